// rsk_common.cuh -- shared declarations of librsk_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/raystrack_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librsk_b200 is written for sm_100a (B200) only"
#endif

// ----------------------------------------------------------------------------- error plumbing
void rsk_set_error(const char *fmt, ...);

#define RSK_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            rsk_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                    \
                          cudaGetErrorString(err__));                                             \
            return err__ == cudaErrorMemoryAllocation ? RSK_ERR_OOM : RSK_ERR_CUDA;               \
        }                                                                                         \
    } while (0)

#define RSK_REQUIRE(cond, msg)                                          \
    do {                                                                \
        if (!(cond)) {                                                  \
            rsk_set_error("%s (%s:%d)", msg, __FILE__, __LINE__);       \
            return RSK_ERR_INVALID;                                     \
        }                                                               \
    } while (0)

#define RSK_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != RSK_OK) return rc__; \
    } while (0)

// ----------------------------------------------------------------------------- device-side layouts

constexpr int RSK_TILE_THREADS = 256;      // threads per CTA of the trace kernels
constexpr int RSK_TILE_RAYS_MAX = 8192;    // rays per CTA tile: chosen per launch between 512 and this (rsk_pick_tile_rays)
constexpr int RSK_TREGENZA_BINS = 145;     // utils/cuda_trace.py:12
constexpr float RSK_INF = 1.0e20f;         // utils/cpu_trace.py:8
constexpr int RSK_WIDE = 8;                // slots of a wide node
#ifndef RSK_LEAF_MAX_TRIS
#define RSK_LEAF_MAX_TRIS 3
#endif
constexpr int RSK_LEAF_MAX = RSK_LEAF_MAX_TRIS;   // triangles per leaf child (<= 3: three triangle bits per slot)
constexpr int RSK_MAX_DEPTH_HOST = 32;     // traversal stack entries per ray (wide-tree depth limit)

// How the quantised plane bytes of a node become ray parameters, per axis (bit a set = axis a):
//   bit clear: t = float(byte) * (cell / d) + (origin - o) / d                 -- one I2F (XU pipe) + FFMA per plane
//   bit set:   one PRMT drops the byte into mantissa bits 8..15 of the float 2.0, f = 2 + byte * 2^-14 exactly;
//              t = f * (cell * 2^14 / d) + ((origin - o) / d - 2 * cell * 2^14 / d)  -- PRMT (ALU pipe) + FFMA, no XU
// The builder stores the matching per-axis scale (cell, or cell * 2^14) in the node, so the kernel and rsk_bvh.cu
// share this constant.
#ifndef RSK_PRMT_AXES
#define RSK_PRMT_AXES 4       // z planes through PRMT, x and y through I2F (balances the XU and ALU pipes)
#endif

// One emitter mesh.  Triangle rows live in EmitterSet::tri[5][*] starting at tri_off.
struct EmitterDesc {
    int32_t tri_off;
    int32_t n_tri;
    int32_t g;
    int32_t grid_off;        // offset of this g's jitter table in EmitterSet::grid (float2 per cell)
    int64_t n_rays_once;     // g*g*rays_per_cell
};

// 96-byte compressed 8-wide BVH node: three 32-byte sectors, fetched with three LDG.256.
//   sector 0  everything the walk needs besides the boxes: quantisation origin, first inner child, first triangle,
//             which slots hold inner nodes (imask) / how many triangles each leaf slot holds (leaf_bits), and the
//             range of mesh ids below the node (lets a ray drop whole sub-trees of surfaces it must ignore)
//   sector 1  low planes x, y, z and high planes x of the 8 child boxes, one byte per slot
//   sector 2  high planes y, z; per-axis scale of the byte grid (RSK_PRMT_AXES picks cell or cell * 2^14)
// Slot s is an inner node when bit s of imask is set (inner children are contiguous from child_base, in slot order);
// otherwise bits 3s..3s+2 of leaf_bits hold its triangle count in unary (0 = empty slot).  The triangles of a node
// are contiguous from tri_base in slot order: triangle bit b is triangle tri_base + popc(leaf_bits & ((1 << b) - 1)).
struct __align__(32) WideNode {
    float ox, oy, oz;        // quantisation origin = node box minimum
    uint32_t child_base;     // index of the first inner child
    uint32_t tri_base;       // index of the first triangle referenced by this node's leaf children
    uint32_t leaf_imask;     // leaf_bits (24 bits) | imask << 24
    int32_t sid_min, sid_max;   // smallest / largest mesh id of the triangles below this node
    uint8_t qlo[3][8];       // quantised child boxes, [axis][slot]
    uint8_t qhi[3][8];
    float scale[3];          // per axis: cell size 2^e, times 2^14 on RSK_PRMT_AXES axes
    uint32_t reserved;
};
static_assert(sizeof(WideNode) == 96, "WideNode must be 96 bytes");
constexpr int RSK_NODE_WORDS = sizeof(WideNode) / 16;

// Scene as the trace kernels see it.
struct SceneView {
    const float4 *tri;       // 3 float4 per triangle: (v0, sid bits) (e1, -) (e2, -); traversal order
    const float4 *nrm;       // (unit normal, -) per triangle, same order
    const uint4 *nodes;      // WideNode array as RSK_NODE_WORDS x uint4 (null without BVH)
    int32_t n_tri;
    int32_t n_surf;
    int32_t use_bvh;
    int32_t mask_words;      // ceil(n_surf/32)
};

struct EmitterView {
    const EmitterDesc *desc;
    const float4 *tri;       // 5 float4 per emitter triangle: (a,eps) (e1,n.x) (e2,n.y) (u,n.z) (v,0)
    const float *cdf;
    const float2 *grid;      // per-cell jitter (u,v)
    const float *halton;     // 5 rows of `halton_stride` floats: bases 5,2,3,7,11
    int64_t halton_stride;
    int32_t rays_per_cell;
};

// ----------------------------------------------------------------------------- host objects

struct rsk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaMemPool_t pool = nullptr;     // private stream-ordered memory pool (see rsk_dev_alloc)
    cudaStream_t stream2 = nullptr;   // second stream of pipelined solves (odd iterations), created on first use
    cudaEvent_t ev_join = nullptr;    // orders `stream` after the work enqueued on stream2
    void *l2_flush = nullptr;         // bench.py: scratch written before every trace launch (rsk_ctx_set_l2_flush)
    size_t l2_flush_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    int sm_count = 0;
    int32_t *h_pinned = nullptr;      // small pinned scratch for asynchronous read-backs (allocated once: cudaFreeHost
                                      // synchronises the whole device, so it must not sit on the per-solve path)
    // QMC caches (device)
    float *halton = nullptr;          // [5][halton_cap]
    int64_t halton_cap = 0;
    std::map<int, std::pair<int64_t, int64_t>> grid_index;   // g -> (offset, cells) inside grid
    float2 *grid = nullptr;
    int64_t grid_cap = 0, grid_used = 0;
    // multi-GPU (rsk_comm.cu): NCCL communicator of this context, small device scratch for host-value all-reduces
    void *comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    long long *comm_scratch = nullptr;
    // result rows in compressed form (rsk_csr_build): column index and F = tally / total rays of every non-zero bin
    int32_t *csr_cols = nullptr;
    double *csr_vals = nullptr;
    int64_t csr_nnz = 0;
    // pinned staging area for large downloads (grow-only)
    void *stage = nullptr;
    size_t stage_cap = 0;
};
constexpr int RSK_COMM_SCRATCH = 4096;     // int64 elements of rsk_ctx::comm_scratch
int rsk_ctx_stage(rsk_ctx *ctx, size_t bytes, void **stage);

struct rsk_scene {
    rsk_ctx *ctx = nullptr;
    int64_t n_tri = 0;
    int32_t n_surf = 0;
    int32_t use_bvh = 0;
    float4 *tri = nullptr;
    float4 *nrm = nullptr;
    uint4 *nodes = nullptr;
    int32_t *tri_index = nullptr;     // traversal slot -> input triangle
    int64_t n_nodes = 0;
    int32_t depth = 0;
    int64_t build_us = 0;
    SceneView view() const {
        SceneView v;
        v.tri = tri; v.nrm = nrm; v.nodes = nodes; v.n_tri = (int32_t)n_tri; v.n_surf = n_surf;
        v.use_bvh = use_bvh; v.mask_words = (n_surf + 31) / 32;
        return v;
    }
};

struct rsk_emitters {
    rsk_ctx *ctx = nullptr;
    int32_t n_emit = 0;
    int32_t rays_per_cell = 0;
    int64_t n_tri_total = 0;
    int64_t max_rays_once = 0;
    std::vector<EmitterDesc> h_desc;
    EmitterDesc *desc = nullptr;
    float4 *tri = nullptr;
    float *cdf = nullptr;
    EmitterView view() const {
        EmitterView v;
        v.desc = desc; v.tri = tri; v.cdf = cdf; v.grid = ctx->grid; v.halton = ctx->halton;
        v.halton_stride = ctx->halton_cap; v.rays_per_cell = rays_per_cell;
        return v;
    }
};

// One CTA's share of a launch: rays [begin, begin + count) of the emitter of local job `job`.
struct TileDesc {
    int32_t job;
    int32_t count;
    int64_t begin;
};

// Arguments of the fused trace kernels (rsk_trace.cu).
struct TraceArgs {
    SceneView sc;
    EmitterView ev;
    const int32_t *emit_ids;        // [n_local]
    const TileDesc *tiles;          // [n_tiles] one CTA each (rsk_build_tiles)
    int32_t n_local;
    int32_t tile_rays;              // rays of the launch's regular tiles (the tail of the launch uses smaller ones)
    int32_t class_mod;              // hand-out units are residue classes of the ray index modulo this (1 = consecutive rays)
    const uint32_t *surf_mask;      // [n_local][mask_words], bit set = surface is a receiver/occluder
    const float *cp_table;          // [n_rot][7]
    const int32_t *rot_base;        // [n_local]
    const int32_t *iters_done;      // [n_local] (device state)
    int32_t iter_index;             // >= 0: the iteration every running job is at (pipelined solves), else read iters_done
    const int32_t *done;            // [n_local] (device state), may be null
    const int32_t *min_sid;         // [n_local] surfaces with a smaller id are ignored (reciprocity), may be null (= 0)
    unsigned long long *tally;      // [n_local][n_hist]
    // MODE_DUAL only: the sky side of the same jobs (any-hit occluder mask, its own progress and tallies)
    const uint32_t *surf_mask2;     // [n_local][mask_words], bit set = active non-emitter surface
    const int32_t *iters_done2;     // [n_local]
    const int32_t *done2;           // [n_local]
    unsigned long long *tally2;     // [n_local][n_hist2]
    int32_t n_hist2;                // 145 or 1
    int32_t n_hist;                 // matrix: 2*n_surf; sky: 145 or 1
    int32_t hist_in_smem;
    const int64_t *ray_begin;       // [n_local] first ray of each job's slice (null: 0)
    const int64_t *ray_end;         // [n_local] one past the last ray of each job's slice (null: n_rays_once)
    unsigned long long *job_ticks;  // [n_local] optional: SM clock ticks the CTAs of each job were resident (rsk_emitter_costs)
    int64_t dbg_base;               // ray index stored at element 0 of the per-ray outputs
    float *dbg_orig, *dbg_dirs;     // optional per-ray outputs (test hook)
    int32_t *dbg_hit;
    uint8_t *dbg_front;
};

enum { MODE_MATRIX = 0, MODE_SKY = 1, MODE_DUAL = 2 };

// internal entry points shared between translation units
int rsk_launch_trace(rsk_ctx *ctx, TraceArgs &a, int mode, int64_t n_tiles, cudaStream_t stream);
int rsk_ctx_stream2(rsk_ctx *ctx, cudaStream_t *out);
int rsk_ctx_join(rsk_ctx *ctx);
// Tile size for a launch over `total_rays` rays: large tiles amortise the per-CTA prologue/flush (8192: +2 % on C5),
// small tiles keep all SMs busy when a scene shoots few rays per iteration and shorten the tail of the launch.
static inline int rsk_pick_tile_rays(int64_t total_rays, int sm_count, bool pipelined = false) {
    static int forced = -1;                 // RSK_TILE_RAYS=<512..8192, power of two>: tuning override
    if (forced < 0) {
        const char *e = getenv("RSK_TILE_RAYS");
        const int v = e ? atoi(e) : 0;
        forced = (v >= 512 && v <= RSK_TILE_RAYS_MAX && (v & (v - 1)) == 0) ? v : 0;
    }
    if (forced) return forced;
    const int64_t wave = 4 * (int64_t)(sm_count > 0 ? sm_count : 148);           // CTAs resident at once (4 per SM)
    int t = RSK_TILE_RAYS_MAX;
    // The last wave leaves the SMs idle for about half a tile's run time: the largest tile only pays with >= 8 waves
    // (a 1/8 shard of the bench scene: 4096-ray tiles +1.5 %), below that at least two full waves are kept.
    // (a pipelined solve fills the last wave with the next iteration's tiles: two waves are enough there, 1/8 shard +0.7 %)
    if (t > 4096 && total_rays / t < (pipelined ? 2 : 8) * wave) t = 4096;
    while (t > 512 && total_rays / t < 2 * wave) t >>= 1;
    return t;
}
// Rays are handed to warps in residue classes of the ray index k modulo M.  The per-ray Halton values are radical
// inverses of k + 1: rays with equal (k + 1) mod 5^a share the leading a base-5 digits of the triangle-pick value (they
// start from the same 1/5^a slice of the emitter's area CDF), rays with equal (k + 1) mod 11 share the leading digit of
// the azimuth value, mod 7 the elevation band.  M = 275 = 25 * 11: 1/25 of the emitter x one azimuth sector of 33
// degrees -- ~30 rays per class in an 8192-ray tile, one warp's worth.  RSK_CLASS_MOD overrides (1 = consecutive rays).
static inline int rsk_pick_class_mod(int tile_rays) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("RSK_CLASS_MOD");
        forced = e ? atoi(e) : 0;
        if (forced < 0 || forced > 4096) forced = 0;
    }
    if (forced) return forced;
    return 1;
}
// Cut the ray ranges [rbeg[k], rend[k]) of n_local jobs into the CTA tiles of one launch: regular tiles of
// rsk_pick_tile_rays() rays in job order, then a tail whose tiles shrink with the work that is left (guided
// self-scheduling: each tail tile takes remaining / (2 x resident CTAs) rays, at least RSK_TAIL_TILE_RAYS) -- all CTA
// slots then run dry within one SMALL tile's time of each other instead of idling for up to a regular tile's time
// (1.4 ms at 8192 rays).  The tail is taken from the ends of the last jobs and amounts to about 1.5 waves of regular
// tiles.  RSK_TAIL_TILE_RAYS overrides the smallest size (default 512; 0 = no tail).
// A pipelined solve (rsk_api.cu) needs no tail: the next iteration's tiles take the slots that run dry.
static inline void rsk_build_tiles(const int64_t *rbeg, const int64_t *rend, int n_local, int sm_count, std::vector<TileDesc> &out,
                                   int *tile_rays_out, bool pipelined = false) {
    static int tail_min = -1;
    if (tail_min < 0) {
        const char *e = getenv("RSK_TAIL_TILE_RAYS");
        tail_min = e ? atoi(e) : 512;
        if (tail_min != 0 && (tail_min < 256 || tail_min > RSK_TILE_RAYS_MAX)) tail_min = 512;
    }
    int64_t total = 0;
    for (int k = 0; k < n_local; ++k) total += rend[k] - rbeg[k];
    const int T = rsk_pick_tile_rays(total, sm_count, pipelined);
    if (tile_rays_out) *tile_rays_out = T;
    const int64_t wave = 4 * (int64_t)(sm_count > 0 ? sm_count : 148);
    int64_t tail = (!pipelined && tail_min > 0 && tail_min < T) ? std::min<int64_t>(total / 2, 3 * wave * T / 2) : 0;
    std::vector<int64_t> take(n_local, 0);
    int64_t left = tail;
    for (int k = n_local - 1; k >= 0 && left > 0; --k) {
        take[k] = std::min<int64_t>(rend[k] - rbeg[k], left);
        left -= take[k];
    }
    out.clear();
    for (int k = 0; k < n_local; ++k)
        for (int64_t b = rbeg[k]; b < rend[k] - take[k]; b += T)
            out.push_back(TileDesc{k, (int32_t)std::min<int64_t>(T, rend[k] - take[k] - b), b});
    int64_t remaining = tail;
    for (int k = 0; k < n_local; ++k) {
        int64_t b = rend[k] - take[k];
        while (b < rend[k]) {
            int64_t sz = remaining / (2 * wave);
            sz = std::max<int64_t>(tail_min, std::min<int64_t>(T, (sz / 256) * 256));
            sz = std::min<int64_t>(sz, rend[k] - b);
            out.push_back(TileDesc{k, (int32_t)sz, b});
            b += sz;
            remaining -= sz;
        }
    }
}
int rsk_qmc_ensure_halton(rsk_ctx *ctx, int64_t n);
int rsk_qmc_ensure_grid(rsk_ctx *ctx, int g, int64_t *offset);
int rsk_bvh_build(rsk_scene *scene, const float4 *tri_in, const float4 *nrm_in);
int rsk_scene_adopt(rsk_ctx *ctx, float4 *d_tri, float4 *d_nrm, int64_t n_tri, int32_t n_surf, int32_t use_bvh, rsk_scene **out);

// Device memory comes from a stream-ordered memory pool that belongs to the context (cudaMallocFromPoolAsync) with an
// unlimited release threshold: repeated solves reuse the same blocks without ever calling cudaMalloc/cudaFree (both of
// which synchronise the device and cost 0.1-0.7 s for the buffers of a million-triangle scene).  The pool is private:
// the device's default pool, which other libraries in the process may use, keeps its settings.  Allocation and
// release are ordered on the stream of the context that is current on the calling thread (RskScope).
extern thread_local cudaStream_t rsk_tl_stream;
extern thread_local cudaMemPool_t rsk_tl_pool;

struct RskScope {
    explicit RskScope(const rsk_ctx *ctx) {
        cudaSetDevice(ctx->device);
        rsk_tl_stream = ctx->stream;
        rsk_tl_pool = ctx->pool;
    }
};

template <typename T>
static inline int rsk_dev_alloc(T **ptr, size_t count) {
    *ptr = nullptr;
    if (count == 0) count = 1;
    if (rsk_tl_pool) RSK_CUDA(cudaMallocFromPoolAsync((void **)ptr, count * sizeof(T), rsk_tl_pool, rsk_tl_stream));
    else RSK_CUDA(cudaMallocAsync((void **)ptr, count * sizeof(T), rsk_tl_stream));
    return RSK_OK;
}

static inline void rsk_dev_free(void *p) {
    if (p) cudaFreeAsync(p, rsk_tl_stream);
}

static inline unsigned rsk_blocks(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
