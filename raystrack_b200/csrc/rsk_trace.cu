// rsk_trace.cu -- fused ray generation + closest-hit / any-hit traversal + tally kernels.
//
// One CTA processes one tile (512..8192 consecutive rays, chosen per launch) of one (emitter, iteration) job.  Rays are
// generated in registers (rsk_raygen.cuh), traced against the 8-wide quantised BVH (or, without BVH, against all
// triangles in input order) and tallied into a per-CTA shared-memory histogram that is flushed with one global
// atomic per touched bin.  Rays never touch HBM.
//
// Replaces, per iteration, the reference's launch sequence kernel_build_rays -> kernel_zero -> kernel_trace_* ->
// kernel_reduce_hits (main.py:633-671) and kernel_trace_[bvh_]tregenza / _count_upward (main.py:2064-2088).
#include "rsk_trace.cuh"


constexpr unsigned FULL = 0xffffffffu;
// RSK_COUNTERS=1 (diagnostic builds, scripts/kernel_variants.py): per-launch work counters read with rsk_trace_counters --
// [0] node visits, [1] triangle tests, [2] triangles skipped by the surface mask, [3] rays, [4] triangle-flush trips,
// [5] stack pushes.  The product build compiles them out.
#ifndef RSK_COUNTERS
#define RSK_COUNTERS 0
#endif
__device__ unsigned long long rsk_work_counters[8];
#if RSK_COUNTERS
#define RSK_COUNT(i) (++cnt[i])
#else
#define RSK_COUNT(i) ((void)0)
#endif
#ifndef RSK_MIN_CTAS_PER_SM
#define RSK_MIN_CTAS_PER_SM 4
#endif
#ifndef RSK_POSTPONE
#define RSK_POSTPONE 10       // closest-hit walks test their pending triangle groups once this many lanes hold one
#endif
#ifndef RSK_POSTPONE_IDLE
#define RSK_POSTPONE_IDLE 4   // ... or once this many lanes have nothing else to do
#endif
#ifndef RSK_TRI_PAIRS
#define RSK_TRI_PAIRS 2       // triangles a closest-hit walk fetches per trip of its triangle loop (0: one, loaded inside the test)
#endif
#ifndef RSK_TRI_PAIRS_SKY
#define RSK_TRI_PAIRS_SKY 1   // the any-hit walk does the same in its immediate triangle loop
#endif
#ifndef RSK_REFILL_BELOW
#define RSK_REFILL_BELOW 24   // leave the traversal loop to fetch new rays when fewer lanes are busy (24-27 measured best)
#endif
constexpr int RSK_MIN_CTAS = RSK_MIN_CTAS_PER_SM;   // 4 CTAs x 256 threads per SM -> at most 64 registers per thread
constexpr int REFILL_BELOW = RSK_REFILL_BELOW;
constexpr int RAY_SLOTS = 64;                       // per-warp ray buffer: a top-up adds <= 32 rays to < 32 leftovers
constexpr uint32_t STACK_STRIDE = RSK_TILE_THREADS * sizeof(uint2);      // bytes between two stack levels of a thread


// Result bin of a finished ray, -1 = nothing to tally.  Matrix bins: 2*receiver + (front ? 0 : 1) (cpu_trace.py:114);
// sky bins (after the matrix bins in MODE_DUAL): Tregenza patch or the single "Sky" counter (cpu_trace.py:735-798).
__device__ __forceinline__ int rsk_result_key(const TraceArgs &a, const Walk &w, bool want_m, bool want_s, bool any_hit,
                                              int sky_base, int n_sky) {
    if (want_m && w.best_tri >= 0) {
        const float4 n = __ldg(a.sc.nrm + w.best_tri);
        const int sid = __float_as_int(n.w);
        const bool front = -(w.dx * n.x + w.dy * n.y + w.dz * n.z) > 0.0f;
        return 2 * sid + (front ? 0 : 1);
    }
    if (want_s && !any_hit && w.best_tri < 0) {
        const int patch = n_sky == 1 ? (w.dz > 0.0f ? 0 : -1) : rsk_tregenza_patch(w.dx, w.dy, w.dz);
        return patch < 0 ? -1 : sky_base + patch;
    }
    return -1;
}

template <int MODE>
__device__ __forceinline__ void rsk_debug_store(const TraceArgs &a, int64_t k, const Walk &w, int key, bool any_hit) {
    const int64_t i = k - a.dbg_base;
    if (a.dbg_orig) { a.dbg_orig[3 * i] = w.ox; a.dbg_orig[3 * i + 1] = w.oy; a.dbg_orig[3 * i + 2] = w.oz; }
    if (a.dbg_dirs) { a.dbg_dirs[3 * i] = w.dx; a.dbg_dirs[3 * i + 1] = w.dy; a.dbg_dirs[3 * i + 2] = w.dz; }
    if (MODE == MODE_MATRIX) {
        if (a.dbg_hit) a.dbg_hit[i] = key < 0 ? -1 : (key >> 1);
        if (a.dbg_front) a.dbg_front[i] = (key >= 0 && !(key & 1)) ? 1 : 0;
    } else {
        if (a.dbg_hit) a.dbg_hit[i] = any_hit ? 1 : 0;
        if (a.dbg_front) a.dbg_front[i] = key < 0 ? 255 : (uint8_t)key;
    }
}

// One step of the wide-BVH walk up to the node test: pop a node group when the current one is used up, pick the
// nearest inner child still to visit, push the rest of the group.  Returns false when the walk has nothing left;
// otherwise `node` is the node to test.  Stack entries 0..RSK_SMEM_STACK-1 of a thread live in shared memory and are
// accessed with predicated loads/stores (no branch: the lanes of a warp that pop, push or do neither stay converged);
// deeper entries spill to local memory on a rarely taken path.
__device__ __forceinline__ bool rsk_walk_next(Walk &w, uint32_t stack_base, uint2 *spill, uint32_t &node) {
    const uint32_t smem_end = stack_base + RSK_SMEM_STACK * STACK_STRIDE;
    const bool used_up = w.ng.y <= 0x00ffffffu;
    const bool bottom = w.sp == stack_base;
    const bool pop = used_up && !bottom;
    if (pop) w.sp -= STACK_STRIDE;
    rsk_lds64_if(w.sp, pop && w.sp < smem_end, w.ng);
    if (pop && w.sp >= smem_end) w.ng = spill[(w.sp - smem_end) / STACK_STRIDE];
    if (used_up && bottom) return false;
    const int bit = 31 - __clz(w.ng.y);
    w.ng.y &= ~(1u << bit);
    const uint32_t slot = (uint32_t)(bit - 24) ^ w.octinv;
    node = w.ng.x + __popc(w.ng.y & 0xffu & ((1u << slot) - 1u));
    const bool push = w.ng.y > 0x00ffffffu;
    rsk_sts64_if(w.sp, push && w.sp < smem_end, w.ng);
    if (push && w.sp >= smem_end) {
        const uint32_t level = (w.sp - smem_end) / STACK_STRIDE;
        if (level < RSK_LOCAL_STACK) spill[level] = w.ng;
    }
    if (push && w.sp < stack_base + RSK_MAX_DEPTH * STACK_STRIDE) w.sp += STACK_STRIDE;
    return true;
}

// MODE_MATRIX: closest hit among the receivers of the job.  MODE_SKY: any hit among the active non-emitter meshes,
// misses binned by direction.  MODE_DUAL: both from one traversal (reference trace_cpu_[bvh_]combined,
// cpu_trace.py:280-522): the closest receiver hit bounds the walk, every hit of an active mesh marks the ray as
// occluded; a job whose matrix (sky) side has converged degrades to the sky-only (matrix-only) walk, exactly as the
// reference's shared-ray loop does (main.py:1380-1547).
template <int MODE, bool BVH>
__global__ void __launch_bounds__(RSK_TILE_THREADS, RSK_MIN_CTAS) rsk_trace_kernel(const TraceArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_next;      // next unclaimed ray of the tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_next = 0;
    const TileDesc td = a.tiles[blockIdx.x];
    const long long t_begin = a.job_ticks ? clock64() : 0ll;
    __syncthreads();
    const int job = td.job;
    const bool side1_done = a.done && a.done[job];
    const bool want_m = MODE == MODE_MATRIX ? true : (MODE == MODE_DUAL ? !side1_done : false);
    const bool want_s = MODE == MODE_SKY ? true : (MODE == MODE_DUAL ? !a.done2[job] : false);
    if (MODE == MODE_DUAL ? (!want_m && !want_s) : side1_done) return;

    const EmitterDesc e = a.ev.desc[a.emit_ids[job]];
    if (e.n_tri <= 0) return;          // a mesh without triangles shoots nothing (the host never schedules it either)
    const int64_t begin = td.begin;
    const int64_t end = begin + td.count;

    float cp[7];
    {
        // both sides of a dual job run the same iteration number while both are alive
        const int it = a.iter_index >= 0 ? a.iter_index : ((MODE == MODE_DUAL && !want_m) ? a.iters_done2[job] : a.iters_done[job]);
        const float *row = a.cp_table + 7 * (int64_t)(a.rot_base[job] + it);
#pragma unroll
        for (int i = 0; i < 7; ++i) cp[i] = __ldg(row + i);
    }

    // shared memory: [traversal stacks][priority table][ray buffers][occluder mask][receiver mask (dual only)][histogram]
    // (the fixed-size parts come first, so the addresses the walk loop uses are compile-time offsets)
    constexpr int STACK_WORDS = BVH ? RSK_SMEM_STACK * RSK_TILE_THREADS * 2 : 0;
    constexpr int LUT_WORDS = BVH ? RSK_PERM_LUT_BYTES / 4 : 0;
    constexpr int RAY_WORDS = (RSK_TILE_THREADS / 32) * 7 * RAY_SLOTS;
    const int mw = a.sc.mask_words;
    const int n_hist_all = a.n_hist + (MODE == MODE_DUAL ? a.n_hist2 : 0);
    uint32_t *s_mask = smem + STACK_WORDS + LUT_WORDS + RAY_WORDS;     // surfaces that stop / receive rays
    uint32_t *s_recv = MODE == MODE_DUAL ? s_mask + mw : s_mask;       // surfaces the matrix may tally
    uint32_t *s_hist = s_mask + (MODE == MODE_DUAL ? 2 * mw : mw);
    const int hist_words = a.hist_in_smem ? n_hist_all : 0;
    {
        // dual jobs whose sky side is finished walk with the receiver mask only (ineligible meshes are invisible)
        const uint32_t *occ = (MODE == MODE_DUAL && want_s) ? a.surf_mask2 : a.surf_mask;
        for (int i = tid; i < mw; i += RSK_TILE_THREADS) {
            s_mask[i] = occ[(int64_t)job * mw + i];
            if (MODE == MODE_DUAL) s_recv[i] = a.surf_mask[(int64_t)job * mw + i];
        }
    }
    for (int i = tid; i < hist_words; i += RSK_TILE_THREADS) s_hist[i] = 0;
    if (BVH) rsk_build_perm_lut(reinterpret_cast<uint8_t *>(smem + STACK_WORDS), tid);
    __syncthreads();
    const uint32_t stack_base = rsk_smem_addr(smem) + tid * (uint32_t)sizeof(uint2);
    const uint32_t lut_base = rsk_smem_addr(smem + STACK_WORDS);
    const uint32_t mask_addr = rsk_smem_addr(s_mask), recv_addr = rsk_smem_addr(s_recv);

    const int job_min_sid = (a.min_sid && !(MODE == MODE_DUAL && want_s)) ? a.min_sid[job] : 0;
    unsigned long long *g_tally = a.tally ? a.tally + (int64_t)job * a.n_hist : nullptr;
    unsigned long long *g_tally2 = (MODE == MODE_DUAL && a.tally2) ? a.tally2 + (int64_t)job * a.n_hist2 : nullptr;
    const int sky_base = MODE == MODE_DUAL ? a.n_hist : 0;
    const int n_sky = MODE == MODE_DUAL ? a.n_hist2 : a.n_hist;
    const bool debug_out = MODE != MODE_DUAL && (a.dbg_hit || a.dbg_orig || a.dbg_dirs || a.dbg_front);

    const int tile_n = (int)(end - begin);
    // hand-out units of the tile's rays (see the refill below)
    const int class_mod = (a.class_mod > 1 && tile_n >= 24 * a.class_mod) ? a.class_mod : 1;
    const int class_phase = (int)((begin + 1) % class_mod);
    const int n_units = class_mod * (((tile_n + class_mod - 1) / class_mod + 31) / 32);
    float *s_rays = reinterpret_cast<float *>(smem + STACK_WORDS + LUT_WORDS) + warp * (7 * RAY_SLOTS);
    int buf_n = 0;
    bool pool_empty = tile_n <= 0;
    Walk w;
    bool active = false;
    bool any_hit = false;     // the current ray has met an occluder (sky / dual modes); lives as long as the ray
    int key = -1;             // finished-ray result waiting to be tallied
    TriGroup ptg = {0u, 0u, 0u};   // triangle group found but not tested yet (lives as long as the ray)
    int my_pos = 0;           // the current ray is ray begin + my_pos of the emitter
    uint2 spill[BVH ? RSK_LOCAL_STACK : 1];
#if RSK_COUNTERS
    unsigned cnt[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#endif

    // One triangle of a pending group against the current ray (Moeller-Trumbore, cpu_trace.py:88-114).  Returns true
    // when the ray is settled by it (sky-only walks stop at the first hit).
    auto test_loaded = [&](int tri, const float4 &V0, const float4 &E1, const float4 &E2) -> bool {
        const int sid = __float_as_int(V0.w);
        if (!rsk_surface_on_s(mask_addr, sid)) { RSK_COUNT(2); return false; }
        RSK_COUNT(1);
        float t;
        if (!rsk_tri_hit(V0, E1, E2, w.ox, w.oy, w.oz, w.dx, w.dy, w.dz, t) || !(t > 1e-6f)) return false;
        if (want_s) {
            any_hit = true;
            if (!want_m) return true;
        }
        if (want_m && t < w.best && (MODE != MODE_DUAL || rsk_surface_on_s(recv_addr, sid))) { w.best = t; w.best_tri = tri; }
        return false;
    };
    auto test_triangle = [&](int tri) -> bool {
        const float4 *tp = a.sc.tri + 3 * (int64_t)tri;
        const float4 V0 = __ldg(tp), E1 = __ldg(tp + 1), E2 = __ldg(tp + 2);
        const int sid = __float_as_int(V0.w);
        if (!rsk_surface_on_s(mask_addr, sid)) { RSK_COUNT(2); return false; }
        RSK_COUNT(1);
        float t;
        if (!rsk_tri_hit(V0, E1, E2, w.ox, w.oy, w.oz, w.dx, w.dy, w.dz, t) || !(t > 1e-6f)) return false;
        if (want_s) {
            any_hit = true;
            if (!want_m) return true;
        }
        if (want_m && t < w.best && (MODE != MODE_DUAL || rsk_surface_on_s(recv_addr, sid))) { w.best = t; w.best_tri = tri; }
        return false;
    };

    for (;;) {
        // ---- warp-aggregated tally of the rays finished since the last refill
        const unsigned has = __ballot_sync(FULL, key >= 0);
        if (key >= 0) {
            const unsigned peers = __match_any_sync(has, key);
            if ((__ffs(peers) - 1) == lane) {
                if (a.hist_in_smem) atomicAdd(&s_hist[key], (uint32_t)__popc(peers));
                else if (key < a.n_hist || MODE != MODE_DUAL) { if (g_tally) atomicAdd(&g_tally[key], (unsigned long long)__popc(peers)); }
                else if (g_tally2) atomicAdd(&g_tally2[key - a.n_hist], (unsigned long long)__popc(peers));
            }
            key = -1;
        }
        // ---- refill idle lanes with fresh rays.  Rays are generated 32 at a time by the whole warp (the float64
        // sampler then runs with every lane busy and the Halton rows are read coalesced) into a small per-warp buffer;
        // idle lanes take their next ray from it.
        const unsigned need = __ballot_sync(FULL, !active);
        if (need) {
            const int n_need = __popc(need);
            if (buf_n < n_need && !pool_empty) {
                // the rays of the tile are one pool: a warp that runs out takes the next hand-out unit, whichever warp
                // it is.  A unit is up to 32 rays of one residue class of the ray index modulo class_mod (see
                // rsk_pick_class_mod): rays whose Halton digits agree start from the same strip of the emitter and
                // leave into the same azimuth sector, so the lanes of a warp walk neighbouring nodes.
                int unit = 0;
                if (lane == 0) unit = atomicAdd(&s_next, 1);
                unit = __shfl_sync(FULL, unit, 0);
                const int cls = unit % class_mod, sub = unit / class_mod;
                int first = cls - class_phase;          // smallest pos >= 0 with (begin + pos + 1) % class_mod == cls
                if (first < 0) first += class_mod;
                const int pos = first + (sub * 32 + lane) * class_mod;
                const bool valid = unit < n_units && pos < tile_n;
                const int made = __popc(__ballot_sync(FULL, valid));       // valid lanes are a prefix (pos grows with lane)
                if (valid) {
                    const Ray r = rsk_make_ray(a.ev, e, begin + pos, cp);
                    float *slot = s_rays + buf_n + lane;
                    slot[0 * RAY_SLOTS] = r.ox; slot[1 * RAY_SLOTS] = r.oy; slot[2 * RAY_SLOTS] = r.oz;
                    slot[3 * RAY_SLOTS] = r.dx; slot[4 * RAY_SLOTS] = r.dy; slot[5 * RAY_SLOTS] = r.dz;
                    slot[6 * RAY_SLOTS] = __int_as_float(pos);
                }
                pool_empty = unit + 1 >= n_units;
                buf_n += made;
                __syncwarp();
            }
            const int rank = __popc(need & ((1u << lane) - 1u));
            if (!active && rank < buf_n) {
                const float *slot = s_rays + (buf_n - 1 - rank);
                Ray r;
                r.ox = slot[0 * RAY_SLOTS]; r.oy = slot[1 * RAY_SLOTS]; r.oz = slot[2 * RAY_SLOTS];
                r.dx = slot[3 * RAY_SLOTS]; r.dy = slot[4 * RAY_SLOTS]; r.dz = slot[5 * RAY_SLOTS];
                my_pos = __float_as_int(slot[6 * RAY_SLOTS]);
                rsk_walk_begin(w, r, stack_base, lut_base);
                RSK_COUNT(3);
                active = true;
                any_hit = false;
                ptg.hits = 0u;
            }
            buf_n -= min(n_need, buf_n);
        }
        if (!__any_sync(FULL, active)) break;
        const bool rays_left = buf_n > 0 || !pool_empty;

        if (BVH && MODE != MODE_SKY) {
            // ---- closest hit: 8-wide BVH walk, one node step per loop trip.  The triangles a node step uncovers are
            // kept as one pending group per lane while the lane goes on stepping nodes (culling with a slightly stale
            // closest hit); the warp tests the pending groups together once RSK_POSTPONE lanes hold one, a lane holds
            // two, or RSK_POSTPONE_IDLE lanes have nothing else to do.  Tested at once, the triangle loop runs one or
            // two lanes wide on nearly every trip (25 % of the issued instructions); batched it is +8 % rays/s.
            // (Postponing until the whole warp reconverges, "while-while", is 30 % slower: lanes idle through other
            // lanes' node steps.)
            bool done = false;
            const int flush_at = want_m ? RSK_POSTPONE : 1;
            while (active) {
                uint32_t node;
                const bool more = rsk_walk_next(w, stack_base, spill, node);
                TriGroup tg = {0u, 0u, 0u};
                if (more) {
                    RSK_COUNT(0);
                    uint2 ng2;
                    rsk_test_node(a.sc.nodes, node, w, want_m ? w.best : RSK_INF, ng2, tg, mask_addr, job_min_sid);
                    w.ng = ng2;
                }
                if (!ptg.hits) { ptg = tg; tg.hits = 0u; }
                const unsigned in_loop = __activemask();
                const unsigned pend = __ballot_sync(in_loop, ptg.hits != 0u);
                const unsigned full = __ballot_sync(in_loop, tg.hits != 0u);
                const unsigned idle = __ballot_sync(in_loop, !more);
                if (__popc(pend) >= flush_at || full || __popc(idle) >= RSK_POSTPONE_IDLE || idle == in_loop) {
                    if (ptg.hits) RSK_COUNT(4);
#if RSK_TRI_PAIRS
                    // RSK_TRI_PAIRS (2) triangles per trip: all are fetched before any is tested, so a lane with several
                    // pending triangles waits for memory once per trip instead of once per triangle (the loop is
                    // latency-bound: a third of the kernel's stall samples at ~9 active lanes)
                    while (ptg.hits) {
                        int tri[RSK_TRI_PAIRS];
                        float4 T0[RSK_TRI_PAIRS], T1[RSK_TRI_PAIRS], T2[RSK_TRI_PAIRS];
                        bool have[RSK_TRI_PAIRS];
#pragma unroll
                        for (int u = 0; u < RSK_TRI_PAIRS; ++u) {
                            have[u] = ptg.hits != 0u;
                            const int b = have[u] ? __ffs(ptg.hits) - 1 : 0;
                            ptg.hits &= ptg.hits - 1u;
                            tri[u] = (int)(ptg.base + __popc(ptg.leaf_bits & ((1u << b) - 1u)));
                            const float4 *tp = a.sc.tri + 3 * (int64_t)tri[u];
                            T0[u] = __ldg(tp); T1[u] = __ldg(tp + 1); T2[u] = __ldg(tp + 2);
                        }
                        bool stop = false;
#pragma unroll
                        for (int u = 0; u < RSK_TRI_PAIRS; ++u)
                            if (!stop && have[u] && test_loaded(tri[u], T0[u], T1[u], T2[u])) stop = true;
                        if (stop) { ptg.hits = 0u; break; }
                    }
#else
                    while (ptg.hits) {
                        const int b = __ffs(ptg.hits) - 1;
                        ptg.hits &= ptg.hits - 1u;
                        if (test_triangle((int)(ptg.base + __popc(ptg.leaf_bits & ((1u << b) - 1u))))) { ptg.hits = 0u; break; }
                    }
#endif
                    if (tg.hits) ptg = tg;
                }
                if ((any_hit && !want_m) || (w.ng.y <= 0x00ffffffu && w.sp == stack_base && !ptg.hits)) {
                    done = true;
                    active = false;
                    break;
                }
                if (rays_left && __popc(__activemask()) < REFILL_BELOW) break;
            }
            if (done) {      // the rays finished during these trips are classified together, after the loop has reconverged
                key = rsk_result_key(a, w, want_m, want_s, any_hit, sky_base, n_sky);
                if (debug_out) rsk_debug_store<MODE>(a, begin + my_pos, w, key, any_hit);
            }
        } else if (BVH) {
            // ---- any hit: the triangles a step uncovers are tested at once (a ray ends at its first hit, so postponing
            // only adds node steps: 3.53 vs 3.71 Grays/s)
            while (active) {
                uint32_t node;
                bool finished = !rsk_walk_next(w, stack_base, spill, node);
                if (!finished) {
                    RSK_COUNT(0);
                    uint2 ng2;
                    TriGroup tg;
                    rsk_test_node(a.sc.nodes, node, w, RSK_INF, ng2, tg, mask_addr, job_min_sid);
                    w.ng = ng2;
#if RSK_TRI_PAIRS_SKY
                    while (tg.hits) {       // two triangles per trip, fetched together (see the closest-hit loop)
                        const int b0 = __ffs(tg.hits) - 1;
                        tg.hits &= tg.hits - 1u;
                        const bool two = tg.hits != 0u;
                        const int b1 = two ? __ffs(tg.hits) - 1 : b0;
                        tg.hits &= tg.hits - 1u;
                        const int tri0 = (int)(tg.base + __popc(tg.leaf_bits & ((1u << b0) - 1u)));
                        const int tri1 = (int)(tg.base + __popc(tg.leaf_bits & ((1u << b1) - 1u)));
                        const float4 *tp0 = a.sc.tri + 3 * (int64_t)tri0, *tp1 = a.sc.tri + 3 * (int64_t)tri1;
                        const float4 A0 = __ldg(tp0), A1 = __ldg(tp0 + 1), A2 = __ldg(tp0 + 2);
                        const float4 B0 = __ldg(tp1), B1 = __ldg(tp1 + 1), B2 = __ldg(tp1 + 2);
                        if (test_loaded(tri0, A0, A1, A2) || (two && test_loaded(tri1, B0, B1, B2))) { finished = true; break; }
                    }
#else
                    while (tg.hits) {
                        const int b = __ffs(tg.hits) - 1;
                        tg.hits &= tg.hits - 1u;
                        if (test_triangle((int)(tg.base + __popc(tg.leaf_bits & ((1u << b) - 1u))))) { finished = true; break; }
                    }
#endif
                }
                if (finished) {
                    key = rsk_result_key(a, w, want_m, want_s, any_hit, sky_base, n_sky);
                    if (debug_out) rsk_debug_store<MODE>(a, begin + my_pos, w, key, any_hit);
                    active = false;
                    break;
                }
                if (rays_left && __popc(__activemask()) < REFILL_BELOW) break;
            }
        } else {
            // ---- no BVH: every triangle in input order, strict t<best (utils/cpu_trace.py:54-117, 280-352, 540-583)
            if (active) {
                for (int tri = 0; tri < a.sc.n_tri; ++tri)
                    if (test_triangle(tri)) break;
                key = rsk_result_key(a, w, want_m, want_s, any_hit, sky_base, n_sky);
                if (debug_out) rsk_debug_store<MODE>(a, begin + my_pos, w, key, any_hit);
                active = false;
            }
        }
        __syncwarp();
    }

#if RSK_COUNTERS
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        unsigned v = cnt[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0 && v) atomicAdd(&rsk_work_counters[i], (unsigned long long)v);
    }
#endif
    if (a.job_ticks && tid == 0) atomicAdd(&a.job_ticks[job], (unsigned long long)(clock64() - t_begin));
    // ---- flush the CTA histogram: one global atomic per touched bin
    if (a.hist_in_smem) {
        __syncthreads();
        for (int i = tid; i < n_hist_all; i += RSK_TILE_THREADS) {
            const uint32_t v = s_hist[i];
            if (!v) continue;
            if (i < a.n_hist) { if (g_tally) atomicAdd(&g_tally[i], (unsigned long long)v); }
            else if (g_tally2) atomicAdd(&g_tally2[i - a.n_hist], (unsigned long long)v);
        }
    }
}

// ----------------------------------------------------------------------------- host launcher

static size_t rsk_trace_smem(const TraceArgs &a, bool bvh, bool dual) {
    const size_t hist = (size_t)a.n_hist + (dual ? a.n_hist2 : 0);
    const size_t words = (size_t)a.sc.mask_words * (dual ? 2 : 1) + (a.hist_in_smem ? hist : 0);
    return (bvh ? (size_t)RSK_SMEM_STACK * RSK_TILE_THREADS * sizeof(uint2) + RSK_PERM_LUT_BYTES : 0)
           + (size_t)(RSK_TILE_THREADS / 32) * 7 * RAY_SLOTS * sizeof(float) + words * 4;
}

template <int MODE, bool BVH>
static int rsk_launch_one(rsk_ctx *ctx, const TraceArgs &a, int64_t n_tiles, cudaStream_t stream) {
    const size_t smem = rsk_trace_smem(a, BVH, MODE == MODE_DUAL);
    if (smem > 48 * 1024)
        RSK_CUDA(cudaFuncSetAttribute(rsk_trace_kernel<MODE, BVH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rsk_trace_kernel<MODE, BVH><<<(unsigned)n_tiles, RSK_TILE_THREADS, smem, stream>>>(a);
    ctx->launches++;
    RSK_CUDA(cudaGetLastError());
    return RSK_OK;
}

int rsk_launch_trace(rsk_ctx *ctx, TraceArgs &a, int mode, int64_t n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0) return RSK_OK;
    if (ctx->l2_flush)          // benchmark mode: evict the L2 before every iteration's trace (counted in the timed region)
        RSK_CUDA(cudaMemsetAsync(ctx->l2_flush, (int)(ctx->launches & 0xff), ctx->l2_flush_bytes, stream));
    RSK_REQUIRE(n_tiles < ((int64_t)1 << 31), "too many ray tiles in one launch");
    // shared-memory histogram when it leaves room for >= 2 CTAs per SM, else warp-aggregated global atomics
    a.hist_in_smem = (((size_t)a.n_hist + (mode == MODE_DUAL ? a.n_hist2 : 0)) * 4 <= 96 * 1024) ? 1 : 0;
    const bool bvh = a.sc.use_bvh != 0;
    a.class_mod = rsk_pick_class_mod(a.tile_rays);
    if (mode == MODE_DUAL) return bvh ? rsk_launch_one<MODE_DUAL, true>(ctx, a, n_tiles, stream) : rsk_launch_one<MODE_DUAL, false>(ctx, a, n_tiles, stream);
    if (mode == MODE_MATRIX) return bvh ? rsk_launch_one<MODE_MATRIX, true>(ctx, a, n_tiles, stream) : rsk_launch_one<MODE_MATRIX, false>(ctx, a, n_tiles, stream);
    return bvh ? rsk_launch_one<MODE_SKY, true>(ctx, a, n_tiles, stream) : rsk_launch_one<MODE_SKY, false>(ctx, a, n_tiles, stream);
}

// Diagnostic work counters of the trace kernels since the last reset (all zero unless the library was built with
// -DRSK_COUNTERS=1): node visits, triangle tests, mask-skipped triangles, rays, triangle-flush trips, stack pushes.
extern "C" int rsk_trace_counters(rsk_ctx *ctx, int64_t *out, int32_t reset) {
    RSK_REQUIRE(ctx && out, "rsk_trace_counters: bad arguments");
    RskScope scope(ctx);
    unsigned long long h[8];
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    RSK_CUDA(cudaMemcpyFromSymbol(h, rsk_work_counters, sizeof(h)));
    for (int i = 0; i < 8; ++i) out[i] = (int64_t)h[i];
    if (reset) {
        memset(h, 0, sizeof(h));
        RSK_CUDA(cudaMemcpyToSymbol(rsk_work_counters, h, sizeof(h)));
    }
    return RSK_OK;
}
