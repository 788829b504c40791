// rsk_trace.cuh -- ray/triangle and ray/wide-node device code shared by the trace kernels.
#pragma once
#include "rsk_common.cuh"
#include "rsk_raygen.cuh"

#ifndef RSK_SMEM_STACK_N
#define RSK_SMEM_STACK_N 8
#endif
constexpr int RSK_SMEM_STACK = RSK_SMEM_STACK_N;        // stack entries per thread kept in shared memory
constexpr int RSK_LOCAL_STACK = 32 - RSK_SMEM_STACK_N;   // spill entries per thread (local memory)
constexpr int RSK_MAX_DEPTH = RSK_SMEM_STACK + RSK_LOCAL_STACK;
static_assert(RSK_MAX_DEPTH == RSK_MAX_DEPTH_HOST, "stack size mismatch");

// Moeller-Trumbore as the reference writes it (utils/cpu_trace.py:88-110): reject |det| < 1e-7, u in [0,1],
// v >= 0, u+v <= 1; the caller applies the t window.  float32 throughout (the reference promotes
// inv_det,u,v,t to float64; measured per-ray disagreement of the float32 form is ~3e-7).
__device__ __forceinline__ bool rsk_tri_hit(const float4 &V0, const float4 &E1, const float4 &E2,
                                            float ox, float oy, float oz, float dx, float dy, float dz, float &t) {
    // Every product/sum is written as an explicit multiply or fused multiply-add: the compiler may not re-associate
    // or contract them differently in different instantiations of the kernel, so a ray on a triangle's edge gets
    // the same verdict in the matrix, sky and dual kernels.
#define RSK_CROSS(a, b, c, d) __fmaf_rn(a, b, -__fmul_rn(c, d))                       /* a*b - c*d */
#define RSK_DOT(ax, ay, az, bx, by, bz) __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)))
    const float px = RSK_CROSS(dy, E2.z, dz, E2.y);
    const float py = RSK_CROSS(dz, E2.x, dx, E2.z);
    const float pz = RSK_CROSS(dx, E2.y, dy, E2.x);
    const float det = RSK_DOT(E1.x, E1.y, E1.z, px, py, pz);
    if (fabsf(det) < 1e-7f) return false;
    const float inv_det = __fdiv_rn(1.0f, det);
    const float tx = __fsub_rn(ox, V0.x), ty = __fsub_rn(oy, V0.y), tz = __fsub_rn(oz, V0.z);
    const float u = __fmul_rn(RSK_DOT(tx, ty, tz, px, py, pz), inv_det);
    if (u < 0.0f || u > 1.0f) return false;
    const float qx = RSK_CROSS(ty, E1.z, tz, E1.y);
    const float qy = RSK_CROSS(tz, E1.x, tx, E1.z);
    const float qz = RSK_CROSS(tx, E1.y, ty, E1.x);
    const float v = __fmul_rn(RSK_DOT(dx, dy, dz, qx, qy, qz), inv_det);
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    t = __fmul_rn(RSK_DOT(E2.x, E2.y, E2.z, qx, qy, qz), inv_det);
#undef RSK_CROSS
#undef RSK_DOT
    return true;
}

__device__ __forceinline__ bool rsk_surface_on(const uint32_t *mask, int sid) {
    return (mask[sid >> 5] >> (sid & 31)) & 1u;
}

// ---- shared memory through 32-bit shared-window addresses: the compiler otherwise rebuilds the window base
// (S2UR SR_CgaCtaId / UMOV / ULEA) in front of every access of the walk loop.
__device__ __forceinline__ uint32_t rsk_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t rsk_lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t rsk_lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// predicated 8-byte stack accesses: no branch, so the lanes of a warp that push / pop / do neither stay converged
__device__ __forceinline__ void rsk_lds64_if(uint32_t addr, bool pred, uint2 &v) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.shared.v2.u32 {%0, %1}, [%2]; }"
                 : "+r"(v.x), "+r"(v.y) : "r"(addr), "r"((uint32_t)pred));
}
__device__ __forceinline__ void rsk_sts64_if(uint32_t addr, bool pred, const uint2 &v) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.shared.v2.u32 [%2], {%0, %1}; }"
                 :: "r"(v.x), "r"(v.y), "r"(addr), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ bool rsk_surface_on_s(uint32_t mask_addr, int sid) {
    return (rsk_lds32(mask_addr + ((uint32_t)(sid >> 5) << 2)) >> (sid & 31)) & 1u;
}

// Inner children are visited front to back: slot s has priority s ^ octinv (octinv bit a set: direction component a
// is >= 0).  The 8 hit bits of a node test come out in slot order; this 8 x 256 byte table (built once per CTA) maps
// them to priority order: lut[o][m] = OR over the set bits s of m of 1 << (s ^ o).
constexpr int RSK_PERM_LUT_BYTES = 8 * 256;
__device__ __forceinline__ void rsk_build_perm_lut(uint8_t *lut, int tid) {
    for (int i = tid; i < RSK_PERM_LUT_BYTES; i += RSK_TILE_THREADS) {
        const uint32_t o = (uint32_t)i >> 8, m = (uint32_t)i & 255u;
        uint32_t r = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) r |= ((m >> s) & 1u) << (s ^ o);
        lut[i] = (uint8_t)r;
    }
}

// Per-ray traversal state of the 8-wide BVH walk.
struct Walk {
    float ox, oy, oz, dx, dy, dz;
    float ix, iy, iz;        // 1/d (clamped)
    float best;              // closest accepted t so far (RSK_INF = none)
    int best_tri;            // slot of the closest triangle, -1 = none
    uint2 ng;                // current node group: x = first inner child, y = hit bits (priority order) << 24 | imask
    uint32_t sp;             // stack pointer: shared-window address of this thread's next free stack entry
    uint32_t octinv;         // bit a set: direction component a >= 0
    uint32_t lut_row;        // shared-window address of perm_lut[octinv]
};

// Triangles a node test uncovered: tri bit b (set in `hits`) is triangle base + popc(leaf_bits & ((1 << b) - 1)).
struct TriGroup {
    uint32_t base, hits, leaf_bits;
};

__device__ __forceinline__ float rsk_safe_inv(float d) {
    const float lim = 1e-20f;
    return 1.0f / (fabsf(d) > lim ? d : copysignf(lim, d));
}

__device__ __forceinline__ void rsk_walk_begin(Walk &w, const Ray &r, uint32_t stack_base, uint32_t lut_base) {
    w.ox = r.ox; w.oy = r.oy; w.oz = r.oz; w.dx = r.dx; w.dy = r.dy; w.dz = r.dz;
    w.ix = rsk_safe_inv(r.dx); w.iy = rsk_safe_inv(r.dy); w.iz = rsk_safe_inv(r.dz);
    w.best = RSK_INF; w.best_tri = -1;
    w.octinv = (r.dx >= 0.0f ? 1u : 0u) | (r.dy >= 0.0f ? 2u : 0u) | (r.dz >= 0.0f ? 4u : 0u);
    w.lut_row = lut_base + (w.octinv << 8);
    w.ng = make_uint2(0u, 0x80000000u);     // pseudo group whose only child is the root
    w.sp = stack_base;
}

// 32 bytes (one sector) per instruction and lane: sm_100 LDG.256 through the read-only path.  A node is three of
// them; six LDG.128 cost twice the L1 wavefronts (the data stage moves one sector per lane and instruction).
__device__ __forceinline__ void rsk_ldg256(const void *p, uint4 &a, uint4 &b) {
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// Slab test of the 8 quantised child boxes of node `idx` against the ray over [0, tmax].
// Returns the new node group (inner children hit, in priority order) and the triangles of the leaf children hit.
// A sub-tree that holds nothing this ray may hit -- every mesh id below min_sid (reciprocity: receivers j <= i are
// ignored, main.py:1181-1182), or a single mesh that is switched off for this emitter (its own mesh, or a mesh
// behind its plane, main.py:167-204) -- reports no hits.
__device__ __forceinline__ void rsk_test_node(const uint4 *__restrict__ nodes, uint32_t idx, const Walk &w, float tmax,
                                              uint2 &ng, TriGroup &tg, uint32_t mask_addr, int min_sid) {
    const uint4 *p = nodes + RSK_NODE_WORDS * (size_t)idx;
    uint4 a0, a1, b0, b1, c0, c1;
    rsk_ldg256(p, a0, a1);
    rsk_ldg256(p + 2, b0, b1);
    rsk_ldg256(p + 4, c0, c1);
    const int sid_lo = (int)a1.z, sid_hi = (int)a1.w;
    const bool ignore = sid_hi < min_sid || (sid_lo == sid_hi && !rsk_surface_on_s(mask_addr, sid_lo));

    // per axis: t = plane * A + B (near / far biases differ on PRMT axes, whose B carries a rounding error of <= 2^-9 cell)
#define RSK_AXIS_SETUP(AXIS, org, scale, o, inv, A, BN, BF)                                                   \
    const float A = __uint_as_float(scale) * inv;                                                              \
    float BN, BF;                                                                                              \
    if ((RSK_PRMT_AXES >> AXIS) & 1) {                                                                         \
        const float c = fmaf(-2.0f, A, (__uint_as_float(org) - o) * inv);                                      \
        const float e = fabsf(A) * 0x1p-22f;                                                                   \
        BN = c - e; BF = c + e;                                                                                \
    } else {                                                                                                   \
        BN = (__uint_as_float(org) - o) * inv; BF = BN;                                                        \
    }
    RSK_AXIS_SETUP(0, a0.x, c1.x, w.ox, w.ix, ax, bnx, bfx)
    RSK_AXIS_SETUP(1, a0.y, c1.y, w.oy, w.iy, ay, bny, bfy)
    RSK_AXIS_SETUP(2, a0.z, c1.z, w.oz, w.iz, az, bnz, bfz)
#undef RSK_AXIS_SETUP
#define RSK_PLANE_I2F(word, j) ((float)(((word) >> (8 * (j))) & 0xffu))
#define RSK_PLANE_PRMT(word, j) __uint_as_float(__byte_perm(word, 0x40000000u, 0x7404u | ((unsigned)(j) << 4)))
#define RSK_PLANE(AXIS, word, j) (((RSK_PRMT_AXES >> AXIS) & 1) ? RSK_PLANE_PRMT(word, j) : RSK_PLANE_I2F(word, j))
    const bool px = w.octinv & 1u, py = w.octinv & 2u, pz = w.octinv & 4u;
    uint32_t hit8 = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t lox = half ? b0.y : b0.x, loy = half ? b0.w : b0.z, loz = half ? b1.y : b1.x;
        const uint32_t hix = half ? b1.w : b1.z, hiy = half ? c0.y : c0.x, hiz = half ? c0.w : c0.z;
        const uint32_t nxw = px ? lox : hix, fxw = px ? hix : lox;
        const uint32_t nyw = py ? loy : hiy, fyw = py ? hiy : loy;
        const uint32_t nzw = pz ? loz : hiz, fzw = pz ? hiz : loz;
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
            // packed FP32 (Blackwell FFMA2): the plane parameters of two children per instruction
#define RSK_PAIR(AXIS, word, A, B) __ffma2_rn(make_float2(RSK_PLANE(AXIS, word, j), RSK_PLANE(AXIS, word, j + 1)), \
                                              make_float2(A, A), make_float2(B, B))
            const float2 tnx = RSK_PAIR(0, nxw, ax, bnx), tfx = RSK_PAIR(0, fxw, ax, bfx);
            const float2 tny = RSK_PAIR(1, nyw, ay, bny), tfy = RSK_PAIR(1, fyw, ay, bfy);
            const float2 tnz = RSK_PAIR(2, nzw, az, bnz), tfz = RSK_PAIR(2, fzw, az, bfz);
#undef RSK_PAIR
            const float tn0 = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, 0.0f)), tf0 = fminf(fminf(tfx.x, tfy.x), fminf(tfz.x, tmax));
            const float tn1 = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, 0.0f)), tf1 = fminf(fminf(tfx.y, tfy.y), fminf(tfz.y, tmax));
            hit8 |= (tn0 <= tf0) ? (1u << (4 * half + j)) : 0u;
            hit8 |= (tn1 <= tf1) ? (2u << (4 * half + j)) : 0u;
        }
    }
#undef RSK_PLANE
#undef RSK_PLANE_I2F
#undef RSK_PLANE_PRMT
    if (ignore) hit8 = 0u;
    const uint32_t imask = a1.y >> 24;
    ng = make_uint2(a0.w, (rsk_lds8(w.lut_row + (hit8 & imask)) << 24) | imask);
    // every hit bit -> the three triangle bits of its slot (the node's leaf_bits keep those that exist)
    uint32_t x = hit8;
    x = (x | (x << 8)) & 0x00f00fu;
    x = (x | (x << 4)) & 0x0c30c3u;
    x = (x | (x << 2)) & 0x249249u;
    tg.base = a1.x;
    tg.hits = (x * 7u) & a1.y & 0x00ffffffu;
    tg.leaf_bits = a1.y;
}

// Tregenza patch of an upward direction (utils/cpu_trace.py:735-777, float32 arguments).  atan2 is evaluated
// in float64 and rounded, which reproduces a correctly-rounded float32 atan2f; degrees() is a float32 multiply.
__device__ __forceinline__ int rsk_tregenza_patch(float dx, float dy, float dz) {
    if ((double)dz <= 0.0) return -1;
    const double ring_hi[8] = {0.20791169081775934, 0.40673664307580015, 0.5877852522924731, 0.7431448254773942,
                               0.8660254037844386,  0.9510565162951535,  0.9945218953682733, 1.0};
    const int ring_n[8] = {30, 30, 24, 24, 18, 12, 6, 1};
    const int ring_start[8] = {0, 30, 60, 84, 108, 126, 138, 144};
    int ridx = 7;
#pragma unroll
    for (int j = 6; j >= 0; --j)
        if ((double)dz < ring_hi[j]) ridx = j;
    const int n_az = ring_n[ridx], base = ring_start[ridx];
    if (n_az == 1) return base;
    const float at = (float)atan2((double)dy, (double)dx);
    double az = (double)__fmul_rn(at, 57.29577951308232f);
    if (az < 0.0) az = __dadd_rn(az, 360.0);
    const double width = __ddiv_rn(360.0, (double)n_az);
    const double off = (ridx & 1) ? __ddiv_rn(180.0, (double)n_az) : 0.0;
    double t = __dsub_rn(az, off);
    if (t < 0.0) t = __dadd_rn(t, 360.0);
    else if (t >= 360.0) t = __dsub_rn(t, 360.0);
    int aidx = (int)floor(__ddiv_rn(t, width));
    if (aidx >= n_az) aidx = n_az - 1;
    return base + aidx;
}
