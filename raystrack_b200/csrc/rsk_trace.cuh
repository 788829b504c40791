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

// Per-ray traversal state of the 8-wide BVH walk.
struct Walk {
    float ox, oy, oz, dx, dy, dz;
    float ix, iy, iz;        // 1/d (clamped)
    float best;              // closest accepted t so far (RSK_INF = none)
    int best_tri;            // slot of the closest triangle, -1 = none
    uint2 ng;                // current node group: x = first inner child, y = hit bits<<24 | imask
    int sp;
    uint32_t octinv;         // bit a set: direction component a >= 0
    uint32_t octinv4;        // octinv replicated into four bytes
};

__device__ __forceinline__ float rsk_safe_inv(float d) {
    const float lim = 1e-20f;
    return 1.0f / (fabsf(d) > lim ? d : copysignf(lim, d));
}

__device__ __forceinline__ void rsk_walk_begin(Walk &w, const Ray &r) {
    w.ox = r.ox; w.oy = r.oy; w.oz = r.oz; w.dx = r.dx; w.dy = r.dy; w.dz = r.dz;
    w.ix = rsk_safe_inv(r.dx); w.iy = rsk_safe_inv(r.dy); w.iz = rsk_safe_inv(r.dz);
    w.best = RSK_INF; w.best_tri = -1;
    w.octinv = (r.dx >= 0.0f ? 1u : 0u) | (r.dy >= 0.0f ? 2u : 0u) | (r.dz >= 0.0f ? 4u : 0u);
    w.octinv4 = w.octinv * 0x01010101u;
    w.ng = make_uint2(0u, 0x80000000u);     // pseudo group whose only child is the root
    w.sp = 0;
}

// How the quantised plane bytes become ray parameters, chosen per axis by the bits of RSK_PRMT_AXES:
//   bit clear: t = float(byte) * (cell * 1/d) + (origin - o) * 1/d           -- one I2F (XU pipe) + one FFMA per plane
//   bit set:   one PRMT drops the byte into mantissa bits 8..15 of the float 2.0, i.e. f = 2 + byte * 2^-14 exactly;
//              t = f * A + B with A = cell * 2^14 / d and B = (origin - o)/d - 2A            -- PRMT (ALU) + FFMA, no XU.
//              B carries a rounding error of <= 2^-9 cell, so the near/far biases are moved outwards by 2^-8 cell.
// All 48 conversions through I2F keep the XU pipe 67 % busy, all through PRMT load the ALU pipe instead (same speed);
// one axis through PRMT balances the two pipes: +1.4 % rays/s, identical tallies (profiles/kernel_variants_r1.md).
#ifndef RSK_PRMT_AXES
#if defined(RSK_BYTE_MODE) && RSK_BYTE_MODE == 3
#define RSK_PRMT_AXES 7
#else
#define RSK_PRMT_AXES 4       // z planes through PRMT, x and y through I2F
#endif
#endif

#ifndef RSK_MASK_PIN
#define RSK_MASK_PIN 2      // 2: hit-mask bytes extracted with PRMT from two pinned words (-14 instructions per node test)
#endif
#ifndef RSK_SUBTREE_SKIP
#define RSK_SUBTREE_SKIP 2      // 0 = off, 2 = range word loaded together with the node (shipped)
#endif
// True when no triangle below node `idx` can matter to this ray: every mesh id there is below the job's `min_sid`
// (reciprocity: receivers j <= i are ignored, main.py:1181-1182), or the sub-tree belongs to a single mesh that is
// switched off for this emitter (its own mesh, or a mesh behind its plane, main.py:167-204).  One 8-byte load.
__device__ __forceinline__ bool rsk_node_ignorable(const uint4 *__restrict__ nodes, uint32_t idx, const uint32_t *mask, int min_sid) {
    const int2 r = __ldg(reinterpret_cast<const int2 *>(nodes + RSK_NODE_WORDS * (size_t)idx + 5));
    return r.y < min_sid || (r.x == r.y && !rsk_surface_on(mask, r.x));
}

// Slab test of the 8 quantised child boxes of node `idx` against the ray over [0, tmax].
// Returns the new node group (inner children hit, priority-permuted) and the 24-bit triangle mask.
__device__ __forceinline__ void rsk_test_node(const uint4 *__restrict__ nodes, uint32_t idx, const Walk &w, float tmax,
                                              uint2 &ng, uint2 &tg, const uint32_t *mask, int min_sid) {
    const uint4 *p = nodes + RSK_NODE_WORDS * (size_t)idx;
    const uint4 n0 = __ldg(p), n1 = __ldg(p + 1), n2 = __ldg(p + 2), n3 = __ldg(p + 3), n4 = __ldg(p + 4);
#if RSK_SUBTREE_SKIP == 2
    {   // all six words are requested together; an ignorable sub-tree costs the loads but no test and no descent
        const int2 r = __ldg(reinterpret_cast<const int2 *>(p + 5));
        if (r.y < min_sid || (r.x == r.y && !rsk_surface_on(mask, r.x))) {
            ng = make_uint2(0u, 0u);
            tg = make_uint2(0u, 0u);
            return;
        }
    }
#endif
    const uint32_t imask = n0.w >> 24;
    // Per axis (bit a of RSK_PRMT_AXES): planes through I2F (t = float(byte) * A + B) or dropped into the mantissa of
    // 2.0 with PRMT (f = 2 + byte * 2^-14; t = f * A' + B', near/far biases moved outwards by 2^-8 cell for B's rounding).
#define RSK_AXIS_SETUP(AXIS, shift, org, o, inv, A, BN, BF)                                                   \
    float A, BN, BF;                                                                                           \
    if ((RSK_PRMT_AXES >> AXIS) & 1) {                                                                         \
        A = __uint_as_float((((n0.w >> shift) & 0xffu) + 14u) << 23) * inv;                                    \
        const float c = fmaf(-2.0f, A, (__uint_as_float(org) - o) * inv);                                      \
        const float e = fabsf(A) * 0x1p-22f;                                                                   \
        BN = c - e; BF = c + e;                                                                                \
    } else {                                                                                                   \
        A = __uint_as_float(((n0.w >> shift) & 0xffu) << 23) * inv;                                            \
        BN = (__uint_as_float(org) - o) * inv; BF = BN;                                                        \
    }
    RSK_AXIS_SETUP(0, 0, n0.x, w.ox, w.ix, ax, bnx, bfx)
    RSK_AXIS_SETUP(1, 8, n0.y, w.oy, w.iy, ay, bny, bfy)
    RSK_AXIS_SETUP(2, 16, n0.z, w.oz, w.iz, az, bnz, bfz)
#undef RSK_AXIS_SETUP
#define RSK_PLANE_I2F(word, j) ((float)(((word) >> (8 * (j))) & 0xffu))
#define RSK_PLANE_PRMT(word, j) __uint_as_float(__byte_perm(word, 0x40000000u, 0x7404u | ((unsigned)(j) << 4)))
#define RSK_PLANE(AXIS, word, j) (((RSK_PRMT_AXES >> AXIS) & 1) ? RSK_PLANE_PRMT(word, j) : RSK_PLANE_I2F(word, j))
    // byte planes: n2 = qlo.x[0..7] qlo.y[0..7]; n3 = qlo.z[0..7] qhi.x[0..7]; n4 = qhi.y[0..7] qhi.z[0..7]
    const bool px = w.octinv & 1u, py = w.octinv & 2u, pz = w.octinv & 4u;
    uint32_t hits = 0;
#pragma unroll
    for (int half = 0; half < RSK_FANOUT / 4; ++half) {
        const uint32_t meta = half ? n1.w : n1.z;
        const uint32_t lox = half ? n2.y : n2.x, loy = half ? n2.w : n2.z, loz = half ? n3.y : n3.x;
        const uint32_t hix = half ? n3.w : n3.z, hiy = half ? n4.y : n4.x, hiz = half ? n4.w : n4.z;
        const uint32_t nxw = px ? lox : hix, fxw = px ? hix : lox;
        const uint32_t nyw = py ? loy : hiy, fyw = py ? hiy : loy;
        const uint32_t nzw = pz ? loz : hiz, fzw = pz ? hiz : loz;
        // four meta bytes at once: inner children (meta = 0b001_11sss) get their slot XOR-ed with the octant
        // permutation, leaf children keep their first-triangle bit; empty slots have no bits to contribute
        const uint32_t inner4 = (((meta & (meta << 1)) & 0x10101010u) >> 4) * 0xffu;
        uint32_t index4 = (meta ^ (w.octinv4 & inner4)) & 0x1f1f1f1fu;
        uint32_t bits4 = (meta >> 5) & 0x07070707u;
#if RSK_MASK_PIN
        asm volatile("" : "+r"(index4), "+r"(bits4));       // keep the two words: ptxas otherwise re-derives them per child
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tnx = fmaf(RSK_PLANE(0, nxw, j), ax, bnx), tfx = fmaf(RSK_PLANE(0, fxw, j), ax, bfx);
            const float tny = fmaf(RSK_PLANE(1, nyw, j), ay, bny), tfy = fmaf(RSK_PLANE(1, fyw, j), ay, bfy);
            const float tnz = fmaf(RSK_PLANE(2, nzw, j), az, bnz), tfz = fmaf(RSK_PLANE(2, fzw, j), az, bfz);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
#if RSK_MASK_PIN >= 2
            const uint32_t contrib = __byte_perm(bits4, 0u, 0x4440u | j) << __byte_perm(index4, 0u, 0x4440u | j);
            hits |= (tn <= tf) ? contrib : 0u;
#else
            if (tn <= tf) hits |= ((bits4 >> (8 * j)) & 0xffu) << ((index4 >> (8 * j)) & 0xffu);
#endif
        }
    }
#undef RSK_PLANE
#undef RSK_PLANE_I2F
#undef RSK_PLANE_PRMT
    ng = make_uint2(n1.x, (hits & 0xff000000u) | imask);
    tg = make_uint2(n1.y, hits & 0x00ffffffu);
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
