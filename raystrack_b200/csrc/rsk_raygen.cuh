// rsk_raygen.cuh -- per-ray QMC emitter sampler + cosine-weighted direction, fused into the trace kernels.
//
// Restates utils/ray_builder.py:25-94 (`build_rays`) with Numba's arithmetic types: float32 table values are
// added to the float32 Cranley-Patterson offsets in float32, everything after the first `% 1.0` is float64, and
// the results are rounded to float32 once, on store.  Every float64 operation is issued as an explicit
// round-to-nearest intrinsic so that nvcc cannot contract a*b+c into an FMA (the CPU reference evaluates
// multiply and add separately).
#pragma once
#include "rsk_common.cuh"

// The five Halton values of a ray are read exactly once per iteration: RSK_HALTON_STREAM=1 loads them with the
// evict-first policy (__ldcs) instead of the read-only path.
#ifndef RSK_HALTON_STREAM
#define RSK_HALTON_STREAM 0
#endif
#if RSK_HALTON_STREAM
#define RSK_HLOAD(p) __ldcs(p)
#else
#define RSK_HLOAD(p) __ldg(p)
#endif

struct Ray {
    float ox, oy, oz;
    float dx, dy, dz;
};

// x % 1.0 for x >= 0 (exact in float64).
__device__ __forceinline__ double rsk_mod1(double x) { return x - floor(x); }

// ray_builder.py:9-22: lower bound of x in the float32 cdf, compared in float64.
__device__ __forceinline__ int rsk_cdf_search(const float *__restrict__ cdf, int n, double x) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if ((double)__ldg(cdf + mid) < x) lo = mid + 1; else hi = mid - 1;
    }
    return lo >= n ? max(n - 1, 0) : lo;
}

__device__ __forceinline__ double rsk_mad3(double a, double x, double b, double y, double c, double z) {
    // (a*x + b*y) + c*z, unfused, source order of ray_builder.py:84-86
    return __dadd_rn(__dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)), __dmul_rn(c, z));
}

// Ray k (0 <= k < n_rays_once) of emitter `e` under rotation cp[0..6] = (cp_grid[0..1], cp_dims[0..4]).
__device__ __forceinline__ Ray rsk_make_ray(const EmitterView &ev, const EmitterDesc &e, int64_t k, const float *cp) {
    // A zero-area emitter (grid_off < 0) has all-zero jitter and Halton tables in the reference (prepared.py:278-287):
    // only the Cranley-Patterson offsets move its rays.
    const bool zero_tables = e.grid_off < 0;
    const int64_t cell = k / ev.rays_per_cell;
    const float2 jit = zero_tables ? make_float2(0.f, 0.f) : __ldg(ev.grid + e.grid_off + cell);
    const double ug = rsk_mod1((double)__fadd_rn(jit.x, cp[0]));                       // :54
    const double vg = rsk_mod1((double)__fadd_rn(jit.y, cp[1]));                       // :55
    const float *h = ev.halton + k;
    const int64_t hs = ev.halton_stride;
#define RSK_HVAL(p) (zero_tables ? 0.0f : RSK_HLOAD(p))
    const double q_tri = rsk_mod1((double)__fadd_rn(RSK_HVAL(h), cp[2]));                  // :57
    const int tri = rsk_cdf_search(ev.cdf + e.tri_off, e.n_tri, q_tri);                // :58
    const double ur = rsk_mod1(__dadd_rn((double)__fadd_rn(RSK_HVAL(h + hs), cp[3]), ug));     // :60
    const double vr = rsk_mod1(__dadd_rn((double)__fadd_rn(RSK_HVAL(h + 2 * hs), cp[4]), vg)); // :61
    const double s = sqrt(ur);                                                         // :63
    const double mix_b = __dmul_rn(s, vr);
    const double mix_c = __dmul_rn(s, __dsub_rn(1.0, vr));

    const float4 *t = ev.tri + 5 * (int64_t)(e.tri_off + tri);
    const float4 A = __ldg(t), E1 = __ldg(t + 1), E2 = __ldg(t + 2), U = __ldg(t + 3), V = __ldg(t + 4);
    const float nx = E1.w, ny = E2.w, nz = U.w, eps = A.w;

    // :71-73  a + mix_b*e1 + mix_c*e2
    const double px = __dadd_rn(__dadd_rn((double)A.x, __dmul_rn(mix_b, (double)E1.x)), __dmul_rn(mix_c, (double)E2.x));
    const double py = __dadd_rn(__dadd_rn((double)A.y, __dmul_rn(mix_b, (double)E1.y)), __dmul_rn(mix_c, (double)E2.y));
    const double pz = __dadd_rn(__dadd_rn((double)A.z, __dmul_rn(mix_b, (double)E1.z)), __dmul_rn(mix_c, (double)E2.z));

    const double r1 = rsk_mod1((double)__fadd_rn(RSK_HVAL(h + 3 * hs), cp[5]));           // :75
    const double r2 = rsk_mod1((double)__fadd_rn(RSK_HVAL(h + 4 * hs), cp[6]));           // :76
    const double sin_t = sqrt(__dsub_rn(1.0, r1));                                     // :78
    const double phi = __dmul_rn(6.283185307179586, r2);
    double sn, cs;
    sincos(phi, &sn, &cs);
    const double x = __dmul_rn(sin_t, cs);
    const double y = __dmul_rn(sin_t, sn);
    const double z = sqrt(r1);

    Ray r;
    r.dx = (float)rsk_mad3(x, (double)U.x, y, (double)V.x, z, (double)nx);             // :84-86
    r.dy = (float)rsk_mad3(x, (double)U.y, y, (double)V.y, z, (double)ny);
    r.dz = (float)rsk_mad3(x, (double)U.z, y, (double)V.z, z, (double)nz);
    r.ox = (float)__dadd_rn(px, (double)__fmul_rn(eps, nx));                           // :89-91
    r.oy = (float)__dadd_rn(py, (double)__fmul_rn(eps, ny));
    r.oz = (float)__dadd_rn(pz, (double)__fmul_rn(eps, nz));
#undef RSK_HVAL
    return r;
}
