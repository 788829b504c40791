/*
 * rsk_oracle.c -- CPU ORACLE for the Raystrack Monte-Carlo view-factor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the reference's
 * Numba CPU kernels (philip-ba/raystrack v1.0.2).  It is the checker for the CUDA
 * path in raystrack_b200/csrc; nothing in the product imports, links or calls it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below
 * against vectors produced by importing the reference itself in the build container
 * (tests/golden/make_golden.py, outputs committed under tests/golden/), and the
 * whole-solve driver (oracle/oracle.py) against the reference's shipped result files.
 *
 * Arithmetic types follow Numba's type inference for the reference source (checked
 * with inspect_types()): float32 array loads stay float32 through +,-,* with other
 * float32 values and are promoted to float64 as soon as they meet a Python float
 * literal (`% 1.0`, `1.0 / det`, `1.0 - vr`, ...).  The reference compiles with
 * fastmath=True (LLVM may contract/reassociate); this file is compiled with
 * -ffp-contract=off and evaluates every expression in source order.
 *
 * All reference citations are relative to /root/reference/src/raystrack/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_INF 1.0e20      /* utils/cpu_trace.py:8 */
#define ORC_STACK 64        /* utils/cpu_trace.py:9 */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ QMC tables */

/* utils/halton.py:9-18 `_halton`: float64 radical inverse, f /= base then r += f*digit. */
static double orc_halton(int64_t i, int64_t base) {
    double f = 1.0, r = 0.0;
    while (i) {
        f /= (double)base;
        r += f * (double)(i % base);
        i /= base;
    }
    return r;
}

/* utils/halton.py:34-39 `_build_halton_dim`: out[i] = float32(H_base(i+1)). */
void orc_halton_dim(int64_t length, int64_t base, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < length; ++i) out[i] = (float)orc_halton(i + 1, base);
}

/* utils/halton.py:21-31 `_build_halton_grid`: u[c]=(H2(c+1)+c//g)/g, v[c]=(H3(c+1)+c%g)/g. */
void orc_halton_grid(int64_t g, float *u, float *v) {
    int64_t cells = g * g;
    for (int64_t c = 0; c < cells; ++c) {
        int64_t i = c / g, j = c % g;
        u[c] = (float)((orc_halton(c + 1, 2) + (double)i) / (double)g);
        v[c] = (float)((orc_halton(c + 1, 3) + (double)j) / (double)g);
    }
}

/* ------------------------------------------------------------------ ray generator */

/* utils/ray_builder.py:9-22 `_binary_search_cdf`: lower bound, float32 cdf vs float64 x. */
static int64_t orc_cdf_search(const float *cdf, int64_t n, double x) {
    int64_t lo = 0, hi = n - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        if ((double)cdf[mid] < x) lo = mid + 1; else hi = mid - 1;
    }
    if (lo < 0) return 0;
    if (lo >= n) return n - 1;
    return lo;
}

/* Python float `%` with a positive divisor (all operands here are >= 0). */
static inline double orc_mod1(double x) {
    double r = fmod(x, 1.0);
    if (r < 0.0) r += 1.0;
    return r;
}

/* utils/ray_builder.py:25-94 `build_rays`.  tri_* are float32 [T,3] row-major. */
void orc_build_rays(const float *u_grid, const float *v_grid,
                    const float *h_tri, const float *h_u, const float *h_v,
                    const float *h_r1, const float *h_r2,
                    const float *cdf, int64_t n_tri,
                    const float *tri_a, const float *tri_e1, const float *tri_e2,
                    const float *tri_u, const float *tri_v, const float *tri_n,
                    const float *tri_eps, int64_t rays_per_cell,
                    int64_t n_rays, float *orig, float *dire,
                    const float *cp_grid, const float *cp_dims) {
    const double two_pi = 6.283185307179586;
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < n_rays; ++idx) {
        int64_t cell = idx / rays_per_cell;
        double ug = orc_mod1((double)(float)(u_grid[cell] + cp_grid[0]));   /* :54 f32+f32, then % 1.0 */
        double vg = orc_mod1((double)(float)(v_grid[cell] + cp_grid[1]));   /* :55 */
        double q_tri = orc_mod1((double)(float)(h_tri[idx] + cp_dims[0]));  /* :57 */
        int64_t tri = orc_cdf_search(cdf, n_tri, q_tri);                    /* :58 */
        double ur = orc_mod1((double)(float)(h_u[idx] + cp_dims[1]) + ug);  /* :60 */
        double vr = orc_mod1((double)(float)(h_v[idx] + cp_dims[2]) + vg);  /* :61 */
        double s = sqrt(ur);                                                /* :63 */
        double mix_b = s * vr;
        double mix_c = s * (1.0 - vr);
        const float *a = tri_a + 3 * tri, *e1 = tri_e1 + 3 * tri, *e2 = tri_e2 + 3 * tri;
        const float *tu = tri_u + 3 * tri, *tv = tri_v + 3 * tri, *tn = tri_n + 3 * tri;
        double px = (double)a[0] + mix_b * (double)e1[0] + mix_c * (double)e2[0];  /* :71-73 */
        double py = (double)a[1] + mix_b * (double)e1[1] + mix_c * (double)e2[1];
        double pz = (double)a[2] + mix_b * (double)e1[2] + mix_c * (double)e2[2];
        double r1 = orc_mod1((double)(float)(h_r1[idx] + cp_dims[3]));      /* :75 */
        double r2 = orc_mod1((double)(float)(h_r2[idx] + cp_dims[4]));      /* :76 */
        double sin_t = sqrt(1.0 - r1);                                      /* :78 */
        double phi = two_pi * r2;
        double x = sin_t * cos(phi);
        double y = sin_t * sin(phi);
        double z = sqrt(r1);
        double dx = x * (double)tu[0] + y * (double)tv[0] + z * (double)tn[0];     /* :84-86 */
        double dy = x * (double)tu[1] + y * (double)tv[1] + z * (double)tn[1];
        double dz = x * (double)tu[2] + y * (double)tv[2] + z * (double)tn[2];
        float eps = tri_eps[tri];                                           /* :87 */
        orig[3 * idx + 0] = (float)(px + (double)(float)(eps * tn[0]));     /* :89-91 eps*n is f32*f32 */
        orig[3 * idx + 1] = (float)(py + (double)(float)(eps * tn[1]));
        orig[3 * idx + 2] = (float)(pz + (double)(float)(eps * tn[2]));
        dire[3 * idx + 0] = (float)dx;
        dire[3 * idx + 1] = (float)dy;
        dire[3 * idx + 2] = (float)dz;
    }
}

/* ------------------------------------------------------------------ tracing */

/* utils/cpu_trace.py:12-42 `_aabb_tmin`: (b - o) is float32, times a float64 inverse. */
static inline double orc_aabb_tmin(float o0, float o1, float o2, double inv0, double inv1, double inv2,
                                   const float *bmin, const float *bmax) {
    double tmin = (double)(float)(bmin[0] - o0) * inv0;
    double tmax = (double)(float)(bmax[0] - o0) * inv0;
    if (tmin > tmax) { double s = tmin; tmin = tmax; tmax = s; }
    double tymin = (double)(float)(bmin[1] - o1) * inv1;
    double tymax = (double)(float)(bmax[1] - o1) * inv1;
    if (tymin > tymax) { double s = tymin; tymin = tymax; tymax = s; }
    if (tmin > tymax || tymin > tmax) return ORC_INF;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    double tzmin = (double)(float)(bmin[2] - o2) * inv2;
    double tzmax = (double)(float)(bmax[2] - o2) * inv2;
    if (tzmin > tzmax) { double s = tzmin; tzmin = tzmax; tzmax = s; }
    if (tmin > tzmax || tzmin > tmax) return ORC_INF;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    if (tmax < 0.0) return ORC_INF;
    return tmin > 0.0 ? tmin : 0.0;
}

/* utils/cpu_trace.py:45-51 `_skip_surface`. */
static inline int orc_skip(int32_t s, const uint8_t *surf_active, int32_t emit_sid, int32_t min_sid) {
    if (surf_active[s] == 0) return 1;
    if (s < min_sid) return 1;
    return s == emit_sid;
}

/* Moeller-Trumbore exactly as utils/cpu_trace.py:88-110 (and every twin of that block):
 * cross/dot products in float32, inv_det/u/v/t in float64.  Returns 1 and *t_out when the
 * triangle is hit at any t (the caller applies its own t window). */
static inline int orc_tri(float o0, float o1, float o2, float d0, float d1, float d2,
                          const float *v0, const float *e1, const float *e2, double *t_out) {
    float px = d1 * e2[2] - d2 * e2[1];
    float py = d2 * e2[0] - d0 * e2[2];
    float pz = d0 * e2[1] - d1 * e2[0];
    float det = e1[0] * px + e1[1] * py + e1[2] * pz;
    if (fabs((double)det) < 1e-7) return 0;
    double inv_det = 1.0 / (double)det;
    float tx = o0 - v0[0], ty = o1 - v0[1], tz = o2 - v0[2];
    double u = (double)(float)(tx * px + ty * py + tz * pz) * inv_det;
    if (u < 0.0 || u > 1.0) return 0;
    float qx = ty * e1[2] - tz * e1[1];
    float qy = tz * e1[0] - tx * e1[2];
    float qz = tx * e1[1] - ty * e1[0];
    double v = (double)(float)(d0 * qx + d1 * qy + d2 * qz) * inv_det;
    if (v < 0.0 || u + v > 1.0) return 0;
    *t_out = (double)(float)(e2[0] * qx + e2[1] * qy + e2[2] * qz) * inv_det;
    return 1;
}

static inline uint8_t orc_front(float d0, float d1, float d2, const float *n) {
    return (-(d0 * n[0] + d1 * n[1] + d2 * n[2]) > 0.0f) ? 1 : 0;   /* cpu_trace.py:114 */
}

/* utils/cpu_trace.py:54-117 `trace_cpu_firsthit` (brute force, triangle order, strict t<best). */
void orc_trace_firsthit(int64_t n_rays, const float *orig, const float *dirs,
                        int64_t n_tri, const float *v0, const float *e1, const float *e2,
                        const float *norm, const int32_t *sid, const uint8_t *surf_active,
                        int32_t emit_sid, int32_t min_sid, int32_t *out_hit_sid, uint8_t *out_front) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n_rays; ++k) {
        float o0 = orig[3 * k], o1 = orig[3 * k + 1], o2 = orig[3 * k + 2];
        float d0 = dirs[3 * k], d1 = dirs[3 * k + 1], d2 = dirs[3 * k + 2];
        double best = ORC_INF; int32_t hit = -1; uint8_t front = 0;
        for (int64_t i = 0; i < n_tri; ++i) {
            int32_t s = sid[i];
            if (orc_skip(s, surf_active, emit_sid, min_sid)) continue;
            double t;
            if (!orc_tri(o0, o1, o2, d0, d1, d2, v0 + 3 * i, e1 + 3 * i, e2 + 3 * i, &t)) continue;
            if (1e-6 < t && t < best) { best = t; hit = s; front = orc_front(d0, d1, d2, norm + 3 * i); }
        }
        out_hit_sid[k] = hit;
        out_front[k] = hit >= 0 ? front : 0;
    }
}

/* utils/cpu_trace.py:120-277 `trace_cpu_bvh_firsthit`: binary BVH, near child popped first,
 * float32 tstack, `node_t >= best` culling.  Optional per-ray work counters (may be NULL):
 * stats[0..3] += inner visits, leaf visits, triangles tested, triangles skipped. */
void orc_trace_bvh_firsthit(int64_t n_rays, const float *orig, const float *dirs,
                            const float *v0, const float *e1, const float *e2,
                            const float *norm, const int32_t *sid, const uint8_t *surf_active,
                            const float *bb_min, const float *bb_max,
                            const int32_t *left, const int32_t *right,
                            const int32_t *start, const int32_t *count,
                            int32_t emit_sid, int32_t min_sid,
                            int32_t *out_hit_sid, uint8_t *out_front, int64_t *stats) {
    int64_t s_inner = 0, s_leaf = 0, s_tri = 0, s_skip = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : s_inner, s_leaf, s_tri, s_skip)
    for (int64_t k = 0; k < n_rays; ++k) {
        float o0 = orig[3 * k], o1 = orig[3 * k + 1], o2 = orig[3 * k + 2];
        float d0 = dirs[3 * k], d1 = dirs[3 * k + 1], d2 = dirs[3 * k + 2];
        double inv0 = fabs((double)d0) > 1e-9 ? 1.0 / (double)d0 : 1e10;   /* :150-152 */
        double inv1 = fabs((double)d1) > 1e-9 ? 1.0 / (double)d1 : 1e10;
        double inv2 = fabs((double)d2) > 1e-9 ? 1.0 / (double)d2 : 1e10;
        double root_t = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min, bb_max);
        if (root_t >= ORC_INF) { out_hit_sid[k] = -1; out_front[k] = 0; continue; }
        int32_t stack[ORC_STACK]; float tstack[ORC_STACK]; int sp = 0;
        stack[sp] = 0; tstack[sp] = (float)root_t; sp++;
        double best = ORC_INF; int32_t hit = -1; uint8_t front = 0;
        while (sp > 0) {
            sp--;
            int32_t node = stack[sp];
            float node_t = tstack[sp];
            if ((double)node_t >= best) continue;
            if (count[node] > 0) {
                s_leaf++;
                for (int32_t t = 0; t < count[node]; ++t) {
                    int64_t tri = (int64_t)start[node] + t;
                    int32_t s = sid[tri];
                    if (orc_skip(s, surf_active, emit_sid, min_sid)) { s_skip++; continue; }
                    s_tri++;
                    double tp;
                    if (!orc_tri(o0, o1, o2, d0, d1, d2, v0 + 3 * tri, e1 + 3 * tri, e2 + 3 * tri, &tp)) continue;
                    if (1e-6 < tp && tp < best) { best = tp; hit = s; front = orc_front(d0, d1, d2, norm + 3 * tri); }
                }
            } else {
                s_inner++;
                int32_t ln = left[node], rn = right[node];
                double tl = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min + 3 * ln, bb_max + 3 * ln);
                double tr = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min + 3 * rn, bb_max + 3 * rn);
                if (tl < tr) {                                               /* :258-274 */
                    if (tr < best && sp < ORC_STACK) { stack[sp] = rn; tstack[sp] = (float)tr; sp++; }
                    if (tl < best && sp < ORC_STACK) { stack[sp] = ln; tstack[sp] = (float)tl; sp++; }
                } else {
                    if (tl < best && sp < ORC_STACK) { stack[sp] = ln; tstack[sp] = (float)tl; sp++; }
                    if (tr < best && sp < ORC_STACK) { stack[sp] = rn; tstack[sp] = (float)tr; sp++; }
                }
            }
        }
        out_hit_sid[k] = hit;
        out_front[k] = hit >= 0 ? front : 0;
    }
    if (stats) { stats[0] += s_inner; stats[1] += s_leaf; stats[2] += s_tri; stats[3] += s_skip; }
}

/* utils/cpu_trace.py:540-583 `trace_cpu_hitmask` (any hit with t > 1e-6, first in triangle order). */
void orc_trace_hitmask(int64_t n_rays, const float *orig, const float *dirs,
                       int64_t n_tri, const float *v0, const float *e1, const float *e2,
                       const int32_t *sid, const uint8_t *surf_active,
                       int32_t emit_sid, int32_t min_sid, uint8_t *out_hitmask) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n_rays; ++k) {
        float o0 = orig[3 * k], o1 = orig[3 * k + 1], o2 = orig[3 * k + 2];
        float d0 = dirs[3 * k], d1 = dirs[3 * k + 1], d2 = dirs[3 * k + 2];
        uint8_t any = 0;
        for (int64_t i = 0; i < n_tri; ++i) {
            if (orc_skip(sid[i], surf_active, emit_sid, min_sid)) continue;
            double t;
            if (!orc_tri(o0, o1, o2, d0, d1, d2, v0 + 3 * i, e1 + 3 * i, e2 + 3 * i, &t)) continue;
            if (t > 1e-6) { any = 1; break; }
        }
        out_hitmask[k] = any;
    }
}

/* utils/cpu_trace.py:586-732 `trace_cpu_bvh_hitmask`: no distance culling, children pushed when t < INF. */
void orc_trace_bvh_hitmask(int64_t n_rays, const float *orig, const float *dirs,
                           const float *v0, const float *e1, const float *e2,
                           const int32_t *sid, const uint8_t *surf_active,
                           const float *bb_min, const float *bb_max,
                           const int32_t *left, const int32_t *right,
                           const int32_t *start, const int32_t *count,
                           int32_t emit_sid, int32_t min_sid, uint8_t *out_hitmask) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n_rays; ++k) {
        float o0 = orig[3 * k], o1 = orig[3 * k + 1], o2 = orig[3 * k + 2];
        float d0 = dirs[3 * k], d1 = dirs[3 * k + 1], d2 = dirs[3 * k + 2];
        double inv0 = fabs((double)d0) > 1e-9 ? 1.0 / (double)d0 : 1e10;
        double inv1 = fabs((double)d1) > 1e-9 ? 1.0 / (double)d1 : 1e10;
        double inv2 = fabs((double)d2) > 1e-9 ? 1.0 / (double)d2 : 1e10;
        double root_t = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min, bb_max);
        if (root_t >= ORC_INF) { out_hitmask[k] = 0; continue; }
        int32_t stack[ORC_STACK]; int sp = 0;
        stack[sp++] = 0;
        uint8_t any = 0;
        while (sp > 0 && !any) {
            int32_t node = stack[--sp];
            if (count[node] > 0) {
                for (int32_t t = 0; t < count[node]; ++t) {
                    int64_t tri = (int64_t)start[node] + t;
                    if (orc_skip(sid[tri], surf_active, emit_sid, min_sid)) continue;
                    double tp;
                    if (!orc_tri(o0, o1, o2, d0, d1, d2, v0 + 3 * tri, e1 + 3 * tri, e2 + 3 * tri, &tp)) continue;
                    if (tp > 1e-6) { any = 1; break; }
                }
            } else {
                int32_t ln = left[node], rn = right[node];
                double tl = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min + 3 * ln, bb_max + 3 * ln);
                double tr = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min + 3 * rn, bb_max + 3 * rn);
                if (tl < tr) {
                    if (tr < ORC_INF && sp < ORC_STACK) stack[sp++] = rn;
                    if (tl < ORC_INF && sp < ORC_STACK) stack[sp++] = ln;
                } else {
                    if (tl < ORC_INF && sp < ORC_STACK) stack[sp++] = ln;
                    if (tr < ORC_INF && sp < ORC_STACK) stack[sp++] = rn;
                }
            }
        }
        out_hitmask[k] = any;
    }
}

/* utils/cpu_trace.py:280-352 `trace_cpu_combined`: closest matrix hit (surf >= matrix_min_sid) plus
 * an any-hit flag over every active non-emitter surface, in one brute-force pass. */
void orc_trace_combined(int64_t n_rays, const float *orig, const float *dirs,
                        int64_t n_tri, const float *v0, const float *e1, const float *e2,
                        const float *norm, const int32_t *sid, const uint8_t *surf_active,
                        int32_t emit_sid, int32_t matrix_min_sid,
                        int32_t *out_hit_sid, uint8_t *out_front, uint8_t *out_any) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n_rays; ++k) {
        float o0 = orig[3 * k], o1 = orig[3 * k + 1], o2 = orig[3 * k + 2];
        float d0 = dirs[3 * k], d1 = dirs[3 * k + 1], d2 = dirs[3 * k + 2];
        double best = ORC_INF; int32_t hit = -1; uint8_t front = 0, any = 0;
        for (int64_t i = 0; i < n_tri; ++i) {
            int32_t s = sid[i];
            if (s == emit_sid || surf_active[s] == 0) continue;
            double t;
            if (!orc_tri(o0, o1, o2, d0, d1, d2, v0 + 3 * i, e1 + 3 * i, e2 + 3 * i, &t)) continue;
            if (t <= 1e-6) continue;
            any = 1;
            if (s < matrix_min_sid) continue;
            if (t < best) { best = t; hit = s; front = orc_front(d0, d1, d2, norm + 3 * i); }
        }
        out_hit_sid[k] = hit;
        out_front[k] = hit >= 0 ? front : 0;
        out_any[k] = any;
    }
}

/* utils/cpu_trace.py:355-522 `trace_cpu_bvh_combined`. */
void orc_trace_bvh_combined(int64_t n_rays, const float *orig, const float *dirs,
                            const float *v0, const float *e1, const float *e2,
                            const float *norm, const int32_t *sid, const uint8_t *surf_active,
                            const float *bb_min, const float *bb_max,
                            const int32_t *left, const int32_t *right,
                            const int32_t *start, const int32_t *count,
                            int32_t emit_sid, int32_t matrix_min_sid,
                            int32_t *out_hit_sid, uint8_t *out_front, uint8_t *out_any) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n_rays; ++k) {
        float o0 = orig[3 * k], o1 = orig[3 * k + 1], o2 = orig[3 * k + 2];
        float d0 = dirs[3 * k], d1 = dirs[3 * k + 1], d2 = dirs[3 * k + 2];
        double inv0 = fabs((double)d0) > 1e-9 ? 1.0 / (double)d0 : 1e10;
        double inv1 = fabs((double)d1) > 1e-9 ? 1.0 / (double)d1 : 1e10;
        double inv2 = fabs((double)d2) > 1e-9 ? 1.0 / (double)d2 : 1e10;
        double root_t = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min, bb_max);
        if (root_t >= ORC_INF) { out_hit_sid[k] = -1; out_front[k] = 0; out_any[k] = 0; continue; }
        int32_t stack[ORC_STACK]; float tstack[ORC_STACK]; int sp = 0;
        stack[sp] = 0; tstack[sp] = (float)root_t; sp++;
        double best = ORC_INF; int32_t hit = -1; uint8_t front = 0, any = 0;
        while (sp > 0) {
            sp--;
            int32_t node = stack[sp];
            float node_t = tstack[sp];
            if ((double)node_t >= best) continue;
            if (count[node] > 0) {
                for (int32_t t = 0; t < count[node]; ++t) {
                    int64_t tri = (int64_t)start[node] + t;
                    int32_t s = sid[tri];
                    if (s == emit_sid || surf_active[s] == 0) continue;
                    double tp;
                    if (!orc_tri(o0, o1, o2, d0, d1, d2, v0 + 3 * tri, e1 + 3 * tri, e2 + 3 * tri, &tp)) continue;
                    if (tp <= 1e-6) continue;
                    any = 1;
                    if (s < matrix_min_sid) continue;
                    if (tp < best) { best = tp; hit = s; front = orc_front(d0, d1, d2, norm + 3 * tri); }
                }
            } else {
                int32_t ln = left[node], rn = right[node];
                double tl = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min + 3 * ln, bb_max + 3 * ln);
                double tr = orc_aabb_tmin(o0, o1, o2, inv0, inv1, inv2, bb_min + 3 * rn, bb_max + 3 * rn);
                if (tl < tr) {
                    if (tr < best && sp < ORC_STACK) { stack[sp] = rn; tstack[sp] = (float)tr; sp++; }
                    if (tl < best && sp < ORC_STACK) { stack[sp] = ln; tstack[sp] = (float)tl; sp++; }
                } else {
                    if (tl < best && sp < ORC_STACK) { stack[sp] = ln; tstack[sp] = (float)tl; sp++; }
                    if (tr < best && sp < ORC_STACK) { stack[sp] = rn; tstack[sp] = (float)tr; sp++; }
                }
            }
        }
        out_hit_sid[k] = hit;
        out_front[k] = hit >= 0 ? front : 0;
        out_any[k] = any;
    }
}

/* ------------------------------------------------------------------ tallies */

/* utils/cpu_trace.py:525-537 `reduce_first_hits`. */
void orc_reduce_first_hits(int64_t n_rays, const int32_t *hit_sid, const uint8_t *front_flag,
                           int64_t n_surf, int64_t *out_front, int64_t *out_back) {
    for (int64_t i = 0; i < n_surf; ++i) { out_front[i] = 0; out_back[i] = 0; }
    for (int64_t i = 0; i < n_rays; ++i) {
        int32_t h = hit_sid[i];
        if (h < 0) continue;
        if (front_flag[i]) out_front[h]++; else out_back[h]++;
    }
}

/* utils/cpu_trace.py:735-777 `_tregenza_patch_id` with float32 arguments: the ring test compares
 * float32 dz with float64 sines; atan2 and degrees are evaluated in float32 (Numba picks the
 * float32 overloads), everything after `az += 360.0` is float64. */
int32_t orc_tregenza_patch_id(float dx, float dy, float dz) {
    static const double ring_hi_sin[8] = {0.20791169081775934, 0.40673664307580015, 0.5877852522924731,
                                          0.7431448254773942,  0.8660254037844386,  0.9510565162951535,
                                          0.9945218953682733,  1.0};
    static const int ring_n[8] = {30, 30, 24, 24, 18, 12, 6, 1};
    static const int ring_start[8] = {0, 30, 60, 84, 108, 126, 138, 144};
    if ((double)dz <= 0.0) return -1;
    int ridx = 7;
    for (int j = 0; j < 8; ++j) {
        if ((double)dz < ring_hi_sin[j] || j == 7) { ridx = j; break; }
    }
    int n_az = ring_n[ridx], base = ring_start[ridx];
    if (n_az == 1) return base;
    float azf = atan2f(dy, dx) * (float)(180.0 / 3.141592653589793);   /* math.degrees on float32 */
    double az = (double)azf;
    if (az < 0.0) az += 360.0;
    double width = 360.0 / (double)n_az;
    double off = (ridx & 1) ? (180.0 / (double)n_az) : 0.0;
    double t = az - off;
    if (t < 0.0) t += 360.0; else if (t >= 360.0) t -= 360.0;
    int aidx = (int)floor(t / width);
    if (aidx >= n_az) aidx = n_az - 1;
    return base + aidx;
}

/* utils/cpu_trace.py:780-789 `bin_tregenza_cpu`. */
void orc_bin_tregenza(int64_t n_rays, const float *dirs, const uint8_t *hitmask, int64_t *counts) {
    for (int i = 0; i < 145; ++i) counts[i] = 0;
    for (int64_t i = 0; i < n_rays; ++i) {
        if (hitmask[i]) continue;
        int32_t pid = orc_tregenza_patch_id(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        if (pid >= 0) counts[pid]++;
    }
}

/* utils/cpu_trace.py:792-798 `count_upward_misses_cpu`. */
int64_t orc_count_upward_misses(int64_t n_rays, const float *dirs, const uint8_t *hitmask) {
    int64_t total = 0;
    for (int64_t i = 0; i < n_rays; ++i)
        if (hitmask[i] == 0 && dirs[3 * i + 2] > 0.0f) total++;
    return total;
}
