// Device-side preparation: raw meshes -> scene triangles and emitter records, on the GPU.
//
// Replaces, for the device path, the per-triangle Python loops of the reference's prepare_scene / prepare_emitters
// (utils/prepared.py:93-321: 115 s for a million triangles there).  Only vertices (float32) and faces (int32) cross
// the PCIe bus (19 MB for the million-triangle scene instead of 139 MB of prepared arrays).
//
// Every arithmetic step reproduces what NumPy does in the reference, bit for bit: float32 multiplies, adds, divides
// and square roots with round-to-nearest and no fused multiply-add (np.cross, np.linalg.norm(axis=1), `v / n`),
// NumPy's pairwise float32 summation for `areas.sum()`, a sequential float64 cumsum for the area CDF.  The only part
// that is not reproduced is the BLAS matrix-vector product inside `_emitter_plane` (prepared.py:155-166); for it the
// kernel returns float64 statistics and the host decides, falling back to the reference arithmetic for a mesh whose
// statistics lie within rounding distance of a threshold (raystrack_b200/prepared.py).
#include "rsk_common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

struct rsk_geometry {
    rsk_ctx *ctx = nullptr;
    int32_t n_mesh = 0;
    int64_t n_vert = 0, n_tri = 0;
    float *verts = nullptr;          // [n_vert][3]
    int32_t *faces = nullptr;        // [n_tri][3], indices local to the mesh
    int64_t *vert_off = nullptr;     // [n_mesh+1]
    int64_t *tri_off = nullptr;      // [n_mesh+1]
    std::vector<int64_t> h_tri_off;
};

typedef rsk_mesh_summary MeshSummary;      // per-mesh by-products (include/raystrack_b200.h)

namespace {

struct F3 { float x, y, z; };

__device__ __forceinline__ F3 sub3(F3 a, F3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ F3 add3(F3 a, F3 b) { return {__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)}; }
// np.cross: c0 = a1*b2 - a2*b1, c1 = a2*b0 - a0*b2, c2 = a0*b1 - a1*b0, every product rounded before the subtraction
__device__ __forceinline__ F3 cross3(F3 a, F3 b) {
    return {__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
            __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}
// np.linalg.norm(v, axis=1): sqrt(((x*x + y*y) + z*z)) in float32
__device__ __forceinline__ float norm3(F3 v) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
}
__device__ __forceinline__ F3 div3(F3 v, float d) { return {__fdiv_rn(v.x, d), __fdiv_rn(v.y, d), __fdiv_rn(v.z, d)}; }

__device__ __forceinline__ int mesh_of(const int64_t *tri_off, int n_mesh, int64_t t) {
    int lo = 0, hi = n_mesh - 1;                 // last m with tri_off[m] <= t
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tri_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// One thread per triangle.  scene_tri/scene_nrm (either both or none) receive the traversal records of the scene
// (prepared.py:170-243, faces as given); em_rec/em_area the emitter records (prepared.py:246-321, faces flipped to
// [0,2,1] when `flip`).  bad[0] counts face indices outside the mesh's vertex range.
__global__ void rsk_prepare_triangles_kernel(const float *__restrict__ verts, const int32_t *__restrict__ faces,
                                             const int64_t *__restrict__ vert_off, const int64_t *__restrict__ tri_off,
                                             int n_mesh, int64_t n_tri, int flip, float4 *scene_tri, float4 *scene_nrm,
                                             float4 *em_rec, float *em_area, int *bad) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tri) return;
    const int m = mesh_of(tri_off, n_mesh, t);
    const int64_t v0 = vert_off[m], nv = vert_off[m + 1] - v0;
    int i0 = faces[3 * t], i1 = faces[3 * t + 1], i2 = faces[3 * t + 2];
    if (flip) { const int s = i1; i1 = i2; i2 = s; }
    // NumPy index semantics: negative indices count from the end; anything else out of range is an IndexError
    if (i0 < 0) i0 += (int)nv;
    if (i1 < 0) i1 += (int)nv;
    if (i2 < 0) i2 += (int)nv;
    if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= nv || i1 >= nv || i2 >= nv) { atomicAdd(bad, 1); return; }
    const float *p0 = verts + 3 * (v0 + i0), *p1 = verts + 3 * (v0 + i1), *p2 = verts + 3 * (v0 + i2);
    const F3 a = {p0[0], p0[1], p0[2]};
    const F3 e1 = sub3({p1[0], p1[1], p1[2]}, a);
    const F3 e2 = sub3({p2[0], p2[1], p2[2]}, a);
    const F3 nraw = cross3(e1, e2);
    const float twice = norm3(nraw);
    const F3 n = div3(nraw, fmaxf(twice, 1e-12f));                       // _unit_rows (prepared.py:93-96)
    const float sbits = __int_as_float(m);
    if (scene_tri) {
        scene_tri[3 * t] = make_float4(a.x, a.y, a.z, sbits);
        scene_tri[3 * t + 1] = make_float4(e1.x, e1.y, e1.z, 0.f);
        scene_tri[3 * t + 2] = make_float4(e2.x, e2.y, e2.z, 0.f);
        scene_nrm[t] = make_float4(n.x, n.y, n.z, sbits);
    }
    if (!em_rec) return;
    // tangent frame (prepared.py:99-122): u = normalise(ref x n), ref = x unless |n.x| >= 0.9; v = n x u
    const F3 ax = {1.f, 0.f, 0.f}, ay = {0.f, 1.f, 0.f};
    const bool use_x = fabs((double)n.x) < 0.9;
    F3 u = cross3(use_x ? ax : ay, n);
    float len = norm3(u);
    if ((double)len <= 1e-12) {
        u = cross3(use_x ? ay : ax, n);
        len = norm3(u);
    }
    F3 v;
    if ((double)len <= 1e-12) { u = ax; v = ay; }
    else { u = div3(u, len); v = cross3(n, u); }
    // ray-origin offset (prepared.py:125-130): 1e-6 of the longest edge, at least 1e-8
    const float scale = fmaxf(norm3(e1), fmaxf(norm3(e2), norm3(sub3(e2, e1))));
    const float eps = fmaxf(__fmul_rn(scale, 1.0e-6f), 1.0e-8f);
    em_rec[5 * t + 0] = make_float4(a.x, a.y, a.z, eps);
    em_rec[5 * t + 1] = make_float4(e1.x, e1.y, e1.z, n.x);
    em_rec[5 * t + 2] = make_float4(e2.x, e2.y, e2.z, n.y);
    em_rec[5 * t + 3] = make_float4(u.x, u.y, u.z, n.z);
    em_rec[5 * t + 4] = make_float4(v.x, v.y, v.z, 0.f);
    em_area[t] = __fmul_rn(0.5f, twice);
}

// NumPy's pairwise summation of n contiguous float32 values (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum):
// fewer than 8 values are added left to right; up to 128 values go through eight strided partial sums combined as
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)) plus the tail; longer runs split at n/2 rounded down to a multiple of 8.
__device__ float np_block_sum(const float *a, int64_t n) {
    if (n < 8) {
        float r = 0.f;
        for (int64_t i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

__device__ float np_pairwise_sum(const float *a, int64_t n) {
    // explicit stack instead of recursion: (offset, length, state) with partial results kept per level
    struct Frame { int64_t off, len; float left; int state; };
    Frame st[48];
    int sp = 0;
    st[0] = {0, n, 0.f, 0};
    float ret = 0.f;
    while (sp >= 0) {
        Frame &f = st[sp];
        if (f.len <= 128) { ret = np_block_sum(a + f.off, f.len); --sp; continue; }
        int64_t n2 = f.len / 2;
        n2 -= n2 % 8;
        if (f.state == 0) { f.state = 1; st[sp + 1] = {f.off, n2, 0.f, 0}; ++sp; }
        else if (f.state == 1) { f.left = ret; f.state = 2; st[sp + 1] = {f.off + n2, f.len - n2, 0.f, 0}; ++sp; }
        else { ret = __fadd_rn(f.left, ret); --sp; }
    }
    return ret;
}

// One warp per mesh: area total (NumPy pairwise order, lane 0), float64 running sum of the areas in triangle order
// (np.cumsum(areas, dtype=float64): the warp loads 32 areas at a time and every lane replays the 32 additions), the
// CDF cum/cum[-1] rounded to float32, and the planarity statistics of prepared.py:133-167.
__global__ void rsk_prepare_meshes_kernel(const float4 *__restrict__ em_rec, const float *__restrict__ em_area,
                                          const int64_t *__restrict__ tri_off, int n_mesh, double *cum, float *cdf,
                                          MeshSummary *summary) {
    const int m = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= n_mesh) return;
    const int64_t t0 = tri_off[m], nt = tri_off[m + 1] - t0;
    MeshSummary s;
    s.total_area = 0.0;
    s.origin[0] = s.origin[1] = s.origin[2] = 0.f;
    s.normal0[0] = s.normal0[1] = s.normal0[2] = 0.f;
    s.eps_max = 0.f; s.reserved = 0.f;
    s.min_dot = 1.0; s.worst = 0.0; s.worst_mag = 0.0;
    if (nt <= 0) { if (lane == 0) summary[m] = s; return; }

    float total = 0.f;
    if (lane == 0) total = np_pairwise_sum(em_area + t0, nt);
    total = __shfl_sync(0xffffffffu, total, 0);

    // running float64 sum in triangle order
    double acc = 0.0;
    for (int64_t base = 0; base < nt; base += 32) {
        const int64_t i = base + lane;
        const float mine = i < nt ? em_area[t0 + i] : 0.f;
        double my_cum = 0.0;
        const int cnt = (int)min((int64_t)32, nt - base);
        for (int j = 0; j < cnt; ++j) {
            acc = __dadd_rn(acc, (double)__shfl_sync(0xffffffffu, mine, j));
            if (j == lane) my_cum = acc;
        }
        if (i < nt) cum[t0 + i] = my_cum;
    }
    const double last = acc;                                   // cum[-1]
    const bool flat = (double)total <= 0.0;                     // cdf = ones (prepared.py:305-307)
    // planarity statistics against the first triangle
    const float4 r0 = em_rec[5 * t0], r1 = em_rec[5 * t0 + 1], r2 = em_rec[5 * t0 + 2], r3 = em_rec[5 * t0 + 3];
    const F3 org = {r0.x, r0.y, r0.z};
    const double nx = r1.w, ny = r2.w, nz = r3.w;
    double min_dot = 1.0e300, worst = 0.0, mag = 0.0;
    float eps_max = 0.f;
    for (int64_t i = lane; i < nt; i += 32) {
        const float4 q0 = em_rec[5 * (t0 + i)], q1 = em_rec[5 * (t0 + i) + 1], q2 = em_rec[5 * (t0 + i) + 2], q3 = em_rec[5 * (t0 + i) + 3];
        cdf[t0 + i] = flat ? 1.0f : (float)__ddiv_rn(cum[t0 + i], last);
        eps_max = fmaxf(eps_max, q0.w);
        min_dot = fmin(min_dot, (double)q1.w * nx + (double)q2.w * ny + (double)q3.w * nz);
        const F3 a = {q0.x, q0.y, q0.z};
        const F3 corner[3] = {a, add3(a, {q1.x, q1.y, q1.z}), add3(a, {q2.x, q2.y, q2.z})};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const F3 d = sub3(corner[c], org);
            const double tx = (double)d.x * nx, ty = (double)d.y * ny, tz = (double)d.z * nz;
            worst = fmax(worst, fabs(tx + ty + tz));
            mag = fmax(mag, fabs(tx) + fabs(ty) + fabs(tz));
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        eps_max = fmaxf(eps_max, __shfl_xor_sync(0xffffffffu, eps_max, o));
        min_dot = fmin(min_dot, __shfl_xor_sync(0xffffffffu, min_dot, o));
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        mag = fmax(mag, __shfl_xor_sync(0xffffffffu, mag, o));
    }
    if (lane == 0) {
        s.total_area = (double)total;
        s.origin[0] = org.x; s.origin[1] = org.y; s.origin[2] = org.z;
        s.normal0[0] = r1.w; s.normal0[1] = r2.w; s.normal0[2] = r3.w;
        s.eps_max = eps_max;
        s.min_dot = min_dot; s.worst = worst; s.worst_mag = mag;
        summary[m] = s;
    }
}

// One thread per (emitter, surface): `surf_active` of reference main.py:167-204.  A planar emitter switches off every
// mesh whose bounding box lies wholly behind its plane; float32 operations in the reference's order
// (dx*nx + dy*ny + dz*nz, then |n|.extent), no fused multiply-add.  The emitter's own mesh is always off.
__global__ void rsk_surface_masks_kernel(int n_emit, int n_surf, const uint8_t *__restrict__ planar, const float *__restrict__ po,
                                         const float *__restrict__ pn, const float *__restrict__ tol,
                                         const float *__restrict__ centers, const float *__restrict__ extents, uint8_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_emit * n_surf) return;
    const int e = (int)(i / n_surf), s = (int)(i - (int64_t)e * n_surf);
    uint8_t on = 1;
    if (planar[e]) {
        const float nx = pn[3 * e], ny = pn[3 * e + 1], nz = pn[3 * e + 2];
        float signed_d = __fmul_rn(__fsub_rn(centers[3 * s], po[3 * e]), nx);
        signed_d = __fadd_rn(signed_d, __fmul_rn(__fsub_rn(centers[3 * s + 1], po[3 * e + 1]), ny));
        signed_d = __fadd_rn(signed_d, __fmul_rn(__fsub_rn(centers[3 * s + 2], po[3 * e + 2]), nz));
        float radius = __fmul_rn(fabsf(nx), extents[3 * s]);
        radius = __fadd_rn(radius, __fmul_rn(fabsf(ny), extents[3 * s + 1]));
        radius = __fadd_rn(radius, __fmul_rn(fabsf(nz), extents[3 * s + 2]));
        signed_d = __fadd_rn(signed_d, radius);
        on = (signed_d <= tol[e]) ? 0 : 1;
    }
    if (e == s) on = 0;
    out[i] = on;
}

}  // namespace

// ----------------------------------------------------------------------------- geometry handle

extern "C" int rsk_geometry_create(rsk_ctx *ctx, int32_t n_mesh, const float *verts, const int64_t *vert_offset,
                                   const int32_t *faces, const int64_t *tri_offset, rsk_geometry **out) {
    RSK_REQUIRE(ctx && out && n_mesh >= 0, "rsk_geometry_create: bad arguments");
    RSK_REQUIRE(n_mesh == 0 || (vert_offset && tri_offset), "rsk_geometry_create: null offsets");
    *out = nullptr;
    const int64_t nv = n_mesh ? vert_offset[n_mesh] : 0, nt = n_mesh ? tri_offset[n_mesh] : 0;
    RSK_REQUIRE(nv >= 0 && nt >= 0 && nt < (1ll << 30) && nv < (1ll << 31), "rsk_geometry_create: sizes out of range");
    RSK_REQUIRE((nv == 0 || verts) && (nt == 0 || faces), "rsk_geometry_create: null arrays");
    for (int i = 0; i < n_mesh; ++i)
        RSK_REQUIRE(vert_offset[i] <= vert_offset[i + 1] && tri_offset[i] <= tri_offset[i + 1], "rsk_geometry_create: offsets must not decrease");
    RskScope scope(ctx);
    rsk_geometry *g = new rsk_geometry();
    g->ctx = ctx; g->n_mesh = n_mesh; g->n_vert = nv; g->n_tri = nt;
    g->h_tri_off.assign(tri_offset, tri_offset + (n_mesh ? n_mesh + 1 : 0));
    int rc = rsk_dev_alloc(&g->verts, (size_t)nv * 3);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&g->faces, (size_t)nt * 3);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&g->vert_off, (size_t)n_mesh + 1);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&g->tri_off, (size_t)n_mesh + 1);
    if (rc == RSK_OK) {
        cudaError_t e = cudaSuccess;
        if (nv) e = cudaMemcpyAsync(g->verts, verts, (size_t)nv * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && nt) e = cudaMemcpyAsync(g->faces, faces, (size_t)nt * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && n_mesh) e = cudaMemcpyAsync(g->vert_off, vert_offset, ((size_t)n_mesh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && n_mesh) e = cudaMemcpyAsync(g->tri_off, tri_offset, ((size_t)n_mesh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);          // the caller's arrays may go away after return
        if (e != cudaSuccess) { rsk_set_error("geometry upload failed: %s", cudaGetErrorString(e)); rc = RSK_ERR_CUDA; }
    }
    if (rc != RSK_OK) { rsk_geometry_destroy(g); return rc; }
    *out = g;
    return RSK_OK;
}

extern "C" int rsk_geometry_destroy(rsk_geometry *g) {
    if (!g) return RSK_OK;
    RskScope scope(g->ctx);
    rsk_dev_free(g->verts); rsk_dev_free(g->faces); rsk_dev_free(g->vert_off); rsk_dev_free(g->tri_off);
    delete g;
    return RSK_OK;
}

static int rsk_check_faces(rsk_ctx *ctx, int *d_bad, const char *who) {
    int bad = 0;
    RSK_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad) { rsk_set_error("%s: %d face indices are outside their mesh's vertex range", who, bad); return RSK_ERR_INVALID; }
    return RSK_OK;
}

extern "C" int rsk_scene_from_geometry(rsk_geometry *g, int32_t use_bvh, rsk_scene **out) {
    RSK_REQUIRE(g && out, "rsk_scene_from_geometry: null argument");
    *out = nullptr;
    rsk_ctx *ctx = g->ctx;
    RskScope scope(ctx);
    float4 *d_tri = nullptr, *d_nrm = nullptr;
    int *d_bad = nullptr;
    int rc = rsk_dev_alloc(&d_tri, (size_t)g->n_tri * 3);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_nrm, (size_t)g->n_tri);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_bad, 1);
    if (rc == RSK_OK && cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream) != cudaSuccess) rc = RSK_ERR_CUDA;
    if (rc == RSK_OK && g->n_tri > 0) {
        rsk_prepare_triangles_kernel<<<rsk_blocks(g->n_tri, 256), 256, 0, ctx->stream>>>(
            g->verts, g->faces, g->vert_off, g->tri_off, g->n_mesh, g->n_tri, 0, d_tri, d_nrm, nullptr, nullptr, d_bad);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = RSK_ERR_CUDA;
    }
    if (rc == RSK_OK) rc = rsk_check_faces(ctx, d_bad, "rsk_scene_from_geometry");
    rsk_dev_free(d_bad);
    if (rc != RSK_OK) { rsk_dev_free(d_tri); rsk_dev_free(d_nrm); return rc; }
    return rsk_scene_adopt(ctx, d_tri, d_nrm, g->n_tri, g->n_mesh, use_bvh, out);
}

extern "C" int rsk_emitters_from_geometry(rsk_geometry *g, double density, int32_t rays_per_cell, int32_t flip_faces,
                                          rsk_emitters **out, rsk_mesh_summary *summary_out) {
    RSK_REQUIRE(g && out && rays_per_cell > 0, "rsk_emitters_from_geometry: bad arguments");
    RSK_REQUIRE(g->n_mesh == 0 || summary_out, "rsk_emitters_from_geometry: null summary array");
    *out = nullptr;
    rsk_ctx *ctx = g->ctx;
    RskScope scope(ctx);
    const int64_t total = g->n_tri;
    rsk_emitters *em = new rsk_emitters();
    em->ctx = ctx; em->n_emit = g->n_mesh; em->rays_per_cell = rays_per_cell; em->n_tri_total = total;
    float *d_area = nullptr; double *d_cum = nullptr; MeshSummary *d_sum = nullptr; int *d_bad = nullptr;
    auto done = [&](int code) {
        rsk_dev_free(d_area); rsk_dev_free(d_cum); rsk_dev_free(d_sum); rsk_dev_free(d_bad);
        if (code != RSK_OK) rsk_emitters_destroy(em);
        return code;
    };
    int rc = rsk_dev_alloc(&em->tri, (size_t)total * 5);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&em->cdf, (size_t)total);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&em->desc, (size_t)g->n_mesh);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_area, (size_t)total);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_cum, (size_t)total);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_sum, (size_t)g->n_mesh);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_bad, 1);
    if (rc != RSK_OK) return done(rc);
    if (cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream) != cudaSuccess) return done(RSK_ERR_CUDA);
    if (total > 0) {
        rsk_prepare_triangles_kernel<<<rsk_blocks(total, 256), 256, 0, ctx->stream>>>(
            g->verts, g->faces, g->vert_off, g->tri_off, g->n_mesh, total, flip_faces ? 1 : 0, nullptr, nullptr, em->tri, d_area, d_bad);
        ctx->launches++;
    }
    rc = rsk_check_faces(ctx, d_bad, "rsk_emitters_from_geometry");
    if (rc != RSK_OK) return done(rc);
    std::vector<MeshSummary> h_sum((size_t)g->n_mesh);
    if (g->n_mesh > 0) {
        rsk_prepare_meshes_kernel<<<rsk_blocks((int64_t)g->n_mesh * 32, 128), 128, 0, ctx->stream>>>(
            em->tri, d_area, g->tri_off, g->n_mesh, d_cum, em->cdf, d_sum);
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(h_sum.data(), d_sum, h_sum.size() * sizeof(MeshSummary), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("emitter preparation failed: %s", cudaGetErrorString(e)); return done(RSK_ERR_CUDA); }
        memcpy(summary_out, h_sum.data(), h_sum.size() * sizeof(MeshSummary));
    }
    // grid side per emitter (helpers.py:8-11; prepared.py:305-311), then the QMC tables as in rsk_emitters_create
    em->h_desc.resize(g->n_mesh);
    for (int i = 0; i < g->n_mesh && rc == RSK_OK; ++i) {
        EmitterDesc &d = em->h_desc[i];
        d.tri_off = (int32_t)g->h_tri_off[i];
        d.n_tri = (int32_t)(g->h_tri_off[i + 1] - g->h_tri_off[i]);
        const double area = h_sum[i].total_area;
        d.g = area <= 0.0 ? 4 : std::max((int)std::ceil(std::sqrt(std::max(area, 0.0) * density)), 4);
        d.n_rays_once = (int64_t)d.g * d.g * rays_per_cell;
        em->max_rays_once = std::max(em->max_rays_once, d.n_rays_once);
        int64_t off = 0;
        rc = rsk_qmc_ensure_grid(ctx, d.g, &off);
        d.grid_off = area <= 0.0 ? -1 : (int32_t)off;          // zero area: all-zero QMC tables (prepared.py:278-287)
    }
    if (rc == RSK_OK) rc = rsk_qmc_ensure_halton(ctx, em->max_rays_once);
    if (rc == RSK_OK && g->n_mesh > 0) {
        cudaError_t e = cudaMemcpyAsync(em->desc, em->h_desc.data(), (size_t)g->n_mesh * sizeof(EmitterDesc), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("emitter upload failed: %s", cudaGetErrorString(e)); rc = RSK_ERR_CUDA; }
    }
    if (rc != RSK_OK) return done(rc);
    *out = em;
    return done(RSK_OK);
}

// ----------------------------------------------------------------------------- read-back (tests)

extern "C" int rsk_emitters_info(rsk_emitters *em, int32_t *g, int64_t *n_rays_once) {
    RSK_REQUIRE(em, "rsk_emitters_info: null emitters");
    for (int i = 0; i < em->n_emit; ++i) {
        if (g) g[i] = em->h_desc[i].g;
        if (n_rays_once) n_rays_once[i] = em->h_desc[i].n_rays_once;
    }
    return RSK_OK;
}

extern "C" int rsk_emitters_download_records(rsk_emitters *em, float *records, float *cdf) {
    RSK_REQUIRE(em, "rsk_emitters_download_records: null emitters");
    RskScope scope(em->ctx);
    RSK_CUDA(cudaStreamSynchronize(em->ctx->stream));
    if (records && em->n_tri_total) RSK_CUDA(cudaMemcpy(records, em->tri, (size_t)em->n_tri_total * 5 * sizeof(float4), cudaMemcpyDeviceToHost));
    if (cdf && em->n_tri_total) RSK_CUDA(cudaMemcpy(cdf, em->cdf, (size_t)em->n_tri_total * sizeof(float), cudaMemcpyDeviceToHost));
    return RSK_OK;
}

extern "C" int rsk_scene_download_triangles(rsk_scene *sc, float *tri, float *normals) {
    RSK_REQUIRE(sc, "rsk_scene_download_triangles: null scene");
    RskScope scope(sc->ctx);
    RSK_CUDA(cudaStreamSynchronize(sc->ctx->stream));
    if (tri && sc->n_tri) RSK_CUDA(cudaMemcpy(tri, sc->tri, (size_t)sc->n_tri * 3 * sizeof(float4), cudaMemcpyDeviceToHost));
    if (normals && sc->n_tri) RSK_CUDA(cudaMemcpy(normals, sc->nrm, (size_t)sc->n_tri * sizeof(float4), cudaMemcpyDeviceToHost));
    return RSK_OK;
}

// ----------------------------------------------------------------------------- surface masks

extern "C" int rsk_surface_masks(rsk_ctx *ctx, int32_t n_emit, int32_t n_surf, const uint8_t *planar, const float *plane_origin,
                                 const float *plane_normal, const float *plane_tol, const float *centers, const float *extents,
                                 uint8_t *active_out) {
    RSK_REQUIRE(ctx && n_emit >= 0 && n_surf >= 0, "rsk_surface_masks: bad arguments");
    const int64_t total = (int64_t)n_emit * n_surf;
    if (total == 0) return RSK_OK;
    RSK_REQUIRE(planar && plane_origin && plane_normal && plane_tol && centers && extents && active_out, "rsk_surface_masks: null array");
    RskScope scope(ctx);
    uint8_t *d_planar = nullptr, *d_out = nullptr;
    float *d_f = nullptr;                      // [po 3e][pn 3e][tol e][centers 3s][extents 3s]
    const size_t ne = (size_t)n_emit, ns = (size_t)n_surf, nf = 7 * ne + 6 * ns;
    int rc = rsk_dev_alloc(&d_planar, ne);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_f, nf);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_out, (size_t)total);
    if (rc == RSK_OK) {
        cudaError_t e = cudaMemcpyAsync(d_planar, planar, ne, cudaMemcpyHostToDevice, ctx->stream);
        const float *src[5] = {plane_origin, plane_normal, plane_tol, centers, extents};
        const size_t cnt[5] = {3 * ne, 3 * ne, ne, 3 * ns, 3 * ns};
        size_t off = 0;
        for (int k = 0; k < 5 && e == cudaSuccess; ++k) {
            e = cudaMemcpyAsync(d_f + off, src[k], cnt[k] * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            off += cnt[k];
        }
        if (e == cudaSuccess) {
            rsk_surface_masks_kernel<<<rsk_blocks(total, 256), 256, 0, ctx->stream>>>(
                n_emit, n_surf, d_planar, d_f, d_f + 3 * ne, d_f + 6 * ne, d_f + 7 * ne, d_f + 7 * ne + 3 * ns, d_out);
            ctx->launches++;
            e = cudaMemcpyAsync(active_out, d_out, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("rsk_surface_masks failed: %s", cudaGetErrorString(e)); rc = RSK_ERR_CUDA; }
    }
    rsk_dev_free(d_planar); rsk_dev_free(d_f); rsk_dev_free(d_out);
    return rc;
}
