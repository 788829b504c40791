// rsk_api.cu -- the extern "C" boundary of librsk_b200 (include/raystrack_b200.h).
#include <cstdarg>

#include "rsk_solve.cuh"

static thread_local char g_error[1024] = "";
thread_local cudaStream_t rsk_tl_stream = nullptr;
thread_local cudaMemPool_t rsk_tl_pool = nullptr;

void rsk_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char *rsk_last_error(void) { return g_error; }
extern "C" int rsk_abi_version(void) { return RSK_ABI_VERSION; }
#ifndef RSK_SOURCE_HASH
#define RSK_SOURCE_HASH "unknown"
#endif
extern "C" const char *rsk_source_hash(void) { return RSK_SOURCE_HASH; }

extern "C" int rsk_device_count(int *count) {
    RSK_REQUIRE(count, "rsk_device_count: null output");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return RSK_OK;
}

// ----------------------------------------------------------------------------- context

extern "C" int rsk_ctx_create(int device, void *stream, rsk_ctx **out) {
    RSK_REQUIRE(out, "rsk_ctx_create: null output");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        rsk_set_error("no CUDA device available: librsk_b200 has no CPU fallback");
        return RSK_ERR_NO_DEVICE;
    }
    RSK_REQUIRE(device >= 0 && device < n, "rsk_ctx_create: device ordinal out of range");
    RSK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RSK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        rsk_set_error("device %d is sm_%d%d; librsk_b200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return RSK_ERR_NO_DEVICE;
    }
    rsk_ctx *ctx = new rsk_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; rsk_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return RSK_ERR_CUDA; }
        ctx->own_stream = true;
    }
    {
        cudaError_t e = cudaEventCreate(&ctx->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
        if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_pinned, 16 * sizeof(int32_t));
        if (e != cudaSuccess) {
            rsk_set_error("rsk_ctx_create: %s", cudaGetErrorString(e));
            rsk_ctx_destroy(ctx);
            return e == cudaErrorMemoryAllocation ? RSK_ERR_OOM : RSK_ERR_CUDA;
        }
    }
    const char *pool_env = getenv("RSK_PRIVATE_POOL");      // 0: allocate from the device's default pool instead (A/B timing)
    if (pool_env && atoi(pool_env) == 0) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    } else {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&ctx->pool, &props);
        if (e == cudaSuccess) {
            unsigned long long keep = ~0ull;
            e = cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        if (e != cudaSuccess) {
            rsk_set_error("rsk_ctx_create: memory pool: %s", cudaGetErrorString(e));
            rsk_ctx_destroy(ctx);
            return RSK_ERR_CUDA;
        }
    }
    *out = ctx;
    return RSK_OK;
}

extern "C" int rsk_ctx_destroy(rsk_ctx *ctx) {
    if (!ctx) return RSK_OK;
    if (ctx->comm) rsk_comm_destroy(ctx);
    RskScope scope(ctx);
    rsk_dev_free(ctx->halton);
    rsk_dev_free(ctx->grid);
    rsk_dev_free(ctx->csr_cols);
    rsk_dev_free(ctx->csr_vals);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    if (ctx->l2_flush) cudaFree(ctx->l2_flush);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); cudaEventDestroy(ctx->ev_join); }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RSK_OK;
}

// Second stream of pipelined solves, and the event that orders `stream` after whatever was enqueued on it.
int rsk_ctx_stream2(rsk_ctx *ctx, cudaStream_t *out) {
    if (!ctx->stream2) {
        RSK_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        RSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    *out = ctx->stream2;
    return RSK_OK;
}

int rsk_ctx_join(rsk_ctx *ctx) {
    if (!ctx->stream2) return RSK_OK;
    RSK_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream2));
    RSK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return RSK_OK;
}

extern "C" int rsk_ctx_synchronize(rsk_ctx *ctx) {
    RSK_REQUIRE(ctx, "null context");
    if (ctx->stream2) RSK_CUDA(cudaStreamSynchronize(ctx->stream2));
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_ctx_set_l2_flush(rsk_ctx *ctx, int64_t bytes) {
    RSK_REQUIRE(ctx && bytes >= 0, "rsk_ctx_set_l2_flush: bad arguments");
    RskScope scope(ctx);
    RSK_TRY(rsk_ctx_synchronize(ctx));
    if (ctx->l2_flush) { cudaFree(ctx->l2_flush); ctx->l2_flush = nullptr; ctx->l2_flush_bytes = 0; }
    if (bytes > 0) {
        RSK_CUDA(cudaMalloc(&ctx->l2_flush, (size_t)bytes));
        ctx->l2_flush_bytes = (size_t)bytes;
    }
    return RSK_OK;
}

extern "C" int rsk_ctx_timer_start(rsk_ctx *ctx) {
    RSK_REQUIRE(ctx, "null context");
    RSK_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_ctx_timer_stop(rsk_ctx *ctx, float *ms) {
    RSK_REQUIRE(ctx && ms, "rsk_ctx_timer_stop: bad arguments");
    RSK_TRY(rsk_ctx_join(ctx));                    // the stopwatch covers the work of pipelined solves on the second stream
    RSK_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    RSK_CUDA(cudaEventSynchronize(ctx->ev1));
    RSK_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return RSK_OK;
}

extern "C" int rsk_ctx_launch_count(rsk_ctx *ctx, int64_t *count) {
    RSK_REQUIRE(ctx && count, "rsk_ctx_launch_count: bad arguments");
    *count = ctx->launches;
    return RSK_OK;
}

extern "C" int rsk_ctx_device_info(rsk_ctx *ctx, char *name, int64_t *info) {
    RSK_REQUIRE(ctx, "null context");
    cudaDeviceProp prop;
    RSK_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    if (name) { strncpy(name, prop.name, 255); name[255] = 0; }
    if (info) { info[0] = prop.multiProcessorCount; info[1] = prop.major; info[2] = prop.minor; info[3] = (int64_t)prop.totalGlobalMem; }
    return RSK_OK;
}

// ----------------------------------------------------------------------------- scene

// pack the reference's SoA scene arrays into traversal records on the device
__global__ void rsk_pack_scene_kernel(const float *v0, const float *e1, const float *e2, const float *nrm, const int32_t *sid,
                                      int64_t n, float4 *tri, float4 *nrm4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sbits = __int_as_float(sid[i]);
    tri[3 * i] = make_float4(v0[3 * i], v0[3 * i + 1], v0[3 * i + 2], sbits);
    tri[3 * i + 1] = make_float4(e1[3 * i], e1[3 * i + 1], e1[3 * i + 2], 0.f);
    tri[3 * i + 2] = make_float4(e2[3 * i], e2[3 * i + 1], e2[3 * i + 2], 0.f);
    nrm4[i] = make_float4(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], sbits);
}

extern "C" int rsk_scene_create(rsk_ctx *ctx, const float *v0, const float *e1, const float *e2, const float *normals,
                                const int32_t *sid, int64_t n_tri, int32_t n_surf, int32_t use_bvh, rsk_scene **out) {
    RSK_REQUIRE(ctx && out, "rsk_scene_create: null context/output");
    RSK_REQUIRE(n_tri >= 0 && n_surf >= 0, "rsk_scene_create: negative sizes");
    RSK_REQUIRE(n_tri == 0 || (v0 && e1 && e2 && normals && sid), "rsk_scene_create: null arrays");
    RSK_REQUIRE(n_tri < (1ll << 30), "rsk_scene_create: too many triangles");
    *out = nullptr;
    RskScope scope(ctx);
    for (int64_t i = 0; i < n_tri; ++i) RSK_REQUIRE(sid[i] >= 0 && sid[i] < n_surf, "rsk_scene_create: sid out of range");
    // the caller's five arrays go to the device as they are; records are packed there
    float *raw = nullptr; int32_t *d_sid = nullptr; float4 *d_tri = nullptr, *d_nrm = nullptr;
    auto fail = [&](int code) { rsk_dev_free(raw); rsk_dev_free(d_sid); rsk_dev_free(d_tri); rsk_dev_free(d_nrm); return code; };
    int rc = rsk_dev_alloc(&raw, (size_t)n_tri * 12);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_sid, (size_t)n_tri);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_tri, (size_t)n_tri * 3);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&d_nrm, (size_t)n_tri);
    if (rc != RSK_OK) return fail(rc);
    if (n_tri > 0) {
        const size_t b3 = (size_t)n_tri * 3 * sizeof(float);
        const float *src[4] = {v0, e1, e2, normals};
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 4 && e == cudaSuccess; ++k)
            e = cudaMemcpyAsync(raw + (size_t)k * n_tri * 3, src[k], b3, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_sid, sid, (size_t)n_tri * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("scene upload failed: %s", cudaGetErrorString(e)); return fail(RSK_ERR_CUDA); }
        rsk_pack_scene_kernel<<<rsk_blocks(n_tri, 256), 256, 0, ctx->stream>>>(raw, raw + (size_t)n_tri * 3, raw + (size_t)n_tri * 6,
                                                                             raw + (size_t)n_tri * 9, d_sid, n_tri, d_tri, d_nrm);
        ctx->launches++;
    }
    rsk_dev_free(raw); raw = nullptr;
    rsk_dev_free(d_sid); d_sid = nullptr;
    return rsk_scene_adopt(ctx, d_tri, d_nrm, n_tri, n_surf, use_bvh, out);
}

// Turn packed device records (ownership passes to the scene) into a scene: build the BVH or keep input order.
int rsk_scene_adopt(rsk_ctx *ctx, float4 *d_tri, float4 *d_nrm, int64_t n_tri, int32_t n_surf, int32_t use_bvh, rsk_scene **out) {
    rsk_scene *sc = new rsk_scene();
    sc->ctx = ctx;
    sc->n_tri = n_tri;
    sc->n_surf = n_surf;
    sc->use_bvh = (use_bvh && n_tri > 0) ? 1 : 0;
    if (sc->use_bvh) {
        const int rc = rsk_bvh_build(sc, d_tri, d_nrm);
        rsk_dev_free(d_tri);
        rsk_dev_free(d_nrm);
        if (rc != RSK_OK) { rsk_scene_destroy(sc); return rc; }
    } else {
        sc->tri = d_tri;
        sc->nrm = d_nrm;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("scene upload failed: %s", cudaGetErrorString(e)); rsk_scene_destroy(sc); return RSK_ERR_CUDA; }
    }
    *out = sc;
    return RSK_OK;
}

extern "C" int rsk_scene_destroy(rsk_scene *sc) {
    if (!sc) return RSK_OK;
    RskScope scope(sc->ctx);
    rsk_dev_free(sc->tri);
    rsk_dev_free(sc->nrm);
    rsk_dev_free(sc->nodes);
    rsk_dev_free(sc->tri_index);
    delete sc;
    return RSK_OK;
}

extern "C" int rsk_scene_info(rsk_scene *sc, int64_t *info) {
    RSK_REQUIRE(sc && info, "rsk_scene_info: bad arguments");
    info[0] = sc->n_tri; info[1] = sc->n_surf; info[2] = sc->use_bvh; info[3] = sc->n_nodes;
    info[4] = sc->n_nodes * (int64_t)sizeof(WideNode); info[5] = sc->n_tri * 64; info[6] = sc->depth; info[7] = sc->build_us;
    return RSK_OK;
}

extern "C" int rsk_scene_download_bvh(rsk_scene *sc, void *nodes, int32_t *tri_index) {
    RSK_REQUIRE(sc && sc->use_bvh, "rsk_scene_download_bvh: scene has no BVH");
    RskScope scope(sc->ctx);
    RSK_CUDA(cudaStreamSynchronize(sc->ctx->stream));
    if (nodes) RSK_CUDA(cudaMemcpy(nodes, sc->nodes, sc->n_nodes * sizeof(WideNode), cudaMemcpyDeviceToHost));
    if (tri_index) RSK_CUDA(cudaMemcpy(tri_index, sc->tri_index, sc->n_tri * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RSK_OK;
}

// ----------------------------------------------------------------------------- emitters

__global__ void rsk_pack_emitters_kernel(const float *a, const float *e1, const float *e2, const float *u, const float *v,
                                         const float *n, const float *eps, int64_t total, float4 *rec) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    rec[5 * t + 0] = make_float4(a[3 * t], a[3 * t + 1], a[3 * t + 2], eps[t]);
    rec[5 * t + 1] = make_float4(e1[3 * t], e1[3 * t + 1], e1[3 * t + 2], n[3 * t]);
    rec[5 * t + 2] = make_float4(e2[3 * t], e2[3 * t + 1], e2[3 * t + 2], n[3 * t + 1]);
    rec[5 * t + 3] = make_float4(u[3 * t], u[3 * t + 1], u[3 * t + 2], n[3 * t + 2]);
    rec[5 * t + 4] = make_float4(v[3 * t], v[3 * t + 1], v[3 * t + 2], 0.f);
}

extern "C" int rsk_emitters_create(rsk_ctx *ctx, int32_t n_emit, const int64_t *tri_offset,
                                   const float *tri_a, const float *tri_e1, const float *tri_e2,
                                   const float *tri_u, const float *tri_v, const float *tri_n,
                                   const float *tri_eps, const float *cdf,
                                   const int32_t *g, int32_t rays_per_cell, rsk_emitters **out) {
    RSK_REQUIRE(ctx && out && n_emit >= 0 && rays_per_cell > 0, "rsk_emitters_create: bad arguments");
    RSK_REQUIRE(n_emit == 0 || (tri_offset && g), "rsk_emitters_create: null tables");
    *out = nullptr;
    RskScope scope(ctx);
    const int64_t total = n_emit ? tri_offset[n_emit] : 0;
    RSK_REQUIRE(total < (1ll << 31), "rsk_emitters_create: too many emitter triangles");
    RSK_REQUIRE(total == 0 || (tri_a && tri_e1 && tri_e2 && tri_u && tri_v && tri_n && tri_eps && cdf), "rsk_emitters_create: null arrays");
    rsk_emitters *em = new rsk_emitters();
    em->ctx = ctx;
    em->n_emit = n_emit;
    em->rays_per_cell = rays_per_cell;
    em->n_tri_total = total;
    em->h_desc.resize(n_emit);
    int rc = RSK_OK;
    for (int i = 0; i < n_emit && rc == RSK_OK; ++i) {
        EmitterDesc &d = em->h_desc[i];
        d.tri_off = (int32_t)tri_offset[i];
        d.n_tri = (int32_t)(tri_offset[i + 1] - tri_offset[i]);
        const bool zero_tables = g[i] < 0;        // -g: the reference's zero-area emitter, all-zero QMC tables
        d.g = zero_tables ? -g[i] : g[i];
        if (d.g < 1 || d.n_tri < 0) { rsk_set_error("rsk_emitters_create: bad grid/triangle count for emitter %d", i); rc = RSK_ERR_INVALID; break; }
        d.n_rays_once = (int64_t)d.g * d.g * rays_per_cell;
        em->max_rays_once = std::max(em->max_rays_once, d.n_rays_once);
        int64_t off = 0;
        rc = rsk_qmc_ensure_grid(ctx, d.g, &off);
        d.grid_off = zero_tables ? -1 : (int32_t)off;
    }
    if (rc == RSK_OK) rc = rsk_qmc_ensure_halton(ctx, em->max_rays_once);
    float *raw = nullptr;
    if (rc == RSK_OK) rc = rsk_dev_alloc(&raw, (size_t)total * 19);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&em->tri, (size_t)total * 5);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&em->cdf, (size_t)total);
    if (rc == RSK_OK) rc = rsk_dev_alloc(&em->desc, (size_t)n_emit);
    if (rc == RSK_OK) {
        cudaError_t e = cudaSuccess;
        if (total > 0) {
            const float *src[6] = {tri_a, tri_e1, tri_e2, tri_u, tri_v, tri_n};
            for (int k = 0; k < 6 && e == cudaSuccess; ++k)
                e = cudaMemcpyAsync(raw + (size_t)k * total * 3, src[k], (size_t)total * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(raw + (size_t)total * 18, tri_eps, total * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(em->cdf, cdf, total * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) {
                rsk_pack_emitters_kernel<<<rsk_blocks(total, 256), 256, 0, ctx->stream>>>(
                    raw, raw + (size_t)total * 3, raw + (size_t)total * 6, raw + (size_t)total * 9, raw + (size_t)total * 12,
                    raw + (size_t)total * 15, raw + (size_t)total * 18, total, em->tri);
                ctx->launches++;
            }
        }
        if (e == cudaSuccess && n_emit > 0)
            e = cudaMemcpyAsync(em->desc, em->h_desc.data(), n_emit * sizeof(EmitterDesc), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("emitter upload failed: %s", cudaGetErrorString(e)); rc = RSK_ERR_CUDA; }
    }
    rsk_dev_free(raw);
    if (rc != RSK_OK) { rsk_emitters_destroy(em); return rc; }
    *out = em;
    return RSK_OK;
}

extern "C" int rsk_emitters_destroy(rsk_emitters *em) {
    if (!em) return RSK_OK;
    RskScope scope(em->ctx);
    rsk_dev_free(em->tri);
    rsk_dev_free(em->cdf);
    rsk_dev_free(em->desc);
    delete em;
    return RSK_OK;
}

extern "C" int rsk_emitters_download_tables(rsk_emitters *em, int64_t n, float *dims, int32_t g, float *grid_u, float *grid_v) {
    RSK_REQUIRE(em, "null emitters");
    rsk_ctx *ctx = em->ctx;
    RskScope scope(ctx);
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (dims) {
        RSK_REQUIRE(n <= ctx->halton_cap, "rsk_emitters_download_tables: n exceeds the cached table");
        for (int r = 0; r < 5; ++r)
            RSK_CUDA(cudaMemcpy(dims + r * n, ctx->halton + r * ctx->halton_cap, n * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (grid_u || grid_v) {
        auto it = ctx->grid_index.find(g);
        RSK_REQUIRE(it != ctx->grid_index.end(), "rsk_emitters_download_tables: grid size not cached");
        std::vector<float2> tmp((size_t)it->second.second);
        RSK_CUDA(cudaMemcpy(tmp.data(), ctx->grid + it->second.first, tmp.size() * sizeof(float2), cudaMemcpyDeviceToHost));
        for (size_t c = 0; c < tmp.size(); ++c) {
            if (grid_u) grid_u[c] = tmp[c].x;
            if (grid_v) grid_v[c] = tmp[c].y;
        }
    }
    return RSK_OK;
}

// ----------------------------------------------------------------------------- helpers for masks / uploads

static void rsk_pack_mask(const uint8_t *active, int n_surf, int emit_sid, int min_sid, uint32_t *words) {
    // utils/cpu_trace.py:45-51 `_skip_surface` folded into one bit per surface (bit set = NOT skipped)
    const int nw = (n_surf + 31) / 32;
    for (int w = 0; w < nw; ++w) words[w] = 0;
    for (int s = 0; s < n_surf; ++s)
        if (active[s] != 0 && s >= min_sid && s != emit_sid) words[s >> 5] |= 1u << (s & 31);
}

template <typename T>
static int rsk_upload(rsk_ctx *ctx, T **dst, const T *src, size_t count) {
    RSK_TRY(rsk_dev_alloc(dst, count));
    if (count) RSK_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return RSK_OK;
}

// ----------------------------------------------------------------------------- per-ray hook

extern "C" int rsk_trace_rays(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, int32_t emitter,
                              const uint8_t *surf_active, int32_t emit_sid, int32_t min_sid, const float *cp,
                              int32_t mode, int64_t first_ray, int64_t n_rays,
                              float *orig, float *dirs, int32_t *hit_sid, uint8_t *hit_front) {
    RSK_REQUIRE(ctx && scene && em && surf_active && cp, "rsk_trace_rays: null argument");
    RSK_REQUIRE(emitter >= 0 && emitter < em->n_emit, "rsk_trace_rays: emitter out of range");
    RSK_REQUIRE(mode == 0 || mode == 1, "rsk_trace_rays: mode must be 0 or 1");
    const int64_t n_once = em->h_desc[emitter].n_rays_once;
    RSK_REQUIRE(first_ray >= 0 && n_rays >= 0 && first_ray + n_rays <= n_once, "rsk_trace_rays: ray range out of bounds");
    if (n_rays == 0) return RSK_OK;
    RskScope scope(ctx);
    const int nw = (scene->n_surf + 31) / 32;
    std::vector<uint32_t> mask(std::max(nw, 1));
    rsk_pack_mask(surf_active, scene->n_surf, emit_sid, min_sid, mask.data());
    const int32_t zero = 0;
    const int64_t range[2] = {first_ray, first_ray + n_rays};
    std::vector<TileDesc> tiles;
    int tile_rays = 0;
    rsk_build_tiles(range, range + 1, 1, ctx->sm_count, tiles, &tile_rays);
    const int64_t n_tiles = (int64_t)tiles.size();
    const int32_t msid1 = mode == MODE_MATRIX ? min_sid : 0;

    uint32_t *d_mask = nullptr; float *d_cp = nullptr; int32_t *d_ids = nullptr, *d_zero = nullptr; TileDesc *d_tiles = nullptr; int64_t *d_range = nullptr; int32_t *d_msid = nullptr;
    float *d_orig = nullptr, *d_dirs = nullptr; int32_t *d_hit = nullptr; uint8_t *d_front = nullptr;
    auto cleanup = [&]() { rsk_dev_free(d_mask); rsk_dev_free(d_cp); rsk_dev_free(d_ids); rsk_dev_free(d_zero); rsk_dev_free(d_tiles); rsk_dev_free(d_range); rsk_dev_free(d_msid);
                           rsk_dev_free(d_orig); rsk_dev_free(d_dirs); rsk_dev_free(d_hit); rsk_dev_free(d_front); };
    int rc = RSK_OK;
#define T_TRY(expr) do { rc = (expr); if (rc != RSK_OK) { cleanup(); return rc; } } while (0)
#define T_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rsk_set_error("%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); return RSK_ERR_CUDA; } } while (0)
    T_TRY(rsk_upload(ctx, &d_mask, mask.data(), mask.size()));
    T_TRY(rsk_upload(ctx, &d_cp, cp, 7));
    T_TRY(rsk_upload(ctx, &d_ids, &emitter, 1));
    T_TRY(rsk_upload(ctx, &d_zero, &zero, 1));
    T_TRY(rsk_upload(ctx, &d_tiles, tiles.data(), tiles.size()));
    T_TRY(rsk_upload(ctx, &d_range, range, 2));
    T_TRY(rsk_upload(ctx, &d_msid, &msid1, 1));
    if (orig) T_TRY(rsk_dev_alloc(&d_orig, (size_t)n_rays * 3));
    if (dirs) T_TRY(rsk_dev_alloc(&d_dirs, (size_t)n_rays * 3));
    if (hit_sid) T_TRY(rsk_dev_alloc(&d_hit, (size_t)n_rays));
    if (hit_front) T_TRY(rsk_dev_alloc(&d_front, (size_t)n_rays));

    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.sc = scene->view();
    a.ev = em->view();
    a.emit_ids = d_ids; a.tiles = d_tiles; a.n_local = 1; a.tile_rays = tile_rays; a.surf_mask = d_mask;
    a.cp_table = d_cp; a.rot_base = d_zero; a.iters_done = d_zero; a.iter_index = -1; a.done = nullptr; a.tally = nullptr;
    a.n_hist = mode == MODE_MATRIX ? 2 * scene->n_surf : RSK_TREGENZA_BINS;
    a.ray_begin = d_range; a.ray_end = d_range + 1; a.dbg_base = first_ray; a.min_sid = d_msid;
    a.dbg_orig = d_orig; a.dbg_dirs = d_dirs; a.dbg_hit = d_hit; a.dbg_front = d_front;
    T_TRY(rsk_launch_trace(ctx, a, mode, n_tiles, ctx->stream));
    if (orig) T_CUDA(cudaMemcpyAsync(orig, d_orig, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (dirs) T_CUDA(cudaMemcpyAsync(dirs, d_dirs, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (hit_sid) T_CUDA(cudaMemcpyAsync(hit_sid, d_hit, (size_t)n_rays * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (hit_front) T_CUDA(cudaMemcpyAsync(hit_front, d_front, (size_t)n_rays, cudaMemcpyDeviceToHost, ctx->stream));
    T_CUDA(cudaStreamSynchronize(ctx->stream));
    T_CUDA(cudaGetLastError());
    cleanup();
#undef T_TRY
#undef T_CUDA
    return RSK_OK;
}

// ----------------------------------------------------------------------------- cost of an emitter's rays

// Relative cost per ray of a set of emitters: one launch traces the first `sample_rays` rays of each (closest hit, the
// masks and skip rules of a matrix solve) and every CTA adds the SM clock ticks it was resident to its job.  Multi-GPU
// plans weight emitters by rays x cost instead of rays alone.  ticks[k] covers rays_out[k] = min(sample_rays, rays of k).
extern "C" int rsk_emitter_costs(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                                 const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid, const float *cp,
                                 int64_t sample_rays, int64_t *ticks, int64_t *rays_out) {
    RSK_REQUIRE(ctx && scene && em && ticks && cp && sample_rays > 0 && n_local >= 0, "rsk_emitter_costs: bad arguments");
    RSK_REQUIRE(n_local == 0 || (emit_ids && surf_active && emit_sid && min_sid), "rsk_emitter_costs: null arrays");
    if (n_local == 0) return RSK_OK;
    RskScope scope(ctx);
    const int nw = std::max((scene->n_surf + 31) / 32, 1);
    std::vector<uint32_t> mask((size_t)n_local * nw);
    std::vector<int64_t> rbeg(n_local, 0), rend(n_local);
    std::vector<int32_t> zeros(n_local, 0);
    for (int k = 0; k < n_local; ++k) {
        RSK_REQUIRE(emit_ids[k] >= 0 && emit_ids[k] < em->n_emit, "rsk_emitter_costs: emitter id out of range");
        rsk_pack_mask(surf_active + (size_t)k * scene->n_surf, scene->n_surf, emit_sid[k], min_sid[k], mask.data() + (size_t)k * nw);
        rend[k] = std::min<int64_t>(sample_rays, em->h_desc[emit_ids[k]].n_rays_once);
        if (rays_out) rays_out[k] = rend[k];
    }
    std::vector<TileDesc> tiles;
    for (int k = 0; k < n_local; ++k)                     // one pass of equal tiles: every CTA sees the same neighbours
        for (int64_t b = 0; b < rend[k]; b += 2048)
            tiles.push_back(TileDesc{k, (int32_t)std::min<int64_t>(2048, rend[k] - b), b});
    uint32_t *d_mask = nullptr; float *d_cp = nullptr; int32_t *d_ids = nullptr, *d_zero = nullptr, *d_msid = nullptr;
    TileDesc *d_tiles = nullptr; int64_t *d_beg = nullptr, *d_end = nullptr; unsigned long long *d_ticks = nullptr;
    auto cleanup = [&]() { rsk_dev_free(d_mask); rsk_dev_free(d_cp); rsk_dev_free(d_ids); rsk_dev_free(d_zero); rsk_dev_free(d_msid);
                           rsk_dev_free(d_tiles); rsk_dev_free(d_beg); rsk_dev_free(d_end); rsk_dev_free(d_ticks); };
    int rc = RSK_OK;
#define E_TRY(expr) do { rc = (expr); if (rc != RSK_OK) { cleanup(); return rc; } } while (0)
#define E_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rsk_set_error("%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); return RSK_ERR_CUDA; } } while (0)
    E_TRY(rsk_upload(ctx, &d_mask, mask.data(), mask.size()));
    E_TRY(rsk_upload(ctx, &d_cp, cp, 7));
    E_TRY(rsk_upload(ctx, &d_ids, emit_ids, (size_t)n_local));
    E_TRY(rsk_upload(ctx, &d_zero, zeros.data(), zeros.size()));
    E_TRY(rsk_upload(ctx, &d_msid, min_sid, (size_t)n_local));
    E_TRY(rsk_upload(ctx, &d_tiles, tiles.data(), tiles.size()));
    E_TRY(rsk_upload(ctx, &d_beg, rbeg.data(), rbeg.size()));
    E_TRY(rsk_upload(ctx, &d_end, rend.data(), rend.size()));
    E_TRY(rsk_dev_alloc(&d_ticks, (size_t)n_local));
    E_CUDA(cudaMemsetAsync(d_ticks, 0, (size_t)n_local * 8, ctx->stream));
    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.sc = scene->view();
    a.ev = em->view();
    a.emit_ids = d_ids; a.tiles = d_tiles; a.n_local = n_local; a.tile_rays = 2048; a.surf_mask = d_mask;
    a.cp_table = d_cp; a.rot_base = d_zero; a.iters_done = d_zero; a.iter_index = 0; a.done = nullptr; a.tally = nullptr;
    a.n_hist = 2 * scene->n_surf; a.ray_begin = d_beg; a.ray_end = d_end; a.min_sid = d_msid; a.job_ticks = d_ticks;
    E_TRY(rsk_launch_trace(ctx, a, MODE_MATRIX, (int64_t)tiles.size(), ctx->stream));
    E_CUDA(cudaMemcpyAsync(ticks, d_ticks, (size_t)n_local * 8, cudaMemcpyDeviceToHost, ctx->stream));
    E_CUDA(cudaStreamSynchronize(ctx->stream));
    cleanup();
#undef E_TRY
#undef E_CUDA
    return RSK_OK;
}

// ----------------------------------------------------------------------------- solves


static int rsk_solve_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, int mode, int discrete,
                           const int32_t *emit_ids, int32_t n_local, const uint8_t *surf_active,
                           const int32_t *emit_sid, const int32_t *min_sid,
                           const float *cp_table, int32_t n_rot, const int32_t *rot_base, const int64_t *ray_range,
                           const rsk_solve_params *params, rsk_solve **out) {
    RSK_REQUIRE(ctx && scene && em && params && out, "solve begin: null argument");
    RSK_REQUIRE(n_local >= 0 && n_rot >= 0, "solve begin: negative sizes");
    RSK_REQUIRE(n_local == 0 || (emit_ids && surf_active && cp_table && rot_base), "solve begin: null arrays");
    RSK_REQUIRE(params->tol_mode == 0 || params->tol_mode == 1, "solve begin: tol_mode must be 0 (stderr) or 1 (delta)");
    *out = nullptr;
    RskScope scope(ctx);
    rsk_solve *s = new rsk_solve();
    s->ctx = ctx; s->scene = scene; s->em = em; s->mode = mode; s->n_local = n_local; s->discrete = discrete;
    s->p = *params;
    s->n_hist = mode == MODE_MATRIX ? 2 * scene->n_surf : (discrete ? RSK_TREGENZA_BINS : 1);
    const int nw = std::max((scene->n_surf + 31) / 32, 1);
    std::vector<uint32_t> mask((size_t)n_local * nw);
    std::vector<int64_t> once(n_local), rbeg(n_local), rend(n_local);
    std::vector<TileDesc> tiles;
    std::vector<int32_t> msid(n_local, 0);
    int rc = RSK_OK;
    for (int k = 0; k < n_local; ++k) {
        if (emit_ids[k] < 0 || emit_ids[k] >= em->n_emit) { rsk_set_error("solve begin: emitter id out of range"); rc = RSK_ERR_INVALID; break; }
        if (rot_base[k] < 0 || (int64_t)rot_base[k] + std::max(params->max_iters, 0) > n_rot) { rsk_set_error("solve begin: rotation table too short"); rc = RSK_ERR_INVALID; break; }
        const int es = emit_sid ? emit_sid[k] : emit_ids[k], ms = min_sid ? min_sid[k] : 0;
        msid[k] = ms;
        rsk_pack_mask(surf_active + (size_t)k * scene->n_surf, scene->n_surf, es, ms, mask.data() + (size_t)k * nw);
        once[k] = em->h_desc[emit_ids[k]].n_rays_once;
        rbeg[k] = ray_range ? ray_range[2 * k] : 0;
        rend[k] = ray_range ? ray_range[2 * k + 1] : once[k];
        if (rbeg[k] < 0 || rend[k] < rbeg[k] || rend[k] > once[k]) { rsk_set_error("solve begin: ray range out of bounds"); rc = RSK_ERR_INVALID; break; }
    }
    static int want_pipeline = -1;           // RSK_PIPELINE=0 turns pipelined stepping off
    if (want_pipeline < 0) { const char *e = getenv("RSK_PIPELINE"); want_pipeline = (e && atoi(e) == 0) ? 0 : 1; }
    const bool pipeline = want_pipeline && n_local > 0 && params->max_iters > 1;
    if (rc == RSK_OK) rsk_build_tiles(rbeg.data(), rend.data(), n_local, ctx->sm_count, tiles, &s->tile_rays, pipeline);
    s->n_tiles = (int64_t)tiles.size();
    const size_t nh = (size_t)n_local * s->n_hist;
    auto fail = [&](int code) { rsk_solve_destroy(s); return code; };
    if (rc != RSK_OK) return fail(rc);
#define S_TRY(expr) do { rc = (expr); if (rc != RSK_OK) return fail(rc); } while (0)
#define S_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rsk_set_error("%s failed: %s", #call, cudaGetErrorString(e__)); return fail(e__ == cudaErrorMemoryAllocation ? RSK_ERR_OOM : RSK_ERR_CUDA); } } while (0)
    S_TRY(rsk_upload(ctx, &s->emit_ids, emit_ids, n_local));
    S_TRY(rsk_upload(ctx, &s->rot_base, rot_base, n_local));
    S_TRY(rsk_upload(ctx, &s->min_sid, msid.data(), msid.size()));
    S_TRY(rsk_upload(ctx, &s->tiles, tiles.data(), tiles.size()));
    S_TRY(rsk_upload(ctx, &s->n_rays_once, once.data(), once.size()));
    S_TRY(rsk_upload(ctx, &s->ray_begin, rbeg.data(), rbeg.size()));
    S_TRY(rsk_upload(ctx, &s->ray_end, rend.data(), rend.size()));
    S_TRY(rsk_upload(ctx, &s->mask, mask.data(), mask.size()));
    S_TRY(rsk_upload(ctx, &s->cp_table, cp_table, (size_t)n_rot * 7));
    S_TRY(rsk_dev_alloc(&s->iters_done, n_local)); S_TRY(rsk_dev_alloc(&s->done, n_local));
    S_TRY(rsk_dev_alloc(&s->not_conv, n_local)); S_TRY(rsk_dev_alloc(&s->have_prev, n_local));
    S_TRY(rsk_dev_alloc(&s->total_rays, n_local));
    S_TRY(rsk_dev_alloc(&s->iter_tally, nh)); S_TRY(rsk_dev_alloc(&s->total, nh));
    S_TRY(rsk_dev_alloc(&s->mean, nh)); S_TRY(rsk_dev_alloc(&s->m2, nh));
    if (params->tol_mode == 1) S_TRY(rsk_dev_alloc(&s->prev, nh));
    S_TRY(rsk_dev_alloc(&s->n_active, 1)); S_TRY(rsk_dev_alloc(&s->rays_traced, 1));
    s->h_pinned = ctx->h_pinned;
    if (!s->h_pinned) { rsk_set_error("solve begin: context has no pinned scratch"); return fail(RSK_ERR_OOM); }
    cudaStream_t st = ctx->stream;
    const size_t nl = std::max(n_local, 1);
    S_CUDA(cudaMemsetAsync(s->iters_done, 0, nl * sizeof(int32_t), st));
    S_CUDA(cudaMemsetAsync(s->done, 0, nl * sizeof(int32_t), st));
    S_CUDA(cudaMemsetAsync(s->not_conv, 0, nl * sizeof(int32_t), st));
    S_CUDA(cudaMemsetAsync(s->have_prev, 0, nl * sizeof(int32_t), st));
    S_CUDA(cudaMemsetAsync(s->total_rays, 0, nl * sizeof(int64_t), st));
    const size_t nh1 = std::max(nh, (size_t)1);
    S_CUDA(cudaMemsetAsync(s->iter_tally, 0, nh1 * 8, st));
    S_CUDA(cudaMemsetAsync(s->total, 0, nh1 * 8, st));
    S_CUDA(cudaMemsetAsync(s->mean, 0, nh1 * 8, st));
    S_CUDA(cudaMemsetAsync(s->m2, 0, nh1 * 8, st));
    if (s->prev) S_CUDA(cudaMemsetAsync(s->prev, 0, nh1 * 8, st));
    S_CUDA(cudaMemsetAsync(s->rays_traced, 0, 8, st));
    S_CUDA(cudaStreamSynchronize(st));
#undef S_TRY
#undef S_CUDA
    s->last_active = (params->max_iters > 0) ? n_local : 0;
    // Pipelined stepping (RSK_PIPELINE=0 turns it off): a second tally buffer for the odd iterations
    {
        if (pipeline) {
            cudaStream_t s2;
            bool ok = rsk_ctx_stream2(ctx, &s2) == RSK_OK && rsk_dev_alloc(&s->iter_tally2, nh) == RSK_OK;
            ok = ok && cudaMemsetAsync(s->iter_tally2, 0, std::max(nh, (size_t)1) * 8, ctx->stream) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&s->ev_fold, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
            if (!ok) { rsk_set_error("solve begin: pipeline buffers: %s", cudaGetErrorString(cudaGetLastError())); return fail(RSK_ERR_CUDA); }
            s->pipelined = true;
        }
    }
    *out = s;
    return RSK_OK;
}

// One iteration = phase A (fused raygen+trace+tally of every unconverged job) + phase B (fold the iteration
// tallies into totals / Welford statistics, decide convergence per job).  The phases are separate entry points
// so that a multi-GPU caller can all-reduce the iteration tallies of ray-split emitters in between.
// Stream and tally buffer of the iteration that is being enqueued.  A dual solve keeps the pipeline state on its
// matrix side (`primary`); both sides must be pipelined for the pair to be.
static inline rsk_solve *rsk_pipe_owner(rsk_solve *s) { return s->primary ? s->primary : s; }
static inline bool rsk_pipe_on(rsk_solve *s) {
    rsk_solve *o = rsk_pipe_owner(s);
    return o->pipelined && (!o->twin || o->twin->pipelined);
}
static inline cudaStream_t rsk_pipe_stream(rsk_solve *s) {
    rsk_solve *o = rsk_pipe_owner(s);
    return (rsk_pipe_on(s) && o->cur) ? s->ctx->stream2 : s->ctx->stream;
}
static inline unsigned long long *rsk_pipe_tally(rsk_solve *s) {
    rsk_solve *o = rsk_pipe_owner(s);
    return (rsk_pipe_on(s) && o->cur) ? s->iter_tally2 : s->iter_tally;
}

cudaStream_t rsk_solve_current_stream(rsk_solve *s) { return rsk_pipe_stream(s); }

static int rsk_solve_enqueue_trace_impl(rsk_solve *s) {
    rsk_ctx *ctx = s->ctx;
    if (s->n_local == 0 || (s->p.max_iters <= 0 && !(s->twin && s->twin->p.max_iters > 0))) return RSK_OK;
    if (rsk_pipe_on(s)) s->cur = s->enq_iters & 1;
    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.sc = s->scene->view();
    a.ev = s->em->view();
    a.emit_ids = s->emit_ids; a.tiles = s->tiles; a.n_local = s->n_local; a.tile_rays = s->tile_rays; a.surf_mask = s->mask;
    a.cp_table = s->cp_table; a.rot_base = s->rot_base; a.iters_done = s->iters_done; a.done = s->done;
    // every running job of a solve is at the same iteration: the number of iterations enqueued so far.  A pipelined
    // trace must not read it from the device (the previous iteration's statistics may still be in flight).
    a.iter_index = rsk_pipe_on(s) ? s->enq_iters : -1;
    a.tally = rsk_pipe_tally(s); a.n_hist = s->n_hist; a.ray_begin = s->ray_begin; a.ray_end = s->ray_end; a.min_sid = s->min_sid;
    s->enq_iters++;
    if (s->twin) {
        rsk_solve *k = s->twin;
        a.surf_mask2 = k->mask; a.iters_done2 = k->iters_done; a.done2 = k->done; a.tally2 = rsk_pipe_tally(k); a.n_hist2 = k->n_hist;
        return rsk_launch_trace(ctx, a, MODE_DUAL, s->n_tiles, rsk_pipe_stream(s));
    }
    return rsk_launch_trace(ctx, a, s->mode, s->n_tiles, rsk_pipe_stream(s));
}

static int rsk_solve_enqueue_fold_impl(rsk_solve *s) {
    rsk_ctx *ctx = s->ctx;
    if (s->n_local == 0 || s->p.max_iters <= 0) return RSK_OK;
    const bool pipe = rsk_pipe_on(s);
    cudaStream_t st = rsk_pipe_stream(s);
    // statistics are folded in iteration order: wait for the previous iteration's fold (it ran on the other stream)
    if (pipe && s->has_fold) RSK_CUDA(cudaStreamWaitEvent(st, s->ev_fold, 0));
    FoldArgs f;
    memset(&f, 0, sizeof(f));
    f.iter_tally = rsk_pipe_tally(s); f.total = s->total; f.mean = s->mean; f.m2 = s->m2; f.prev = s->prev;
    f.surf_mask = s->mode == MODE_MATRIX ? s->mask : nullptr;
    f.n_rays_once = s->n_rays_once; f.iters_done = s->iters_done; f.total_rays = s->total_rays; f.done = s->done;
    f.not_converged = s->not_conv; f.n_local = s->n_local; f.n_hist = s->n_hist; f.n_surf = s->scene->n_surf;
    f.mask_words = (s->scene->n_surf + 31) / 32;
    f.max_iters = s->p.max_iters; f.min_iters = s->p.min_iters; f.interval = s->p.interval; f.tol_mode = s->p.tol_mode;
    f.scalar_sky = (s->mode == MODE_SKY && !s->discrete) ? 1 : 0;
    f.tol = s->p.tol;
    DecideArgs d;
    memset(&d, 0, sizeof(d));
    d.iters_done = s->iters_done; d.total_rays = s->total_rays; d.done = s->done; d.not_converged = s->not_conv;
    d.have_prev = s->have_prev; d.n_rays_once = s->n_rays_once; d.n_active = s->n_active; d.rays_traced = s->rays_traced;
    d.ray_begin = s->ray_begin; d.ray_end = s->ray_end;
    d.n_local = s->n_local; d.max_iters = s->p.max_iters; d.min_iters = s->p.min_iters; d.interval = s->p.interval;
    d.tol_mode = s->p.tol_mode;
    RSK_TRY(rsk_launch_fold(ctx, f, st));
    RSK_CUDA(cudaMemsetAsync(s->n_active, 0, sizeof(int32_t), st));
    RSK_TRY(rsk_launch_decide(ctx, d, st));
    if (pipe) {
        RSK_CUDA(cudaEventRecord(s->ev_fold, st));
        s->has_fold = true;
    }
    return RSK_OK;
}

static int rsk_solve_poll_impl(rsk_solve *s, int32_t *n_active) {
    rsk_ctx *ctx = s->ctx;
    if (s->n_local == 0 || s->p.max_iters <= 0 || !s->stepped) {
        if (n_active) *n_active = s->last_active;
        return RSK_OK;
    }
    RSK_TRY(rsk_ctx_join(ctx));                    // order `stream` after the iterations enqueued on the second stream
    RSK_CUDA(cudaMemcpyAsync(s->h_pinned, s->n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    RSK_CUDA(cudaGetLastError());
    s->last_active = s->h_pinned[0];
    if (n_active) *n_active = s->last_active;
    return RSK_OK;
}

static int rsk_solve_step(rsk_solve *s, int32_t n_iters, int32_t *n_active) {
    RSK_REQUIRE(s && n_iters >= 0, "solve step: bad arguments");
    RskScope scope(s->ctx);
    if (s->last_active > 0) {
        for (int it = 0; it < n_iters; ++it) {
            RSK_TRY(rsk_solve_enqueue_trace_impl(s));
            RSK_TRY(rsk_solve_enqueue_fold_impl(s));
            s->stepped = true;
        }
    }
    return rsk_solve_poll_impl(s, n_active);
}

extern "C" int rsk_solve_enqueue_trace(rsk_solve *s) {
    RSK_REQUIRE(s, "null solve");
    RskScope scope(s->ctx);
    return rsk_solve_enqueue_trace_impl(s);
}

extern "C" int rsk_solve_enqueue_fold(rsk_solve *s) {
    RSK_REQUIRE(s, "null solve");
    RskScope scope(s->ctx);
    s->stepped = true;
    return rsk_solve_enqueue_fold_impl(s);
}

extern "C" int rsk_solve_poll(rsk_solve *s, int32_t *n_active) {
    RSK_REQUIRE(s, "null solve");
    RskScope scope(s->ctx);
    return rsk_solve_poll_impl(s, n_active);
}

extern "C" int rsk_solve_set_iter_tally_buffer(rsk_solve *s, void *device_ptr, int64_t n_elements) {
    RSK_REQUIRE(s && device_ptr, "rsk_solve_set_iter_tally_buffer: null argument");
    RSK_REQUIRE(n_elements >= (int64_t)s->n_local * s->n_hist, "rsk_solve_set_iter_tally_buffer: buffer too small");
    RSK_REQUIRE(!s->stepped, "rsk_solve_set_iter_tally_buffer: must be called before the first iteration");
    RskScope scope(s->ctx);
    if (!s->external_tally) rsk_dev_free(s->iter_tally);
    s->iter_tally = (unsigned long long *)device_ptr;
    s->external_tally = true;
    s->pipelined = false;            // one caller-owned buffer: iterations run strictly one after the other
    RSK_CUDA(cudaMemsetAsync(s->iter_tally, 0, (size_t)s->n_local * s->n_hist * 8, s->ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_solve_device_iter_tallies(rsk_solve *s, void **device_ptr, int64_t *n_per_job) {
    RSK_REQUIRE(s && device_ptr && n_per_job, "rsk_solve_device_iter_tallies: bad arguments");
    *device_ptr = rsk_pipe_tally(s);      // the buffer of the iteration enqueued last
    *n_per_job = s->n_hist;
    return RSK_OK;
}

extern "C" int rsk_matrix_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                                const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid,
                                const float *cp_table, int32_t n_rot, const int32_t *rot_base, const int64_t *ray_range,
                                const rsk_solve_params *params, rsk_solve **out) {
    RSK_REQUIRE(n_local == 0 || (emit_sid && min_sid), "rsk_matrix_begin: null emit_sid/min_sid");
    return rsk_solve_begin(ctx, scene, em, MODE_MATRIX, 0, emit_ids, n_local, surf_active, emit_sid, min_sid, cp_table, n_rot, rot_base, ray_range, params, out);
}

extern "C" int rsk_matrix_step(rsk_solve *s, int32_t n_iters, int32_t *n_active) {
    RSK_REQUIRE(s && s->mode == MODE_MATRIX, "rsk_matrix_step: not a matrix solve");
    return rsk_solve_step(s, n_iters, n_active);
}

static int rsk_read_common(rsk_solve *s, int32_t *iters, int64_t *total_rays) {
    RSK_TRY(rsk_ctx_join(s->ctx));         // reads are ordered after the iterations of a pipelined solve on the second stream
    if (iters && s->n_local) RSK_CUDA(cudaMemcpyAsync(iters, s->iters_done, s->n_local * sizeof(int32_t), cudaMemcpyDeviceToHost, s->ctx->stream));
    if (total_rays && s->n_local) RSK_CUDA(cudaMemcpyAsync(total_rays, s->total_rays, s->n_local * sizeof(int64_t), cudaMemcpyDeviceToHost, s->ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_matrix_read(rsk_solve *s, int64_t *hits_front, int64_t *hits_back, int32_t *iters, int64_t *total_rays,
                               double *stderr_front, double *stderr_back) {
    RSK_REQUIRE(s && s->mode == MODE_MATRIX, "rsk_matrix_read: not a matrix solve");
    RskScope scope(s->ctx);
    const int ns = s->scene->n_surf;
    const size_t nh = (size_t)s->n_local * s->n_hist;
    std::vector<int32_t> h_iters(std::max(s->n_local, 1));
    RSK_TRY(rsk_read_common(s, h_iters.data(), total_rays));
    std::vector<long long> tot(nh);
    std::vector<double> m2;
    if (nh) RSK_CUDA(cudaMemcpyAsync(tot.data(), s->total, nh * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    if ((stderr_front || stderr_back) && nh) {
        m2.resize(nh);
        RSK_CUDA(cudaMemcpyAsync(m2.data(), s->m2, nh * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    }
    RSK_CUDA(cudaStreamSynchronize(s->ctx->stream));
    for (int k = 0; k < s->n_local; ++k) {
        if (iters) iters[k] = h_iters[k];
        for (int j = 0; j < ns; ++j) {
            if (hits_front) hits_front[(size_t)k * ns + j] = tot[(size_t)k * s->n_hist + 2 * j];
            if (hits_back) hits_back[(size_t)k * ns + j] = tot[(size_t)k * s->n_hist + 2 * j + 1];
            if (!m2.empty()) {
                const int n = h_iters[k];
                for (int side = 0; side < 2; ++side) {
                    double *dst = side ? stderr_back : stderr_front;
                    if (!dst) continue;
                    const double v = m2[(size_t)k * s->n_hist + 2 * j + side];
                    dst[(size_t)k * ns + j] = n > 1 ? sqrt(std::max(v / (n - 1), 0.0) / n) : INFINITY;   // main.py:1911-1916
                }
            }
        }
    }
    return RSK_OK;
}

extern "C" int rsk_solve_read_block(rsk_solve *s, int64_t *tallies, int32_t *iters, int64_t *total_rays) {
    RSK_REQUIRE(s, "rsk_solve_read_block: null solve");
    RskScope scope(s->ctx);
    RSK_TRY(rsk_read_common(s, iters, total_rays));
    const size_t nh = (size_t)s->n_local * s->n_hist;
    if (tallies && nh) RSK_CUDA(cudaMemcpyAsync(tallies, s->total, nh * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return RSK_OK;
}

// The same through the context's pinned staging area, without a second host copy: *tallies_view points INTO that
// area (valid until the next staged download on this context).
extern "C" int rsk_solve_read_block_view(rsk_solve *s, int64_t **tallies_view, int32_t *iters, int64_t *total_rays) {
    RSK_REQUIRE(s && tallies_view, "rsk_solve_read_block_view: null argument");
    RskScope scope(s->ctx);
    *tallies_view = nullptr;
    RSK_TRY(rsk_read_common(s, iters, total_rays));
    const size_t nh = (size_t)s->n_local * s->n_hist;
    if (nh) {
        void *stage = nullptr;
        RSK_TRY(rsk_ctx_stage(s->ctx, nh * 8, &stage));
        RSK_CUDA(cudaMemcpyAsync(stage, s->total, nh * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
        *tallies_view = (int64_t *)stage;
    }
    RSK_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_matrix_device_tallies(rsk_solve *s, void **device_ptr, int64_t *n_elements) {
    RSK_REQUIRE(s && device_ptr && n_elements, "rsk_matrix_device_tallies: bad arguments");
    *device_ptr = s->total;
    *n_elements = (int64_t)s->n_local * s->n_hist;
    return RSK_OK;
}

// ----------------------------------------------------------------------------- dual (shared-ray) solve
extern "C" int rsk_dual_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                              const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid,
                              const float *cp_table, int32_t n_rot, const int32_t *rot_base,
                              const rsk_solve_params *matrix_params, const rsk_solve_params *sky_params, int32_t discrete,
                              rsk_solve **out) {
    return rsk_dual_begin_sliced(ctx, scene, em, emit_ids, n_local, surf_active, emit_sid, min_sid, cp_table, n_rot, rot_base, nullptr,
                                 matrix_params, sky_params, discrete, out);
}

extern "C" int rsk_dual_begin_sliced(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                                     const uint8_t *surf_active, const int32_t *emit_sid, const int32_t *min_sid,
                                     const float *cp_table, int32_t n_rot, const int32_t *rot_base, const int64_t *ray_range,
                                     const rsk_solve_params *matrix_params, const rsk_solve_params *sky_params, int32_t discrete,
                                     rsk_solve **out) {
    RSK_REQUIRE(out && matrix_params && sky_params, "rsk_dual_begin: null argument");
    RSK_REQUIRE(n_local == 0 || (emit_sid && min_sid), "rsk_dual_begin: null emit_sid/min_sid");
    *out = nullptr;
    rsk_solve *m = nullptr, *k = nullptr;
    RSK_TRY(rsk_solve_begin(ctx, scene, em, MODE_MATRIX, 0, emit_ids, n_local, surf_active, emit_sid, min_sid, cp_table, n_rot, rot_base,
                            ray_range, matrix_params, &m));
    int rc = rsk_solve_begin(ctx, scene, em, MODE_SKY, discrete ? 1 : 0, emit_ids, n_local, surf_active, nullptr, nullptr, cp_table, n_rot,
                             rot_base, ray_range, sky_params, &k);
    if (rc != RSK_OK) { rsk_solve_destroy(m); return rc; }
    m->twin = k;
    k->primary = m;
    // an emitter without receivers never starts its matrix side (main.py:1285-1287): mark it done up front
    RskScope scope(ctx);
    std::vector<int32_t> done(std::max(n_local, 1), 0);
    int n_running = 0;
    for (int j = 0; j < n_local; ++j) {
        const int es = emit_sid[j], ms = min_sid[j];
        bool any = false;
        const uint8_t *row = surf_active + (size_t)j * scene->n_surf;
        for (int s = std::max(ms, 0); s < scene->n_surf && !any; ++s) any = row[s] != 0 && s != es;
        done[j] = (any && matrix_params->max_iters > 0) ? 0 : 1;
        n_running += done[j] ? 0 : 1;
    }
    if (n_local > 0) {
        cudaError_t e = cudaMemcpyAsync(m->done, done.data(), n_local * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rsk_set_error("rsk_dual_begin: %s", cudaGetErrorString(e)); rsk_solve_destroy(m); return RSK_ERR_CUDA; }
    }
    m->last_active = n_running;
    *out = m;
    return RSK_OK;
}

extern "C" int rsk_dual_step(rsk_solve *s, int32_t n_iters, int32_t *n_active) {
    RSK_REQUIRE(s && s->twin && n_iters >= 0, "rsk_dual_step: not a dual solve");
    rsk_solve *k = s->twin;
    RskScope scope(s->ctx);
    if (s->last_active + k->last_active > 0) {
        for (int it = 0; it < n_iters; ++it) {
            RSK_TRY(rsk_solve_enqueue_trace_impl(s));
            RSK_TRY(rsk_solve_enqueue_fold_impl(s));
            RSK_TRY(rsk_solve_enqueue_fold_impl(k));
            s->stepped = k->stepped = true;
        }
    }
    int32_t am = 0, ak = 0;
    RSK_TRY(rsk_solve_poll_impl(s, &am));
    RSK_TRY(rsk_solve_poll_impl(k, &ak));
    if (n_active) *n_active = am + ak;
    return RSK_OK;
}

extern "C" int rsk_dual_sky_part(rsk_solve *s, rsk_solve **sky) {
    RSK_REQUIRE(s && s->twin && sky, "rsk_dual_sky_part: not a dual solve");
    *sky = s->twin;
    return RSK_OK;
}

extern "C" int rsk_sky_begin(rsk_ctx *ctx, rsk_scene *scene, rsk_emitters *em, const int32_t *emit_ids, int32_t n_local,
                             const uint8_t *surf_active, const float *cp_table, int32_t n_rot, const int32_t *rot_base,
                             const int64_t *ray_range, const rsk_solve_params *params, int32_t discrete, rsk_solve **out) {
    // main.py:2112: emit_sid = emitter index, min_sid = 0
    return rsk_solve_begin(ctx, scene, em, MODE_SKY, discrete ? 1 : 0, emit_ids, n_local, surf_active, nullptr, nullptr, cp_table, n_rot, rot_base, ray_range, params, out);
}

extern "C" int rsk_sky_step(rsk_solve *s, int32_t n_iters, int32_t *n_active) {
    RSK_REQUIRE(s && s->mode == MODE_SKY, "rsk_sky_step: not a sky solve");
    return rsk_solve_step(s, n_iters, n_active);
}

extern "C" int rsk_sky_read(rsk_solve *s, int64_t *counts, int32_t *iters, int64_t *total_rays) {
    RSK_REQUIRE(s && s->mode == MODE_SKY, "rsk_sky_read: not a sky solve");
    RskScope scope(s->ctx);
    RSK_TRY(rsk_read_common(s, iters, total_rays));
    const size_t nh = (size_t)s->n_local * s->n_hist;
    if (counts && nh) RSK_CUDA(cudaMemcpyAsync(counts, s->total, nh * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_solve_rays_traced(rsk_solve *s, int64_t *rays) {
    RSK_REQUIRE(s && rays, "rsk_solve_rays_traced: bad arguments");
    RskScope scope(s->ctx);
    RSK_TRY(rsk_ctx_join(s->ctx));
    unsigned long long v = 0;
    RSK_CUDA(cudaMemcpyAsync(&v, s->rays_traced, 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(s->ctx->stream));
    *rays = (int64_t)v;
    return RSK_OK;
}

extern "C" int rsk_solve_destroy(rsk_solve *s) {
    if (!s) return RSK_OK;
    if (s->twin) { rsk_solve_destroy(s->twin); s->twin = nullptr; }
    RskScope scope(s->ctx);
    rsk_ctx_join(s->ctx);             // frees are ordered on `stream`: after the iterations still running on the second one
    rsk_dev_free(s->iter_tally2);
    if (s->ev_fold) cudaEventDestroy(s->ev_fold);
    rsk_dev_free(s->emit_ids); rsk_dev_free(s->min_sid); rsk_dev_free(s->rot_base); rsk_dev_free(s->iters_done); rsk_dev_free(s->done); rsk_dev_free(s->not_conv);
    rsk_dev_free(s->have_prev); rsk_dev_free(s->tiles); rsk_dev_free(s->n_rays_once); rsk_dev_free(s->total_rays); rsk_dev_free(s->ray_begin); rsk_dev_free(s->ray_end); rsk_dev_free(s->mask);
    rsk_dev_free(s->cp_table); if (!s->external_tally) rsk_dev_free(s->iter_tally); rsk_dev_free(s->rays_traced); rsk_dev_free(s->total); rsk_dev_free(s->mean);
    rsk_dev_free(s->m2); rsk_dev_free(s->prev); rsk_dev_free(s->n_active);
    delete s;
    return RSK_OK;
}
