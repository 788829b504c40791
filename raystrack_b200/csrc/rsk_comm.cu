// rsk_comm.cu -- the multi-GPU half of the C ABI: one NCCL communicator per context, int64 all-reduces on the
// context's stream, and the device-side tally block that the ranks of a sharded solve sum over NVLink.
//
// The reference has no multi-device code; this is the exchange step SURVEY.md 8(b)/(e) specifies:
// emitters are sharded over the GPUs, the integer tallies [n_emit][n_hist] are summed once at the end (and, for
// emitters whose rays are split over ranks, the per-iteration tallies before every statistics update).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded into the process by PyTorch if there is
// one, else $RSK_NCCL_LIBRARY, else the system library), so librsk_b200.so itself loads on machines without NCCL
// and single-GPU callers never touch it.  Only the handful of NCCL entry points below are used; their signatures
// and enum values are those of NCCL 2.x (nccl.h).
#include <dlfcn.h>

#include "rsk_solve.cuh"

namespace {

typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId_ { char internal[128]; };
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT64 = 4 };                        // ncclDataType_t
enum { NCCL_SUM = 0, NCCL_MAX = 2 };            // ncclRedOp_t

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

int rsk_nccl_load() {
    if (g_nccl.handle) return RSK_OK;
    void *h = nullptr;
    const char *env = getenv("RSK_NCCL_LIBRARY");
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // already in the process (PyTorch's copy)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        rsk_set_error("NCCL not found (dlopen libnccl.so.2: %s); set RSK_NCCL_LIBRARY", dlerror());
        return RSK_ERR_INVALID;
    }
    NcclApi api;
    api.handle = h;
#define RSK_SYM(field, name)                                                                              \
    *(void **)(&api.field) = dlsym(h, name);                                                              \
    if (!api.field) { rsk_set_error("NCCL library lacks %s", name); return RSK_ERR_INVALID; }
    RSK_SYM(GetUniqueId, "ncclGetUniqueId")
    RSK_SYM(CommInitRank, "ncclCommInitRank")
    RSK_SYM(CommDestroy, "ncclCommDestroy")
    RSK_SYM(AllReduce, "ncclAllReduce")
    RSK_SYM(GetErrorString, "ncclGetErrorString")
    RSK_SYM(GetVersion, "ncclGetVersion")
#undef RSK_SYM
    g_nccl = api;
    return RSK_OK;
}

#define RSK_NCCL(call)                                                                                    \
    do {                                                                                                  \
        int r__ = (call);                                                                                 \
        if (r__ != NCCL_SUCCESS) {                                                                        \
            rsk_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
            return RSK_ERR_CUDA;                                                                          \
        }                                                                                                 \
    } while (0)

}  // namespace

extern "C" int rsk_comm_unique_id(uint8_t *id) {
    RSK_REQUIRE(id, "rsk_comm_unique_id: null output");
    RSK_TRY(rsk_nccl_load());
    ncclUniqueId_ u;
    RSK_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, RSK_COMM_ID_BYTES);
    return RSK_OK;
}

extern "C" int rsk_comm_init(rsk_ctx *ctx, const uint8_t *id, int32_t rank, int32_t nranks) {
    RSK_REQUIRE(ctx && id, "rsk_comm_init: null argument");
    RSK_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rsk_comm_init: rank out of range");
    RSK_REQUIRE(!ctx->comm, "rsk_comm_init: context already has a communicator");
    RSK_TRY(rsk_nccl_load());
    RskScope scope(ctx);
    ncclUniqueId_ u;
    memcpy(u.internal, id, RSK_COMM_ID_BYTES);
    ncclComm_t comm = nullptr;
    RSK_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    if (!ctx->comm_scratch) RSK_CUDA(cudaMalloc((void **)&ctx->comm_scratch, RSK_COMM_SCRATCH * sizeof(long long)));
    return RSK_OK;
}

extern "C" int rsk_comm_destroy(rsk_ctx *ctx) {
    RSK_REQUIRE(ctx, "rsk_comm_destroy: null context");
    if (!ctx->comm) return RSK_OK;
    RskScope scope(ctx);
    cudaStreamSynchronize(ctx->stream);
    RSK_NCCL(g_nccl.CommDestroy((ncclComm_t)ctx->comm));
    ctx->comm = nullptr;
    ctx->comm_rank = 0;
    ctx->comm_size = 1;
    if (ctx->comm_scratch) { cudaFree(ctx->comm_scratch); ctx->comm_scratch = nullptr; }
    return RSK_OK;
}

extern "C" int rsk_comm_info(rsk_ctx *ctx, int32_t *rank, int32_t *nranks, int32_t *nccl_version) {
    RSK_REQUIRE(ctx, "rsk_comm_info: null context");
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (nranks) *nranks = ctx->comm ? ctx->comm_size : 1;
    if (nccl_version) {
        *nccl_version = 0;
        if (g_nccl.handle) g_nccl.GetVersion(nccl_version);
    }
    return RSK_OK;
}

static int rsk_allreduce_impl(rsk_ctx *ctx, void *device_ptr, int64_t n, int32_t op, cudaStream_t stream = nullptr) {
    RSK_REQUIRE(op == 0 || op == 1, "all-reduce: op must be 0 (sum) or 1 (max)");
    if (n <= 0 || !ctx->comm || ctx->comm_size <= 1) return RSK_OK;
    RSK_NCCL(g_nccl.AllReduce(device_ptr, device_ptr, (size_t)n, NCCL_INT64, op == 0 ? NCCL_SUM : NCCL_MAX, (ncclComm_t)ctx->comm,
                              stream ? stream : ctx->stream));
    return RSK_OK;
}

extern "C" int rsk_allreduce_i64(rsk_ctx *ctx, void *device_ptr, int64_t n, int32_t op) {
    RSK_REQUIRE(ctx && (device_ptr || n == 0), "rsk_allreduce_i64: null argument");
    RskScope scope(ctx);
    return rsk_allreduce_impl(ctx, device_ptr, n, op);
}

extern "C" int rsk_allreduce_host_i64(rsk_ctx *ctx, int64_t *values, int64_t n, int32_t op) {
    RSK_REQUIRE(ctx && values && n >= 0 && n <= RSK_COMM_SCRATCH, "rsk_allreduce_host_i64: bad arguments (n <= 4096)");
    if (n == 0 || !ctx->comm || ctx->comm_size <= 1) return RSK_OK;
    RskScope scope(ctx);
    RSK_CUDA(cudaMemcpyAsync(ctx->comm_scratch, values, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    RSK_TRY(rsk_allreduce_impl(ctx, ctx->comm_scratch, n, op));
    RSK_CUDA(cudaMemcpyAsync(values, ctx->comm_scratch, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return RSK_OK;
}

// Split-phase stepping without a caller-side collective: sum the iteration tallies of the first n_jobs jobs (the
// ray-split ones, which come first on every rank) over the communicator, on the context stream, between
// rsk_solve_enqueue_trace and rsk_solve_enqueue_fold.
extern "C" int rsk_solve_allreduce_iter_tallies(rsk_solve *s, int32_t n_jobs) {
    RSK_REQUIRE(s && n_jobs >= 0 && n_jobs <= s->n_local, "rsk_solve_allreduce_iter_tallies: bad arguments");
    RskScope scope(s->ctx);
    // a pipelined solve traced this iteration on one of two streams into one of two buffers: reduce there
    void *buf = nullptr;
    int64_t per_job = 0;
    RSK_TRY(rsk_solve_device_iter_tallies(s, &buf, &per_job));
    return rsk_allreduce_impl(s->ctx, buf, (int64_t)n_jobs * per_job, 0, rsk_solve_current_stream(s));
}

// ----------------------------------------------------------------------------- tally block


// full[rows[k]][:] = part[k][:] for the kept jobs k (rows < 0 = skip): one CTA column-strides over one row
__global__ void rsk_scatter_rows_kernel(const long long *__restrict__ part, const int32_t *__restrict__ rows, long long *__restrict__ full,
                                        int64_t n_cols) {
    const int32_t r = rows[blockIdx.x];
    if (r < 0) return;
    const long long *src = part + (int64_t)blockIdx.x * n_cols;
    long long *dst = full + (int64_t)r * n_cols;
    for (int64_t c = threadIdx.x; c < n_cols; c += blockDim.x) dst[c] = src[c];
}

extern "C" int rsk_tally_block_create(rsk_ctx *ctx, int64_t n_rows, int64_t n_cols, rsk_tally_block **out) {
    RSK_REQUIRE(ctx && out && n_rows >= 0 && n_cols >= 0, "rsk_tally_block_create: bad arguments");
    *out = nullptr;
    RskScope scope(ctx);
    rsk_tally_block *b = new rsk_tally_block();
    b->ctx = ctx; b->n_rows = n_rows; b->n_cols = n_cols;
    const size_t n = (size_t)std::max<int64_t>(n_rows * n_cols, 1);
    int rc = rsk_dev_alloc(&b->d, n);
    if (rc == RSK_OK && cudaMemsetAsync(b->d, 0, n * 8, ctx->stream) != cudaSuccess) { rsk_set_error("tally block: memset failed"); rc = RSK_ERR_CUDA; }
    if (rc != RSK_OK) { rsk_dev_free(b->d); delete b; return rc; }
    *out = b;
    return RSK_OK;
}

extern "C" int rsk_tally_block_destroy(rsk_tally_block *b) {
    if (!b) return RSK_OK;
    RskScope scope(b->ctx);
    rsk_dev_free(b->d);
    delete b;
    return RSK_OK;
}

extern "C" int rsk_tally_block_add_solve(rsk_tally_block *b, rsk_solve *s, const uint8_t *keep) {
    RSK_REQUIRE(b && s && b->ctx == s->ctx, "rsk_tally_block_add_solve: bad arguments");
    RSK_REQUIRE(s->n_hist == b->n_cols, "rsk_tally_block_add_solve: column count mismatch");
    if (s->n_local == 0) return RSK_OK;
    rsk_ctx *ctx = b->ctx;
    RskScope scope(ctx);
    RSK_TRY(rsk_ctx_join(ctx));
    // destination rows = the solve's emitter ids (already on the device); jobs not kept get row -1
    std::vector<int32_t> rows(s->n_local);
    RSK_CUDA(cudaMemcpyAsync(rows.data(), s->emit_ids, s->n_local * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < s->n_local; ++k) {
        if (keep && !keep[k]) rows[k] = -1;
        RSK_REQUIRE(rows[k] < b->n_rows, "rsk_tally_block_add_solve: emitter id beyond the block");
    }
    int32_t *d_rows = nullptr;
    RSK_TRY(rsk_dev_alloc(&d_rows, rows.size()));
    cudaError_t e = cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        rsk_scatter_rows_kernel<<<(unsigned)s->n_local, 256, 0, ctx->stream>>>(s->total, d_rows, b->d, b->n_cols);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);      // `rows` is pageable host memory
    rsk_dev_free(d_rows);
    if (e != cudaSuccess) { rsk_set_error("rsk_tally_block_add_solve: %s", cudaGetErrorString(e)); return RSK_ERR_CUDA; }
    return RSK_OK;
}

extern "C" int rsk_tally_block_allreduce(rsk_tally_block *b) {
    RSK_REQUIRE(b, "rsk_tally_block_allreduce: null block");
    RskScope scope(b->ctx);
    return rsk_allreduce_impl(b->ctx, b->d, b->n_rows * b->n_cols, 0);
}

extern "C" int rsk_tally_block_device(rsk_tally_block *b, void **device_ptr, int64_t *n_elements) {
    RSK_REQUIRE(b && device_ptr, "rsk_tally_block_device: bad arguments");
    *device_ptr = b->d;
    if (n_elements) *n_elements = b->n_rows * b->n_cols;
    return RSK_OK;
}

// Download through the context's pinned staging area (grow-only, allocated once: cudaMallocHost/cudaFreeHost
// synchronise the device).  `view` receives a pointer INTO that area: valid until the next staged download on the
// same context.  dst (optional) additionally receives a copy in caller memory.
int rsk_ctx_stage(rsk_ctx *ctx, size_t bytes, void **stage) {
    if (ctx->stage_cap < bytes) {
        if (ctx->stage) cudaFreeHost(ctx->stage);
        ctx->stage = nullptr;
        ctx->stage_cap = 0;
        RSK_CUDA(cudaMallocHost(&ctx->stage, bytes));
        ctx->stage_cap = bytes;
    }
    *stage = ctx->stage;
    return RSK_OK;
}

extern "C" int rsk_tally_block_download(rsk_tally_block *b, int64_t *dst, int64_t **view) {
    RSK_REQUIRE(b && (dst || view), "rsk_tally_block_download: bad arguments");
    rsk_ctx *ctx = b->ctx;
    RskScope scope(ctx);
    const size_t bytes = (size_t)(b->n_rows * b->n_cols) * 8;
    if (view) *view = nullptr;
    if (bytes == 0) return RSK_OK;
    void *stage = nullptr;
    RSK_TRY(rsk_ctx_stage(ctx, bytes, &stage));
    RSK_CUDA(cudaMemcpyAsync(stage, b->d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (dst) memcpy(dst, stage, bytes);
    if (view) *view = (int64_t *)stage;
    return RSK_OK;
}
