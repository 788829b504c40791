// rsk_qmc.cu -- device generation of the reference's QMC tables (utils/halton.py:9-58).
//
// The reference evaluates the radical inverse in float64 with the recurrence  f /= base; r += f * digit
// and rounds the result to float32.  The kernels below run the same recurrence with IEEE division and
// explicit non-fused multiply/add, so the float32 tables are bit-identical to the reference's.
#include "rsk_common.cuh"

__device__ __forceinline__ double rsk_halton_f64(int64_t i, int base) {
    double f = 1.0, r = 0.0;
    const double b = (double)base;
    while (i) {
        f = __ddiv_rn(f, b);
        r = __dadd_rn(r, __dmul_rn(f, (double)(i % base)));
        i /= base;
    }
    return r;
}

// out[row][k] = float32(H_base(k+1)) for k in [first, n); rows: bases 5,2,3,7,11 (halton.py:52-58).
__global__ void rsk_halton_dims_kernel(float *out, int64_t stride, int64_t first, int64_t n) {
    const int bases[5] = {5, 2, 3, 7, 11};
    int64_t k = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
#pragma unroll
    for (int r = 0; r < 5; ++r) out[r * stride + k] = (float)rsk_halton_f64(k + 1, bases[r]);
}

// halton.py:21-31: u[c] = (H2(c+1) + c//g)/g, v[c] = (H3(c+1) + c%g)/g, stored interleaved.
__global__ void rsk_halton_grid_kernel(float2 *out, int g) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g * g) return;
    int i = c / g, j = c % g;
    double u = __ddiv_rn(__dadd_rn(rsk_halton_f64(c + 1, 2), (double)i), (double)g);
    double v = __ddiv_rn(__dadd_rn(rsk_halton_f64(c + 1, 3), (double)j), (double)g);
    out[c] = make_float2((float)u, (float)v);
}

int rsk_qmc_ensure_halton(rsk_ctx *ctx, int64_t n) {
    if (n <= ctx->halton_cap) return RSK_OK;
    // grow geometrically; tables are prefixes of one another, so the old part is regenerated in place
    int64_t cap = ctx->halton_cap > 0 ? ctx->halton_cap : 65536;
    while (cap < n) cap *= 2;
    if (cap > n && cap > (int64_t)1 << 26) cap = (n + 1023) / 1024 * 1024;   // do not double huge tables
    float *fresh = nullptr;
    RSK_TRY(rsk_dev_alloc(&fresh, (size_t)cap * 5));
    rsk_halton_dims_kernel<<<rsk_blocks(cap, 256), 256, 0, ctx->stream>>>(fresh, cap, 0, cap);
    ctx->launches++;
    RSK_CUDA(cudaGetLastError());
    RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->halton) rsk_dev_free(ctx->halton);
    ctx->halton = fresh;
    ctx->halton_cap = cap;
    return RSK_OK;
}

int rsk_qmc_ensure_grid(rsk_ctx *ctx, int g, int64_t *offset) {
    auto it = ctx->grid_index.find(g);
    if (it != ctx->grid_index.end()) {
        *offset = it->second.first;
        return RSK_OK;
    }
    int64_t cells = (int64_t)g * g;
    if (ctx->grid_used + cells > ctx->grid_cap) {
        int64_t cap = ctx->grid_cap > 0 ? ctx->grid_cap : 65536;
        while (cap < ctx->grid_used + cells) cap *= 2;
        float2 *fresh = nullptr;
        RSK_TRY(rsk_dev_alloc(&fresh, (size_t)cap));
        if (ctx->grid_used > 0)
            RSK_CUDA(cudaMemcpyAsync(fresh, ctx->grid, ctx->grid_used * sizeof(float2), cudaMemcpyDeviceToDevice, ctx->stream));
        RSK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->grid) rsk_dev_free(ctx->grid);
        ctx->grid = fresh;
        ctx->grid_cap = cap;
    }
    rsk_halton_grid_kernel<<<rsk_blocks(cells, 256), 256, 0, ctx->stream>>>(ctx->grid + ctx->grid_used, g);
    ctx->launches++;
    RSK_CUDA(cudaGetLastError());
    ctx->grid_index[g] = std::make_pair(ctx->grid_used, cells);
    *offset = ctx->grid_used;
    ctx->grid_used += cells;
    return RSK_OK;
}
