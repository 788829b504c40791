// rsk_stats.cu -- on-device iteration statistics, convergence decisions and the reciprocity/row-sum solver.
//
// Replaces the host-side bookkeeping of the reference's iteration loops: `hits += hits_iter`, the Welford update
// of the per-iteration fractions, `_convergence_checkpoint` and the stderr/delta stop rules
// (main.py:217-228, 1872-1909 for the matrix; main.py:2122-2174 for the sky), plus kernel_accumulate_hits[_stats]
// (utils/cuda_trace.py:581-616).  All float64 arithmetic uses explicit round-to-nearest intrinsics in the
// reference's NumPy operation order (no FMA contraction), so decisions are bit-identical to the CPU reference
// whenever the integer tallies are.
#include "rsk_solve.cuh"


// main.py:217-228 `_convergence_checkpoint`.
__host__ __device__ inline bool rsk_checkpoint(int iters_done, int min_iters, int interval, int max_iters, bool needs_variance) {
    const int start = min_iters > 1 ? min_iters : 1;
    if (iters_done < start) return false;
    if (needs_variance && iters_done <= 1) return false;
    if (iters_done >= max_iters) return true;
    const int span = interval > 1 ? interval : 1;
    if (span <= 1) return true;
    return ((iters_done - start) % span) == 0;
}

__global__ void rsk_fold_kernel(const FoldArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.n_local * a.n_hist) return;
    const int k = (int)(idx / a.n_hist), j = (int)(idx % a.n_hist);
    if (a.done[k]) {
        // a pipelined solve may have traced this job once more before its stop decision was known: drop those tallies
        if (a.iter_tally[idx] != 0ull) a.iter_tally[idx] = 0ull;
        return;
    }
    const unsigned long long cnt = a.iter_tally[idx];
    const long long tot0 = a.total[idx];
    // A bin that has never been hit stays all-zero (x = 0, mean = M2 = 0, standard error 0, cumulative estimate 0) and
    // passes either stop rule whenever the tolerance admits a zero error: nothing to read, update or write.  Nine bins
    // in ten of a city-sized matrix are of this kind.
    if (cnt == 0ull && tot0 == 0 && (a.tol_mode == 0 ? a.tol >= 0.0 : a.tol > 0.0)) return;
    a.iter_tally[idx] = 0ull;
    const long long tot = tot0 + (long long)cnt;
    a.total[idx] = tot;

    const int n = a.iters_done[k] + 1;
    const double n_once = (double)a.n_rays_once[k];
    // main.py:1877-1884 / 2133-2140: x = hits_iter / n_rays_once; Welford mean/M2
    const double x = __ddiv_rn((double)cnt, n_once);
    double mean = a.mean[idx], m2 = a.m2[idx];
    const double delta = __dsub_rn(x, mean);
    mean = __dadd_rn(mean, __ddiv_rn(delta, (double)n));
    m2 = __dadd_rn(m2, __dmul_rn(delta, __dsub_rn(x, mean)));
    a.mean[idx] = mean;
    a.m2[idx] = m2;

    const bool check = rsk_checkpoint(n, a.min_iters, a.interval, a.max_iters, a.tol_mode == 0);
    if (!check) return;
    if (a.tol_mode == 0) {
        // main.py:1904-1906: only active receivers are tested (front and back); sky: every bin
        if (a.surf_mask) {
            const int s = j >> 1;                       // matrix bins are (receiver, side) pairs
            if (!((a.surf_mask[(int64_t)k * a.mask_words + (s >> 5)] >> (s & 31)) & 1u)) return;
        }
        double se;
        if (a.scalar_sky) se = __ddiv_rn(sqrt(fmax(__ddiv_rn(m2, (double)(n - 1)), 0.0)), sqrt((double)n));
        else se = sqrt(__ddiv_rn(fmax(__ddiv_rn(m2, (double)(n - 1)), 0.0), (double)n));
        if (!(se <= a.tol)) a.not_converged[k] = 1;
    } else {
        // main.py:1893-1901: cumulative estimate against the previous checkpoint, all surfaces
        const double total_rays = (double)(a.total_rays[k] + a.n_rays_once[k]);
        const double cur = __ddiv_rn((double)tot, total_rays);
        const double prev = a.prev[idx];
        if (!(fabs(__dsub_rn(cur, prev)) < a.tol)) a.not_converged[k] = 1;
        a.prev[idx] = cur;
    }
}


__global__ void rsk_decide_kernel(const DecideArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_local) return;
    if (a.done[k]) return;
    const int n = a.iters_done[k] + 1;
    a.iters_done[k] = n;
    a.total_rays[k] += a.n_rays_once[k];
    atomicAdd(a.rays_traced, (unsigned long long)(a.ray_end[k] - a.ray_begin[k]));
    const bool check = rsk_checkpoint(n, a.min_iters, a.interval, a.max_iters, a.tol_mode == 0);
    bool converged = false;
    if (check) {
        converged = a.not_converged[k] == 0;
        if (a.tol_mode == 1) {
            if (!a.have_prev[k]) converged = false;     // `prev_f is not None` (main.py:1896)
            a.have_prev[k] = 1;
        }
    }
    a.not_converged[k] = 0;
    const bool stop = converged || n >= a.max_iters;
    a.done[k] = stop ? 1 : 0;
    if (!stop) atomicAdd(a.n_active, 1);
}

int rsk_launch_fold(rsk_ctx *ctx, const FoldArgs &a, cudaStream_t stream) {
    const int64_t n = (int64_t)a.n_local * a.n_hist;
    if (n == 0) return RSK_OK;
    rsk_fold_kernel<<<rsk_blocks(n, 256), 256, 0, stream>>>(a);
    ctx->launches++;
    RSK_CUDA(cudaGetLastError());
    return RSK_OK;
}

int rsk_launch_decide(rsk_ctx *ctx, const DecideArgs &a, cudaStream_t stream) {
    if (a.n_local == 0) return RSK_OK;
    rsk_decide_kernel<<<rsk_blocks(a.n_local, 128), 128, 0, stream>>>(a);
    ctx->launches++;
    RSK_CUDA(cudaGetLastError());
    return RSK_OK;
}

// ----------------------------------------------------------------------------- reciprocity + row sums
// utils/helpers.py:70-96.  G = 0.5*(A_i F_ij + A_j F_ji); d <- d * sqrt(max(target/max(d*(G d),1e-30),0)).

__global__ void rsk_recip_symmetrize(const double *F, const double *area, double *G, int n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    G[idx] = 0.5 * (area[i] * F[idx] + area[j] * F[(int64_t)j * n + i]);
}

// one warp per row: row_i = d_i * sum_j G_ij d_j ; d_new_i = d_i * sqrt(max(target_i / max(row_i,1e-30), 0))
__global__ void rsk_recip_sweep(const double *G, const double *target, const double *d, double *d_new, double *max_delta_bits, int n) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double *row = G + (int64_t)warp * n;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc += row[j] * d[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const double di = d[warp];
        const double r = fmax(di * acc, 1e-30);
        const double upd = fmax(target[warp] / r, 0.0);
        const double dn = di * sqrt(upd);
        d_new[warp] = dn;
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(max_delta_bits), (unsigned long long)__double_as_longlong(fabs(dn - di)));
    }
}

__global__ void rsk_recip_apply(const double *G, const double *area, const double *d, double *F, int n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    const double gp = (d[i] * G[idx]) * d[j];
    F[idx] = area[i] > 0.0 ? gp / area[i] : 0.0;
}

extern "C" int rsk_reciprocity_rowsum(rsk_ctx *ctx, int32_t n, const double *area, const double *target, double *F,
                                      double tol, int32_t max_iter, int32_t *sweeps) {
    RSK_REQUIRE(ctx && area && F && n >= 0, "rsk_reciprocity_rowsum: bad arguments");
    if (sweeps) *sweeps = 0;
    if (n == 0) return RSK_OK;
    RskScope scope(ctx);
    const int64_t nn = (int64_t)n * n;
    double *dF = nullptr, *dG = nullptr, *dA = nullptr, *dT = nullptr, *dd = nullptr, *dd2 = nullptr, *dmax = nullptr;
    int rc = RSK_OK;
    auto cleanup = [&]() { rsk_dev_free(dF); rsk_dev_free(dG); rsk_dev_free(dA); rsk_dev_free(dT); rsk_dev_free(dd); rsk_dev_free(dd2); rsk_dev_free(dmax); };
#define RSK_R(expr) do { rc = (expr); if (rc != RSK_OK) { cleanup(); return rc; } } while (0)
#define RSK_RC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rsk_set_error("%s: %s", #call, cudaGetErrorString(e__)); cleanup(); return RSK_ERR_CUDA; } } while (0)
    RSK_R(rsk_dev_alloc(&dF, nn)); RSK_R(rsk_dev_alloc(&dG, nn)); RSK_R(rsk_dev_alloc(&dA, n)); RSK_R(rsk_dev_alloc(&dT, n));
    RSK_R(rsk_dev_alloc(&dd, n)); RSK_R(rsk_dev_alloc(&dd2, n)); RSK_R(rsk_dev_alloc(&dmax, 1));
    std::vector<double> ones(n, 1.0), tgt(n);
    for (int i = 0; i < n; ++i) tgt[i] = target ? area[i] * (target[i] > 0.0 ? target[i] : 0.0) : area[i];   // helpers.py:50-58
    cudaStream_t s = ctx->stream;
    RSK_RC(cudaMemcpyAsync(dF, F, nn * sizeof(double), cudaMemcpyHostToDevice, s));
    RSK_RC(cudaMemcpyAsync(dA, area, n * sizeof(double), cudaMemcpyHostToDevice, s));
    RSK_RC(cudaMemcpyAsync(dT, tgt.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
    RSK_RC(cudaMemcpyAsync(dd, ones.data(), n * sizeof(double), cudaMemcpyHostToDevice, s));
    rsk_recip_symmetrize<<<rsk_blocks(nn, 256), 256, 0, s>>>(dF, dA, dG, n);
    ctx->launches++;
    int it = 0;
    for (; it < max_iter; ++it) {
        RSK_RC(cudaMemsetAsync(dmax, 0, sizeof(double), s));
        rsk_recip_sweep<<<rsk_blocks((int64_t)n * 32, 256), 256, 0, s>>>(dG, dT, dd, dd2, dmax, n);
        ctx->launches++;
        double h_max = 0.0;
        RSK_RC(cudaMemcpyAsync(&h_max, dmax, sizeof(double), cudaMemcpyDeviceToHost, s));
        RSK_RC(cudaStreamSynchronize(s));
        std::swap(dd, dd2);
        if (h_max < tol) { ++it; break; }
    }
    rsk_recip_apply<<<rsk_blocks(nn, 256), 256, 0, s>>>(dG, dA, dd, dF, n);
    ctx->launches++;
    RSK_RC(cudaMemcpyAsync(F, dF, nn * sizeof(double), cudaMemcpyDeviceToHost, s));
    RSK_RC(cudaStreamSynchronize(s));
    RSK_RC(cudaGetLastError());
    if (sweeps) *sweeps = it;
    cleanup();
#undef RSK_R
#undef RSK_RC
    return RSK_OK;
}


// ----------------------------------------------------------------------------- result rows in compressed form
// Replaces, for city-sized matrices, the dense read-back of the tallies followed by the host loop of main.py:1918-1934
// (`F = hits / total_rays`, keys only for F > 0): nine bins in ten are zero, so the device emits per row the column
// indices and values of the non-zero bins (CSR, columns ascending = the reference's key order) and only those cross
// PCIe.  The division is the reference's: int64 -> float64 conversions (exact below 2^53), one IEEE division.

__global__ void rsk_csr_count_kernel(const long long *__restrict__ block, int64_t n_cols, int32_t *__restrict__ row_count) {
    const long long *row = block + (int64_t)blockIdx.x * n_cols;
    int c = 0;
    for (int64_t j = threadIdx.x; j < n_cols; j += blockDim.x) c += row[j] != 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int s_part[32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        row_count[blockIdx.x] = t;
    }
}

// exclusive prefix sum of n counts into row_ptr[0..n] (one CTA; n is the number of meshes)
__global__ void rsk_csr_scan_kernel(const int32_t *__restrict__ row_count, int64_t n, long long *__restrict__ row_ptr) {
    __shared__ long long s_carry;
    __shared__ long long s_warp[32];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < n ? row_count[i] : 0;
        long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        long long before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (i < n) row_ptr[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = s_carry;
            for (int w = 0; w < nw; ++w) t += s_warp[w];
            s_carry = t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) row_ptr[n] = s_carry;
}

__global__ void rsk_csr_fill_kernel(const long long *__restrict__ block, int64_t n_cols, const long long *__restrict__ total_rays,
                                    const long long *__restrict__ row_ptr, int32_t *__restrict__ cols, double *__restrict__ vals) {
    const long long *row = block + (int64_t)blockIdx.x * n_cols;
    const double denom = (double)total_rays[blockIdx.x];
    __shared__ int s_warp[32];
    __shared__ long long s_base;
    if (threadIdx.x == 0) s_base = row_ptr[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t base = 0; base < n_cols; base += blockDim.x) {
        const int64_t j = base + threadIdx.x;
        const long long t = j < n_cols ? row[j] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, t != 0);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
        for (int w = 0; w < nw; ++w) { const int c = s_warp[w]; if (w < warp) before += c; all += c; }
        if (t != 0) {
            const long long dst = s_base + before + __popc(m & ((1u << lane) - 1u));
            cols[dst] = (int32_t)j;
            vals[dst] = __ddiv_rn((double)t, denom);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += all;
        __syncthreads();
    }
}

static int rsk_csr_build_impl(rsk_ctx *ctx, const long long *block, int64_t n_rows, int64_t n_cols, const long long *d_total_rays,
                              const int64_t *h_total_rays, int64_t *row_ptr) {
    RSK_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_cols < (1ll << 31) && n_rows < (1ll << 31), "rsk_csr_build: bad shape");
    cudaStream_t st = ctx->stream;
    rsk_dev_free(ctx->csr_cols); rsk_dev_free(ctx->csr_vals);
    ctx->csr_cols = nullptr; ctx->csr_vals = nullptr; ctx->csr_nnz = 0;
    if (n_rows == 0) { row_ptr[0] = 0; return RSK_OK; }
    int32_t *d_count = nullptr;
    long long *d_ptr = nullptr, *d_tot = nullptr;
    int rc = RSK_OK;
    auto cleanup = [&]() { rsk_dev_free(d_count); rsk_dev_free(d_ptr); rsk_dev_free(d_tot); };
#define C_TRY(expr) do { rc = (expr); if (rc != RSK_OK) { cleanup(); return rc; } } while (0)
#define C_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rsk_set_error("%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); return RSK_ERR_CUDA; } } while (0)
    C_TRY(rsk_dev_alloc(&d_count, (size_t)n_rows));
    C_TRY(rsk_dev_alloc(&d_ptr, (size_t)n_rows + 1));
    if (!d_total_rays) {
        C_TRY(rsk_dev_alloc(&d_tot, (size_t)n_rows));
        C_CUDA(cudaMemcpyAsync(d_tot, h_total_rays, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
        d_total_rays = d_tot;
    }
    rsk_csr_count_kernel<<<(unsigned)n_rows, 256, 0, st>>>(block, n_cols, d_count);
    rsk_csr_scan_kernel<<<1, 1024, 0, st>>>(d_count, n_rows, d_ptr);
    ctx->launches += 2;
    C_CUDA(cudaMemcpyAsync(row_ptr, d_ptr, ((size_t)n_rows + 1) * 8, cudaMemcpyDeviceToHost, st));
    C_CUDA(cudaStreamSynchronize(st));
    const int64_t nnz = row_ptr[n_rows];
    if (nnz > 0) {
        C_TRY(rsk_dev_alloc(&ctx->csr_cols, (size_t)nnz));
        C_TRY(rsk_dev_alloc(&ctx->csr_vals, (size_t)nnz));
        rsk_csr_fill_kernel<<<(unsigned)n_rows, 256, 0, st>>>(block, n_cols, d_total_rays, d_ptr, ctx->csr_cols, ctx->csr_vals);
        ctx->launches++;
        C_CUDA(cudaGetLastError());
    }
    ctx->csr_nnz = nnz;
    cleanup();
#undef C_TRY
#undef C_CUDA
    return RSK_OK;
}

extern "C" int rsk_solve_csr(rsk_solve *s, int64_t *row_ptr) {
    RSK_REQUIRE(s && row_ptr, "rsk_solve_csr: null argument");
    RskScope scope(s->ctx);
    RSK_TRY(rsk_ctx_join(s->ctx));
    return rsk_csr_build_impl(s->ctx, s->total, s->n_local, s->n_hist, (const long long *)s->total_rays, nullptr, row_ptr);
}

extern "C" int rsk_tally_block_csr(rsk_tally_block *b, const int64_t *total_rays, int64_t *row_ptr) {
    RSK_REQUIRE(b && total_rays && row_ptr, "rsk_tally_block_csr: null argument");
    RskScope scope(b->ctx);
    return rsk_csr_build_impl(b->ctx, b->d, b->n_rows, b->n_cols, nullptr, total_rays, row_ptr);
}

extern "C" int rsk_csr_fetch(rsk_ctx *ctx, int32_t *cols, double *vals) {
    RSK_REQUIRE(ctx, "rsk_csr_fetch: null context");
    RskScope scope(ctx);
    if (ctx->csr_nnz > 0) {
        RSK_REQUIRE(cols && vals, "rsk_csr_fetch: null output");
        RSK_CUDA(cudaMemcpyAsync(cols, ctx->csr_cols, (size_t)ctx->csr_nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        RSK_CUDA(cudaMemcpyAsync(vals, ctx->csr_vals, (size_t)ctx->csr_nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        RSK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    rsk_dev_free(ctx->csr_cols); rsk_dev_free(ctx->csr_vals);
    ctx->csr_cols = nullptr; ctx->csr_vals = nullptr; ctx->csr_nnz = 0;
    return RSK_OK;
}
