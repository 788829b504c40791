// rsk_stats.cu -- on-device iteration statistics, convergence decisions and the reciprocity/row-sum solver.
//
// Replaces the host-side bookkeeping of the reference's iteration loops: `hits += hits_iter`, the Welford update
// of the per-iteration fractions, `_convergence_checkpoint` and the stderr/delta stop rules
// (main.py:217-228, 1872-1909 for the matrix; main.py:2122-2174 for the sky), plus kernel_accumulate_hits[_stats]
// (utils/cuda_trace.py:581-616).  All float64 arithmetic uses explicit round-to-nearest intrinsics in the
// reference's NumPy operation order (no FMA contraction), so decisions are bit-identical to the CPU reference
// whenever the integer tallies are.
#include "rsk_stats.cuh"


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
    if (a.done[k]) return;
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

int rsk_launch_fold(rsk_ctx *ctx, const FoldArgs &a) {
    const int64_t n = (int64_t)a.n_local * a.n_hist;
    if (n == 0) return RSK_OK;
    rsk_fold_kernel<<<rsk_blocks(n, 256), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    RSK_CUDA(cudaGetLastError());
    return RSK_OK;
}

int rsk_launch_decide(rsk_ctx *ctx, const DecideArgs &a) {
    if (a.n_local == 0) return RSK_OK;
    rsk_decide_kernel<<<rsk_blocks(a.n_local, 128), 128, 0, ctx->stream>>>(a);
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
