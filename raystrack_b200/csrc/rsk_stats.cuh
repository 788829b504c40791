// rsk_stats.cuh -- argument blocks of the statistics / convergence kernels (rsk_stats.cu).
#pragma once
#include "rsk_common.cuh"

struct FoldArgs {
    unsigned long long *iter_tally;  // [n_local][n_hist], zeroed after folding
    long long *total;                // [n_local][n_hist]
    double *mean, *m2;               // [n_local][n_hist]
    double *prev;                    // [n_local][n_hist] (delta mode) or null
    const uint32_t *surf_mask;       // [n_local][mask_words] or null (sky: every bin counts)
    const int64_t *n_rays_once;      // [n_local]
    const int32_t *iters_done;       // [n_local]
    const int64_t *total_rays;       // [n_local]
    const int32_t *done;             // [n_local]
    int32_t *not_converged;          // [n_local]
    int32_t n_local, n_hist, n_surf, mask_words;
    int32_t max_iters, min_iters, interval, tol_mode;   // tol_mode 0 = stderr, 1 = delta
    int32_t scalar_sky;              // 1: merged sky formula sqrt(max(M2/(n-1),0))/sqrt(n)  (main.py:2170)
    double tol;
};

struct DecideArgs {
    int32_t *iters_done;
    int64_t *total_rays;
    int32_t *done;
    int32_t *not_converged;
    int32_t *have_prev;              // delta mode: a previous checkpoint exists
    const int64_t *n_rays_once;
    const int64_t *ray_begin, *ray_end;   // this rank's slice of each job (for the rays-traced counter)
    int32_t *n_active;               // scalar counter, zeroed by the launcher
    unsigned long long *rays_traced; // scalar
    int32_t n_local;
    int32_t max_iters, min_iters, interval, tol_mode;
};

int rsk_launch_fold(rsk_ctx *ctx, const FoldArgs &a, cudaStream_t stream);
int rsk_launch_decide(rsk_ctx *ctx, const DecideArgs &a, cudaStream_t stream);
