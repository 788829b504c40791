// rsk_solve.cuh -- host-side state of a device solve, shared by rsk_api.cu and rsk_comm.cu.
#pragma once
#include "rsk_stats.cuh"

struct rsk_solve {
    rsk_ctx *ctx = nullptr;
    rsk_scene *scene = nullptr;
    rsk_emitters *em = nullptr;
    int mode = MODE_MATRIX;
    int32_t n_local = 0, n_hist = 0, discrete = 0;
    rsk_solve_params p{};
    int64_t n_tiles = 0;
    int32_t tile_rays = RSK_TILE_RAYS_MAX;
    // device state
    int32_t *min_sid = nullptr;
    int32_t *emit_ids = nullptr, *rot_base = nullptr, *iters_done = nullptr, *done = nullptr, *not_conv = nullptr, *have_prev = nullptr;
    TileDesc *tiles = nullptr;
    int64_t *n_rays_once = nullptr, *total_rays = nullptr, *ray_begin = nullptr, *ray_end = nullptr;
    uint32_t *mask = nullptr;
    float *cp_table = nullptr;
    unsigned long long *iter_tally = nullptr, *rays_traced = nullptr;
    long long *total = nullptr;
    double *mean = nullptr, *m2 = nullptr, *prev = nullptr;
    int32_t *n_active = nullptr;
    int32_t *h_pinned = nullptr;     // [0] n_active
    int32_t last_active = 0;
    bool stepped = false;
    bool external_tally = false;     // iter_tally belongs to the caller (rsk_solve_set_iter_tally_buffer)
    rsk_solve *twin = nullptr;       // dual solves: the sky side (this object is the matrix side)
    rsk_solve *primary = nullptr;    // dual solves, sky side: the matrix side (owner of the pipeline state below)
    // Pipelined stepping: iteration i is traced and folded on stream (i & 1) with its own tally buffer, so the trace
    // of iteration i + 1 fills the SMs while iteration i drains and its statistics run (rsk_api.cu).
    bool pipelined = false;
    unsigned long long *iter_tally2 = nullptr;   // tallies of odd iterations
    cudaEvent_t ev_fold = nullptr;   // recorded after every fold/decide of this solve
    int32_t enq_iters = 0;           // iterations enqueued so far = the iteration index every running job is at
    int32_t cur = 0;                 // stream / buffer of the iteration being enqueued (0 or 1)
    bool has_fold = false;
};

// Device-resident int64 [n_rows][n_cols] block in which the ranks of a sharded solve assemble their results (rsk_comm.cu).
struct rsk_tally_block {
    rsk_ctx *ctx = nullptr;
    int64_t n_rows = 0, n_cols = 0;
    long long *d = nullptr;
};

// stream on which the iteration enqueued last runs (rsk_api.cu)
cudaStream_t rsk_solve_current_stream(rsk_solve *s);
