"""Host orchestrator of the view-factor solves -- the drop-in for the reference's ``raystrack.main``.

Public functions keep the reference's names, arguments, return shapes, errors and log format
(reference src/raystrack/main.py:1689-2185).  What changed underneath: instead of a Python loop that launches
six Numba kernels per (emitter, iteration) and synchronises after each iteration, the solve is handed to
librsk_b200 as ONE device-resident job set: every iteration of every unconverged emitter is a tile range of a
single fused raygen+trace+tally kernel, followed by an on-device statistics/convergence kernel; the host only
reads "how many emitters are still running".
"""
from __future__ import annotations

import functools
import heapq
from concurrent.futures import ThreadPoolExecutor
import os
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .params import MatrixParams, SkyParams
from .prepared import PreparedEmitter, PreparedSolver

Mesh = Tuple[str, np.ndarray, np.ndarray]
_BVH_AUTO_THRESHOLD = 512      # reference main.py:48


LAST_TIMING: Dict[str, float] = {}     # wall-clock seconds per phase of the most recent solve (diagnostics / bench.py)
LAST_PLAN: Dict[str, int] = {}         # how the most recent matrix solve was cut: {"emitters", "first_part"} (0 = one piece)


class _Phase:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.t = time.perf_counter()

    def __exit__(self, *exc):
        LAST_TIMING[self.name] = LAST_TIMING.get(self.name, 0.0) + time.perf_counter() - self.t
        return False


def _log(msg: str) -> None:
    """Progress line sink.  A module attribute on purpose: the reference's validation harness and examples
    replace ``raystrack.main._log`` to capture iteration counts (validation/common_validation.py:139-141)."""
    print(msg)


def _select_bvh(bvh: Optional[str], total_faces: int) -> bool:
    """reference main.py:125-133."""
    mode = (bvh or "auto").lower()
    if mode not in ("auto", "off", "builtin"):
        raise ValueError(f"bvh must be 'auto', 'off', or 'builtin' (got {bvh!r})")
    if mode == "builtin":
        return True
    if mode == "off":
        return False
    return total_faces >= _BVH_AUTO_THRESHOLD


def _resolve_device(device: Optional[str]) -> str:
    """reference main.py:136-147.  Every value runs on the GPU (this package has no CPU path); the returned
    label selects the convergence schedule -- "cpu" checks after every iteration like the reference's CPU loop
    (main.py:1889), "gpu" honours ``convergence_interval`` -- and is what the progress lines print after
    ``device=``, as the reference prints its resolved device (main.py:1938)."""
    dev = (device or "auto").lower()
    if dev not in ("auto", "gpu", "cpu"):
        raise ValueError(f"device must be 'auto', 'gpu', or 'cpu' (got {device!r})")
    if _native.device_count() <= 0:
        if dev == "gpu":
            raise RuntimeError("device='gpu' requested but CUDA is not available")
        raise RuntimeError("CUDA is not available: raystrack_b200 runs on a B200 only and has no CPU path")
    return "cpu" if dev == "cpu" else "gpu"


def _ensure_prepared(meshes: List[Mesh], prepared: Optional[PreparedSolver]) -> PreparedSolver:
    if prepared is None:
        return PreparedSolver(meshes)
    if not isinstance(prepared, PreparedSolver):
        raise TypeError("prepared must be a PreparedSolver instance")
    return prepared


def _surface_masks(emitters: Sequence[PreparedEmitter], centers: np.ndarray, extents: np.ndarray) -> np.ndarray:
    """``surf_active`` for every emitter at once, uint8 [n_emit, n_surf] (reference main.py:167-204): the emitter's
    own mesh is off; for a planar emitter every mesh whose bounding box lies wholly behind its plane is off.
    Same float32 operation order as the reference's scalar loop (dx*nx + dy*ny + dz*nz; |n|.extent)."""
    n = centers.shape[0]
    ne = len(emitters)
    active = np.ones((ne, n), np.uint8)
    if ne == 0 or n == 0:
        return active
    planar = np.fromiter((em.plane_is_planar for em in emitters), bool, ne)
    po = np.stack([em.plane_origin for em in emitters]).astype(np.float32, copy=False)
    pn = np.stack([em.plane_normal for em in emitters]).astype(np.float32, copy=False)
    tol = np.fromiter((em.plane_tol for em in emitters), np.float64, ne).astype(np.float32)
    an = np.abs(pn)
    c = [np.ascontiguousarray(centers[:, k]) for k in range(3)]
    x = [np.ascontiguousarray(extents[:, k]) for k in range(3)]
    rows = np.nonzero(planar)[0]

    def block(lo: int) -> None:                             # row blocks keep the temporaries cache-sized
        r = rows[lo:lo + 128]
        signed = (c[0][None, :] - po[r, 0:1]) * pn[r, 0:1]
        signed += (c[1][None, :] - po[r, 1:2]) * pn[r, 1:2]
        signed += (c[2][None, :] - po[r, 2:3]) * pn[r, 2:3]
        radius = an[r, 0:1] * x[0][None, :]
        radius += an[r, 1:2] * x[1][None, :]
        radius += an[r, 2:3] * x[2][None, :]
        signed += radius
        active[r] = ~(signed <= tol[r, None])

    starts = range(0, rows.size, 128)
    if rows.size * n >= 1 << 20:                            # NumPy releases the GIL: a few threads share the blocks
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            list(pool.map(block, starts))
    else:
        for lo in starts:
            block(lo)
    idx = np.arange(min(ne, n))
    active[idx, idx] = 0
    return active


def _cached_masks(solver: PreparedSolver, emitters, centers, extents, flip_faces: bool,
                  ctx: Optional[_native.Context] = None) -> np.ndarray:
    """Surface masks depend only on the emitter planes and the mesh bounds, so a PreparedSolver keeps them.  With a
    context they are computed on the GPU (``rsk_surface_masks``: the same float32 operations as ``_surface_masks``)."""
    key = ("surface_masks", bool(flip_faces))
    got = solver._derived_cache.get(key)
    if got is None:
        if ctx is not None and len(emitters) and centers.shape[0]:
            ne = len(emitters)
            soa = getattr(emitters, "soa", None)           # summaries that came from the device already hold the arrays
            if soa is None:
                soa = (np.fromiter((em.plane_is_planar for em in emitters), bool, ne), np.stack([em.plane_origin for em in emitters]),
                       np.stack([em.plane_normal for em in emitters]),
                       np.fromiter((em.plane_tol for em in emitters), np.float64, ne).astype(np.float32))
            got = ctx.surface_masks(*soa, centers, extents)
        else:
            got = _surface_masks(emitters, centers, extents)
        got.setflags(write=False)
        solver._derived_cache[key] = got
    return got


def _rotation_table(seed: int, n_emit: int, max_iters: int) -> np.ndarray:
    """Cranley-Patterson offsets.  The reference draws ``default_rng(seed + idx_emit + itr)`` per emitter and
    iteration (main.py:1810-1812); only the sum matters, so one row per distinct sum: row s = rotation of
    ``seed + s``; emitter i, iteration it uses row i + it."""
    rows = max(0, n_emit + max(int(max_iters), 0))
    return _rotation_rows(int(seed), max(rows, 1))


@functools.lru_cache(maxsize=8)
def _rotation_rows(seed: int, rows: int) -> np.ndarray:
    if 0 <= seed and seed + rows <= (1 << 32):
        # every row at once: SeedSequence + PCG64 + the float32 draw restated on integer arrays, bit-identical to the
        # generators (``_pcg.py``; 2041 rows: 1.2 ms instead of 39 ms of generator constructions)
        from ._pcg import rotation_rows
        table = rotation_rows(seed, rows)
    else:
        table = _rotation_rows_generators(seed, rows)
    table.setflags(write=False)
    return table


def _rotation_rows_generators(seed: int, rows: int) -> np.ndarray:
    """The reference's own draws, one generator per row (main.py:1810-1812); also raises what NumPy raises for a
    negative seed."""
    table = np.zeros((rows, 7), np.float32)
    for s in range(rows):
        rng = np.random.default_rng(seed + s)
        table[s, :2] = rng.random(2, dtype=np.float32)
        table[s, 2:] = rng.random(5, dtype=np.float32)
    return table


TILE_RAYS = 4096               # ray slices of split emitters are cut on multiples of this (half the largest CTA tile)


def plan_shards(todo: Sequence[int], n_rays_once: Sequence[int], world: int, allow_split: bool = True,
                cost_per_ray: Optional[Sequence[float]] = None) -> List[List[Tuple[int, int, int, bool]]]:
    """Partition the (emitter, ray range) work of one iteration over ``world`` GPUs.

    Returns, per rank, a list of jobs ``(emitter, ray_begin, ray_end, shared)``.  Emitters are independent units
    (reference main.py:1758-1939) and are assigned whole, costliest first, to the least loaded rank; an emitter that
    alone exceeds an eighth of a rank's fair share is cut into ``world`` tile-aligned ray slices instead (``shared``):
    its per-iteration tallies are summed across ranks before the statistics update, so every rank takes the same
    convergence decision for it.  Shared jobs come first in every rank's list, in the same order.
    ``allow_split=False`` assigns every emitter whole (no per-iteration exchange at all).  ``cost_per_ray`` (indexed by
    emitter, e.g. from ``_emitter_cost_per_ray``) weights the load of an emitter by rays x cost instead of rays alone;
    every rank must pass the same values."""
    world = max(1, int(world))
    plans: List[List[Tuple[int, int, int, bool]]] = [[] for _ in range(world)]
    if world == 1:
        plans[0] = [(int(i), 0, int(n_rays_once[i]), False) for i in todo]
        return plans
    cost = (lambda i: 1.0) if cost_per_ray is None else (lambda i: float(cost_per_ray[i]))
    total = float(sum(int(n_rays_once[i]) for i in todo))
    limit = total / (8.0 * world)
    shared = [int(i) for i in todo if allow_split and n_rays_once[i] > limit and n_rays_once[i] >= 2 * world * TILE_RAYS]
    shared_set = set(shared)
    whole = [int(i) for i in todo if int(i) not in shared_set]
    loads = [0.0] * world
    for i in shared:
        n = int(n_rays_once[i])
        tiles = (n + TILE_RAYS - 1) // TILE_RAYS
        cuts = [min(n, (tiles * r // world) * TILE_RAYS) for r in range(world)] + [n]
        for r in range(world):
            plans[r].append((i, cuts[r], cuts[r + 1], True))
            loads[r] += (cuts[r + 1] - cuts[r]) * cost(i)
    per_rank: List[List[int]] = [[] for _ in range(world)]
    heap = [(loads[q], q) for q in range(world)]                 # least loaded rank first, ties to the lower rank
    heapq.heapify(heap)
    for i in sorted(whole, key=lambda k: (-int(n_rays_once[k]) * cost(k), k)):
        load, r = heapq.heappop(heap)
        heapq.heappush(heap, (load + int(n_rays_once[i]) * cost(i), r))
        per_rank[r].append(i)
    for r in range(world):
        plans[r].extend((i, 0, int(n_rays_once[i]), False) for i in sorted(per_rank[r]))
    return plans


COST_SAMPLE_RAYS = 2048        # rays per emitter of the cost measurement (a QMC prefix covers the emitter evenly)


def _emitter_cost_per_ray(ctx, d_scene, d_em, todo, n_rays_once, active, table, emit_sid, min_sid, rank: int, world: int):
    """Relative cost per ray of every emitter in ``todo`` (1.0 = the mean), for ``plan_shards``: rays are balanced to
    1e-4 by construction, but a ray from a roof (mostly sky) costs less than one from a street-level wall, and the
    spread between ranks at 8 GPUs is ~1.5 % of an iteration.  Every rank measures the emitters ``k % world == rank`` of
    ``todo`` (``rsk_emitter_costs``: SM clock ticks of the first COST_SAMPLE_RAYS rays), the tick and ray counts are summed
    over the ranks, so all ranks plan with identical numbers.  ~1 ms per call."""
    from .dist import allreduce_sum_
    n_emit = active.shape[0]
    ticks = np.zeros(n_emit, np.int64)
    rays = np.zeros(n_emit, np.int64)
    mine = np.asarray([i for k, i in enumerate(todo) if k % world == rank], np.int32)
    if mine.size:
        t, r = _native.emitter_costs(ctx, d_scene.native, d_em.native, mine, active[mine], np.asarray(emit_sid)[mine],
                                     np.asarray(min_sid)[mine], table[0], COST_SAMPLE_RAYS)
        ticks[mine], rays[mine] = t, r
    allreduce_sum_([ticks, rays])
    cost = np.ones(n_emit, np.float64)
    seen = rays > 0
    if seen.any():
        per_ray = ticks[seen] / rays[seen].astype(np.float64)
        mean = float(np.sum(per_ray * np.asarray(n_rays_once, np.float64)[seen]) / max(1.0, float(np.asarray(n_rays_once, np.float64)[seen].sum())))
        if mean > 0.0:
            cost[seen] = np.clip(per_ray / mean, 0.25, 4.0)
    return cost


def _run_solve(solve: _native.Solve, min_iters: int, max_iters: int) -> None:
    """Drive a device solve to completion: first the iterations no stopping rule can interrupt, then small
    batches (the device skips converged emitters on its own, so over-enqueueing is harmless)."""
    if max_iters <= 0 or solve.n_local == 0:
        return
    first = max(1, min(int(max_iters), max(int(min_iters), 1)))
    active = solve.step(first)
    done = first
    while active > 0 and done < max_iters:
        chunk = min(4, max_iters - done)
        active = solve.step(chunk)
        done += chunk


def _run_solve_shared(solve, n_shared: int, min_iters: int, max_iters: int, exchange: str) -> None:
    """Multi-GPU variant when some emitters are ray-split over the ranks: per iteration trace -> all-reduce of the
    shared jobs' iteration tallies -> statistics.  ``exchange`` "native": the library's NCCL communicator, on the
    solve's stream (rsk_solve_allreduce_iter_tallies); "host": torch.distributed on host tensors (CPU stand-ins of
    the tests).  Every rank issues the same sequence of collectives; the loop ends when no rank has a running job."""
    from . import dist as D
    if max_iters <= 0:
        return
    tally = D.attach_tally_tensor(solve, n_shared) if exchange == "host" else None
    done = 0
    chunk = max(1, min(int(max_iters), max(int(min_iters), 1)))
    while done < max_iters:
        for _ in range(chunk):
            solve.enqueue_trace()
            if exchange == "host":
                D.all_reduce_tensor_(tally)
            elif n_shared and solve.n_local:
                solve.allreduce_iter_tallies(n_shared)
            solve.enqueue_fold()
        done += chunk
        active = solve.poll() if solve.n_local else 0
        if D.max_over_ranks(float(active)) <= 0:
            break
        chunk = min(4, max_iters - done)


_DIST_OVERRIDE: Optional[Tuple[int, int]] = None     # tests / bench.py: (0, 1) solves unsharded inside a process group


def _dist_env() -> Tuple[int, int]:
    """(rank, world) of this process: the library's own communicator if there is one, else the torch process group."""
    if _DIST_OVERRIDE is not None:
        return _DIST_OVERRIDE
    from . import dist as D
    if D.native_comm_active():
        return D.native_comm_env()
    return D._torch_group()


def _context() -> _native.Context:
    """The CUDA context of this process.  With more than one rank it is the context that owns the library's NCCL
    communicator (joined on first use when a torch process group exists), so that collectives and kernels are ordered
    on one stream without host synchronisation."""
    if _DIST_OVERRIDE is None:
        from . import dist as D
        ctx = D.native_comm_context(create=True)
        if ctx is not None:
            return ctx
    return _native.Context.for_device()


def _exchange_kind(ctx, world: int) -> Optional[str]:
    """How the ranks exchange tallies DURING a solve (needed to ray-split an emitter): "native" = the library's NCCL
    communicator on the context's stream; "host" = torch.distributed on host tensors, only for the CPU stand-ins of
    the tests; None = no ordered exchange available (e.g. a gloo group on a GPU box without NCCL) -- emitters are then
    assigned whole and the results summed once at the end."""
    if world <= 1:
        return None
    if isinstance(ctx, _native.Context):
        from . import dist as D
        return "native" if D.native_comm_active() and D.native_comm_context() is ctx else None
    return "host"


def _plan_chunks(plan: List[Tuple[int, int, int, bool]], n_hist: int) -> List[List[Tuple[int, int, int, bool]]]:
    """The device keeps ~48 bytes per (job, bin): two iteration tallies, total, Welford mean/M2 (+ previous estimate).  Jobs
    are solved in chunks that fit a memory budget (RSK_SOLVE_MEMORY_MB, default 16384); emitters are independent, so
    chunking cannot change any result.  Ray-split jobs (first in every rank's list, same order everywhere) stay
    together in the first chunk."""
    budget = max(1.0, float(os.environ.get("RSK_SOLVE_MEMORY_MB", "16384")) * (1 << 20))
    max_jobs = max(1, int(budget // (48 * max(1, n_hist))))
    n_shared = sum(1 for j in plan if j[3])
    chunks: List[List[Tuple[int, int, int, bool]]] = []
    head = plan[:max(n_shared, min(len(plan), max_jobs))] if plan else []
    if head or not plan:
        chunks.append(head)
    for lo in range(len(head), len(plan), max_jobs):
        chunks.append(plan[lo:lo + max_jobs])
    return chunks


def _csr_from_dense(tallies: np.ndarray, totals: np.ndarray):
    """Host form of ``rsk_solve_csr``: (row_ptr, cols, vals) of F = tallies / totals over the non-zero bins."""
    n_rows, n_cols = tallies.shape
    nz = np.flatnonzero(tallies.reshape(-1))
    rows = nz // max(n_cols, 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        vals = tallies.reshape(-1)[nz] / totals[rows].astype(np.float64)              # main.py:1922-1923
    return np.searchsorted(rows, np.arange(n_rows + 1)).astype(np.int64), (nz - rows * n_cols).astype(np.int32), vals


def _solve_sharded(ctx, d_scene, d_em, todo, n_rays_once, active, table, *, max_iters, min_iters, interval, tol_mode, tol,
                   emit_sid=None, min_sid=None, sky=False, discrete=False, want_csr=False):
    """Run the jobs of ``todo`` (emitter indices) on this rank's shard and return full-size, rank-summed integer
    results: (tallies int64 [n_emit, n_hist], iterations int64 [n_emit], total rays int64 [n_emit]).  With
    ``want_csr`` the first element is instead the compressed form of the result rows, ``(row_ptr, cols, vals)`` with
    vals = tally / total rays (built on the device when the library holds the rank-summed block)."""
    n_emit = active.shape[0]
    n_surf = active.shape[1]
    rank, world = _dist_env()
    exchange = _exchange_kind(ctx, world)
    cost = None
    if exchange == "native" and not sky and len(todo) >= 4 * world and os.environ.get("RSK_COST_PLAN", "1") != "0":
        with _Phase("cost_plan"):
            cost = _emitter_cost_per_ray(ctx, d_scene, d_em, todo, n_rays_once, active, table, emit_sid, min_sid, rank, world)
    with _Phase("plan"):
        all_plans = plan_shards(todo, n_rays_once, world, allow_split=exchange is not None, cost_per_ray=cost)
        plan = all_plans[rank]
        any_shared = world > 1 and any(j[3] for shard in all_plans for j in shard)
        n_hist = (145 if discrete else 1) if sky else 2 * n_surf
        chunks = _plan_chunks(plan, n_hist)

    single = world == 1 and len(chunks) == 1 and len(plan) == n_emit          # every emitter, in order: no scatter needed
    # one GPU, one solve, emitters in ascending order (some may be missing: emitters without receivers are never
    # traced): the solve's compressed rows only need empty rows spliced in
    single_csr = (want_csr and world == 1 and len(chunks) == 1 and hasattr(_native.Solve, "read_csr")
                  and all(a[0] < b[0] for a, b in zip(plan, plan[1:])))
    # With the library communicator the tally blocks never visit the host before they are summed: every rank scatters
    # the rows of its solves into a zeroed device block, one NCCL all-reduce adds the blocks up over NVLink, and the
    # result comes back in a single copy through pinned memory.
    block = _native.TallyBlock(ctx, n_emit, n_hist) if exchange == "native" else None
    tallies = iters = totals = None
    if single_csr or not single:
        tallies = None if (block is not None or single_csr) else np.zeros((n_emit, n_hist), np.int64)
        iters = np.zeros(n_emit, np.int64)
        totals = np.zeros(n_emit, np.int64)
    try:
        for c, chunk in enumerate(chunks):
            ids = np.asarray([j[0] for j in chunk], np.int32)
            ranges = np.asarray([[j[1], j[2]] for j in chunk], np.int64).reshape(-1, 2)
            kw = {}
            if not sky:
                kw = dict(emit_sid=np.asarray(emit_sid)[ids], min_sid=np.asarray(min_sid)[ids])
            with _Phase("solve_begin"):
                solve = _native.Solve(ctx, d_scene.native, d_em.native, ids,
                                      active[ids] if len(chunk) else np.zeros((0, n_surf), np.uint8),
                                      table, ids.copy(), max_iters=max_iters, min_iters=min_iters, interval=interval,
                                      tol_mode=tol_mode, tol=tol, sky=sky, discrete=discrete, ray_range=ranges, **kw)
            try:
                with _Phase("iterate"):
                    if any_shared and c == 0:
                        _run_solve_shared(solve, sum(1 for j in chunk if j[3]), min_iters, max_iters, exchange)
                    else:
                        _run_solve(solve, min_iters, max_iters)
                keep = np.asarray([not (j[3] and rank != 0) for j in chunk], bool)     # replicated (ray-split) jobs count once
                with _Phase("download"):
                    if block is not None:
                        it_loc, tot_loc = solve.read_counters()
                        block.add_solve(solve, keep)
                        loc = None
                    elif single_csr:
                        it_loc, tot_loc = solve.read_counters()
                        loc = solve.read_csr()
                    elif single and hasattr(solve, "read_block_view"):
                        loc, it_loc, tot_loc = solve.read_block_view()
                    else:
                        loc, it_loc, tot_loc = solve.read_block()
            finally:
                with _Phase("solve_end"):
                    solve.close()
            if single_csr:
                counts = np.zeros(n_emit, np.int64)
                counts[ids] = np.diff(loc[0])
                iters[ids], totals[ids] = it_loc, tot_loc
                return (np.concatenate([np.zeros(1, np.int64), np.cumsum(counts)]), loc[1], loc[2]), iters, totals
            if single:
                if want_csr:
                    loc = _csr_from_dense(loc, tot_loc)
                return loc, it_loc.astype(np.int64), tot_loc
            if keep.any():
                if loc is not None:
                    tallies[ids[keep]] = loc[keep]
                iters[ids[keep]] = it_loc[keep]
                totals[ids[keep]] = tot_loc[keep]
        if world > 1:
            from .dist import allreduce_sum_
            with _Phase("all_reduce"):
                if block is not None:
                    block.allreduce()
                    allreduce_sum_([iters, totals])
                else:
                    allreduce_sum_([tallies, iters, totals], device=getattr(ctx, "device", 0))
        if block is not None:
            with _Phase("rows"):
                tallies = block.read_csr(totals) if want_csr else block.download()
    finally:
        if block is not None:
            block.close()
    if want_csr and not isinstance(tallies, tuple):
        tallies = _csr_from_dense(tallies, totals)
    return tallies, iters, totals


OVERLAP_MIN_EMITTERS = 256            # a solve is cut in two only with at least this many emitters to solve ...
OVERLAP_MIN_RAYS = 1.0e9              # ... and this many rays certain to be traced (rays per iteration x min_iters)
OVERLAP_TAIL_FRACTION = 0.15          # share of the rays left for the second part


def _overlap_split(todo: Sequence[int], n_rays_once: Sequence[int], iters_floor: int) -> int:
    """Where to cut a large matrix solve in two so that the result rows of the first part are assembled (600 k Python
    floats at C5: ~23 ms, a fixed cost that does not shrink with the GPU count) while the GPUs trace the second part.
    Returns k: ``todo[:k]`` is solved first, ``todo[k:]`` second; 0 = solve in one piece.  The cut keeps emitter order
    (rows and reciprocity fill-ins are then written in the reference's order, main.py:1918-1934) and leaves the second
    part the shortest suffix holding OVERLAP_TAIL_FRACTION of the rays -- enough GPU time to hide the first part's rows,
    few rows of its own.  Depends only on values every rank holds, so all ranks cut at the same place."""
    if os.environ.get("RSK_OVERLAP_ASSEMBLY", "1") == "0" or len(todo) < max(2, OVERLAP_MIN_EMITTERS):
        return 0
    rays = np.asarray([int(n_rays_once[i]) for i in todo], np.float64)
    total = float(rays.sum())
    if total * max(1, int(iters_floor)) < OVERLAP_MIN_RAYS:
        return 0
    suffix = np.cumsum(rays[::-1])[::-1]                   # rays of todo[k:]
    ks = np.nonzero(suffix >= OVERLAP_TAIL_FRACTION * total)[0]
    k = int(ks[-1]) if ks.size else 0
    return k if 2 * k >= len(todo) else 0                  # the hidden part must be the larger one in rows


def view_factor_matrix(meshes: List[Mesh], params: MatrixParams, *, prepared: Optional[PreparedSolver] = None,
                       _hook: Optional[dict] = None):
    """View factors between all meshes (reference main.py:1689-1945).

    Returns ``{emitter: {"<receiver>_front" | "<receiver>_back": F}}`` with only positive entries; with
    ``reciprocity=True`` receivers ``j > i`` are traced and ``F_ji = F_ij * A_i / A_j`` is filled in for front hits.
    Inside an initialised ``torch.distributed`` group the emitters are sharded over the ranks and the integer
    tallies all-reduced, every rank returns the full result."""
    if not isinstance(params, MatrixParams):
        raise TypeError("params must be a MatrixParams instance")
    t_call = time.perf_counter()
    p = params.as_dict()
    samples, rays, seed = p["samples"], p["rays"], p["seed"]
    max_iters, tol, tol_mode, min_iters = p["max_iters"], p["tol"], p["tol_mode"], p["min_iters"]
    interval = max(1, int(p["convergence_interval"]))
    reciprocity, flip_faces = p["reciprocity"], p["flip_faces"]

    schedule = _resolve_device(p["device"])
    solver = _ensure_prepared(meshes, prepared)
    use_bvh = _select_bvh(p["bvh"], solver.total_faces)
    if tol_mode not in ("stderr", "delta"):
        raise ValueError(f"Unknown tol_mode: {tol_mode}")
    n_surf = len(meshes)
    result: Dict[str, Dict[str, float]] = {name: {} for name, _, _ in meshes}
    LAST_TIMING.clear()
    with _Phase("host_prepare"):
        centers, extents = solver.get_mesh_bounds()
        ctx = _context()
    with _Phase("scene_upload_bvh"):
        d_scene = solver.get_device_scene(use_bvh=use_bvh, ctx=ctx)
    with _Phase("emitter_upload"):
        d_em = solver.get_device_emitters(samples=samples, rays=rays, flip_faces=flip_faces, ctx=ctx)
        emitters = solver.get_emitter_summaries(samples=samples, rays=rays, flip_faces=flip_faces, ctx=ctx)
        areas = [em.total_area for em in emitters] if reciprocity else None

    t0 = time.time()
    with _Phase("masks"):
        active = _cached_masks(solver, emitters, centers, extents, flip_faces, ctx)
    t_recv = time.perf_counter()
    # receivers of emitter i (main.py:161-164, 207-214): active meshes j > i (reciprocity) or j != i
    emit_sid = np.arange(n_surf, dtype=np.int32)
    min_sid = (emit_sid + 1) if reciprocity else np.zeros(n_surf, np.int32)
    if reciprocity:
        # suffix-any: emitter i has receivers iff some active mesh j > i exists (the diagonal is already off)
        has_recv = np.asarray([bool(active[i, i + 1:].any()) for i in range(n_surf)], bool)
    else:
        has_recv = active.any(axis=1)
    # a mesh without triangles shoots nothing (the reference fails on it; here it reports "0 iter, 0 rays")
    has_recv = has_recv & np.fromiter((np.asarray(f).shape[0] > 0 for _, _, f in meshes), bool, n_surf)
    todo = [i for i in range(n_surf) if has_recv[i]]

    weights = [float(em.n_cells * rays) for em in emitters]
    n_once = [int(em.n_cells * rays) for em in emitters]
    LAST_TIMING["receivers"] = time.perf_counter() - t_recv
    # result rows (main.py:1918-1934).  The tally block is [emitter][receiver][front, back], i.e. already in the
    # reference's key order "<r0>_front, <r0>_back, <r1>_front, ..."; only non-zero bins become keys.
    label = "builtin" if use_bvh else "off"
    keys = [f"{name}{suffix}" for name, _, _ in meshes for suffix in ("_front", "_back")]
    keys_obj = np.asarray(keys, dtype=object) if keys else None
    has_recv_l = has_recv.tolist()
    tail = f"(BVH={label}, device={schedule})"
    weights_a = np.asarray(weights)

    def assemble(i_lo: int, i_hi: int, tallies, iters, totals, elapsed: float) -> None:
        """Rows of emitters i_lo <= i < i_hi from the compressed (or dense) tallies of a finished solve."""
        work = weights_a[i_lo:i_hi] * np.maximum(iters[i_lo:i_hi], 1)
        work_sum = max(1.0, float(work.sum()))
        # non-zero bins per emitter: column indices (row-major, i.e. keys in order) and F = hits / total rays (main.py:1922-1923)
        row_ptr, nz_cols, nz_vals = tallies if isinstance(tallies, tuple) else _csr_from_dense(tallies, totals)
        bounds = row_ptr.tolist()
        lo_all, hi_all = bounds[i_lo], bounds[i_hi]
        vals_all = nz_vals[lo_all:hi_all].tolist()
        cols_all = nz_cols[lo_all:hi_all].tolist() if reciprocity else None
        keys_nz = keys_obj[nz_cols[lo_all:hi_all]].tolist() if keys_obj is not None else []
        iters_l, totals_l = iters[i_lo:i_hi].tolist(), totals[i_lo:i_hi].tolist()     # plain Python numbers for the log lines
        share_l = (elapsed * work / work_sum).tolist()
        for i in range(i_lo, i_hi):
            name_e = meshes[i][0]
            if not has_recv_l[i]:
                if _hook is None:
                    _log(f"({i+1}/{n_surf}) [{name_e}] 0 iter, 0 rays -> 0.000s  (BVH={label}, device={schedule})")
                continue
            lo, hi = bounds[i] - lo_all, bounds[i + 1] - lo_all
            vals = vals_all[lo:hi]
            row = dict(zip(keys_nz[lo:hi], vals))
            if reciprocity and areas is not None:
                for c, f in zip(cols_all[lo:hi], vals):
                    j = c >> 1
                    if not (c & 1) and areas[j] > 0.0:
                        result[meshes[j][0]][f"{name_e}_front"] = f * (areas[i] / areas[j])     # main.py:1926-1927
            result[name_e].update(row)
            if _hook is None:
                _log(f"({i+1}/{n_surf}) [{name_e}] {iters_l[i - i_lo]} iter, {totals_l[i - i_lo]:,} rays -> {share_l[i - i_lo]:0.3f}s  {tail}")

    if _hook is not None and "precomputed" in _hook:
        tallies, iters, totals = _hook["precomputed"]            # shared-ray solve already ran (view_factor_matrix_and_sky)
        elapsed = time.time() - t0
        t_asm = time.perf_counter()
        assemble(0, n_surf, tallies, iters, totals, elapsed)
    else:
        with _Phase("rotations"):
            table = _rotation_table(seed, n_surf, max_iters)
        solve_kw = dict(max_iters=max_iters, min_iters=min_iters, interval=interval if schedule == "gpu" else 1,
                        tol_mode=tol_mode, tol=tol, emit_sid=emit_sid, min_sid=min_sid, want_csr=True)
        cut = _overlap_split(todo, n_once, min_iters)
        LAST_PLAN.update(emitters=len(todo), first_part=cut)
        if cut:
            # two solves: the rows of the first are built on a worker thread while the GPUs trace the second (the
            # main thread spends that time inside the library with the GIL released)
            import sys
            import threading
            first_of_second = todo[cut]
            res_a = _solve_sharded(ctx, d_scene, d_em, todo[:cut], n_once, active, table, **solve_kw)
            err: List[BaseException] = []

            def work_a(elapsed_a=time.time() - t0):
                try:
                    assemble(0, first_of_second, *res_a, elapsed_a)
                except BaseException as exc:                     # re-raised on the calling thread
                    err.append(exc)

            worker = threading.Thread(target=work_a, name="rsk-rows")
            switch = sys.getswitchinterval()
            sys.setswitchinterval(1e-4)                          # the enqueueing thread must get the GIL back at once
            t1 = time.time()
            worker.start()
            try:
                res_b = _solve_sharded(ctx, d_scene, d_em, todo[cut:], n_once, active, table, **solve_kw)
            finally:
                t_asm = time.perf_counter()
                worker.join()
                sys.setswitchinterval(switch)
            if err:
                raise err[0]
            iters, totals = res_a[1] + res_b[1], res_a[2] + res_b[2]
            elapsed = time.time() - t0
            assemble(first_of_second, n_surf, res_b[0], res_b[1], res_b[2], time.time() - t1)
        else:
            tallies, iters, totals = _solve_sharded(ctx, d_scene, d_em, todo, n_once, active, table, **solve_kw)
            elapsed = time.time() - t0
            t_asm = time.perf_counter()
            assemble(0, n_surf, tallies, iters, totals, elapsed)

    LAST_TIMING["assemble"] = time.perf_counter() - t_asm
    LAST_TIMING["other"] = (time.perf_counter() - t_call) - sum(LAST_TIMING.values())      # everything no phase above covers
    if _hook is not None:
        _hook.update(iters=iters, totals=totals, n_once=n_once, label=label, elapsed=elapsed, device=schedule)
        return result
    if p["enforce_reciprocity_rowsum"]:
        from .reciprocity import enforce_reciprocity_and_rowsum
        with _Phase("reciprocity_rowsum"):
            enforce_reciprocity_and_rowsum(result, meshes, areas, ctx=ctx)
    return result


def view_factor(sender, receiver, params: MatrixParams, *, prepared: Optional[PreparedSolver] = None):
    """Rows of the sender meshes only (reference main.py:1948-1954)."""
    senders = [sender] if isinstance(sender, tuple) else list(sender)
    receivers = [receiver] if isinstance(receiver, tuple) else list(receiver)
    vf_all = view_factor_matrix(senders + receivers, params=params, prepared=prepared)
    return {s[0]: vf_all.get(s[0], {}) for s in senders}


def view_factor_to_tregenza_sky(meshes: List[Mesh], params: SkyParams, *, prepared: Optional[PreparedSolver] = None,
                                _hook: Optional[dict] = None):
    """Sky view factors per mesh (reference main.py:1957-2185): rays that hit no other active mesh and point
    upward are binned into the 145 Tregenza patches (``discrete=True``) or counted as one "Sky" entry."""
    if not isinstance(params, SkyParams):
        raise TypeError("params must be a SkyParams instance")
    if len(meshes) == 0:
        raise ValueError("meshes must not be empty")
    p = params.as_dict()
    samples, rays, seed = p["samples"], p["rays"], p["seed"]
    max_iters, tol, tol_mode, min_iters = p["max_iters"], p["tol"], p["tol_mode"], p["min_iters"]
    interval = max(1, int(p["convergence_interval"]))
    discrete = bool(p["discrete"])

    schedule = _resolve_device(p["device"])
    solver = _ensure_prepared(meshes, prepared)
    use_bvh = _select_bvh(p["bvh"], solver.total_faces)
    if tol_mode not in ("stderr", "delta"):
        raise ValueError(f"Unknown tol_mode: {tol_mode}")
    keys = [f"Sky_Patch_{i}" for i in range(1, 146)] if discrete else ["Sky"]
    result: Dict[str, Dict[str, float]] = {name: {k: 0.0 for k in keys} for name, _, _ in meshes}
    n_surf = len(meshes)
    if n_surf <= 1 and _hook is None:                                                  # main.py:1998-1999
        return result

    ctx = _context()
    centers, extents = solver.get_mesh_bounds()
    d_scene = solver.get_device_scene(use_bvh=use_bvh, ctx=ctx)
    d_em = solver.get_device_emitters(samples=samples, rays=rays, flip_faces=False, ctx=ctx)     # main.py:1985
    emitters = solver.get_emitter_summaries(samples=samples, rays=rays, flip_faces=False, ctx=ctx)
    t0 = time.time()
    active = _cached_masks(solver, emitters, centers, extents, False, ctx)
    weights = [float(em.n_cells * rays) for em in emitters]
    n_once = [int(em.n_cells * rays) for em in emitters]
    if _hook is not None and "precomputed" in _hook:
        counts, iters, totals = _hook["precomputed"]
    else:
        table = _rotation_table(seed, n_surf, max_iters)
        nonempty = [i for i, (_, _, f) in enumerate(meshes) if np.asarray(f).shape[0] > 0]
        counts, iters, totals = _solve_sharded(ctx, d_scene, d_em, nonempty, n_once, active, table, max_iters=max_iters,
                                               min_iters=min_iters, interval=interval if schedule == "gpu" else 1,
                                               tol_mode=tol_mode, tol=tol, sky=True, discrete=discrete)
    elapsed = time.time() - t0

    label = "builtin" if use_bvh else "off"
    work = np.asarray(weights) * np.maximum(iters, 1)
    work_sum = max(1.0, float(work.sum()))
    # counts / max(1, total rays) per mesh (main.py:2160-2178): one float64 division per bin, as the reference's
    # ``counts.astype(np.float64) / denom``; rows become dictionaries over the pre-built key list
    denom = np.maximum(1, np.asarray(totals, np.int64)).astype(np.float64)
    rows = (np.asarray(counts)[:, :len(keys)].astype(np.float64) / denom[:, None]).tolist()
    iters_l, totals_l = np.asarray(iters).tolist(), np.asarray(totals).tolist()
    share_l = (elapsed * work / work_sum).tolist()
    for i, (name_e, _, _) in enumerate(meshes):
        result[name_e] = dict(zip(keys, rows[i]))
        if _hook is None:
            _log(f"({i+1}/{n_surf}) [{name_e}] {int(iters_l[i])} iter, {int(totals_l[i]):,} rays -> {share_l[i]:0.3f}s  "
                 f"(BVH={label}, device={schedule})")
    if _hook is not None:
        _hook.update(iters=iters, totals=totals, n_once=n_once, label=label, elapsed=elapsed, device=schedule)
    return result


def outside_workflow_shareable(matrix_params: MatrixParams, sky_params: SkyParams) -> bool:
    """True when the matrix and the sky solve may share one ray set (reference main.py:1185-1206): identical
    ``samples, rays, seed, bvh, device, cuda_async, gpu_raygen`` and ``flip_faces=False`` on the matrix side."""
    if bool(matrix_params.flip_faces):
        return False
    fields = ("samples", "rays", "seed", "bvh", "device", "cuda_async", "gpu_raygen")
    return all(getattr(matrix_params, k) == getattr(sky_params, k) for k in fields)


def _shared_ray_solve(meshes, matrix_params: MatrixParams, sky_params: SkyParams, solver: PreparedSolver):
    """One dual device solve (csrc: MODE_DUAL): every ray is traced once, the closest receiver hit feeds the matrix
    tallies and the any-hit flag the sky bins (reference trace_cpu_[bvh_]combined, cpu_trace.py:280-522)."""
    mp, sp = matrix_params.as_dict(), sky_params.as_dict()
    schedule = _resolve_device(mp["device"])
    use_bvh = _select_bvh(mp["bvh"], solver.total_faces)
    n_surf = len(meshes)
    centers, extents = solver.get_mesh_bounds()
    ctx = _context()
    d_scene = solver.get_device_scene(use_bvh=use_bvh, ctx=ctx)
    d_em = solver.get_device_emitters(samples=mp["samples"], rays=mp["rays"], flip_faces=False, ctx=ctx)
    emitters = solver.get_emitter_summaries(samples=mp["samples"], rays=mp["rays"], flip_faces=False, ctx=ctx)
    active = _cached_masks(solver, emitters, centers, extents, False, ctx)
    ids = np.arange(n_surf, dtype=np.int32)
    min_sid = (ids + 1) if mp["reciprocity"] else np.zeros(n_surf, np.int32)
    table = _rotation_table(mp["seed"], n_surf, max(int(mp["max_iters"]), int(sp["max_iters"])))

    def side(p):
        return dict(max_iters=p["max_iters"], min_iters=p["min_iters"], tol=p["tol"], tol_mode=p["tol_mode"],
                    interval=max(1, int(p["convergence_interval"])) if schedule == "gpu" else 1)

    # Emitters are independent: with several ranks every rank solves its share (whole emitters, plus a ray slice of
    # every oversized one, exactly as _solve_sharded does) and the integer results are summed once at the end.
    rank, world = _dist_env()
    exchange = _exchange_kind(ctx, world)
    n_once = [int(em.n_cells * int(mp["rays"])) for em in emitters]
    n_sky = 145 if sp["discrete"] else 1
    blocks = (_native.TallyBlock(ctx, n_surf, 2 * n_surf), _native.TallyBlock(ctx, n_surf, n_sky)) if exchange == "native" else None
    out_m = [None if blocks else np.zeros((n_surf, 2 * n_surf), np.int64), np.zeros(n_surf, np.int64), np.zeros(n_surf, np.int64)]
    out_s = [None if blocks else np.zeros((n_surf, n_sky), np.int64), np.zeros(n_surf, np.int64), np.zeros(n_surf, np.int64)]
    limit = max(int(mp["max_iters"]), int(sp["max_iters"]))
    first = max(1, min(limit, max(1, min(int(mp["min_iters"]), int(sp["min_iters"])))))
    all_plans = plan_shards(list(range(n_surf)), n_once, world, allow_split=exchange == "native")
    any_shared = world > 1 and any(j[3] for shard in all_plans for j in shard)
    chunks = _plan_chunks(all_plans[rank], 2 * n_surf + n_sky)
    # one GPU, one chunk: every emitter is a job, in order -- the matrix rows leave the device compressed (non-zero bins
    # only) instead of as the dense [n, 2n] int64 block (64 MB at C5), exactly as in the matrix-only solve
    csr_single = blocks is None and world == 1 and len(chunks) == 1 and len(chunks[0]) == n_surf and hasattr(_native.Solve, "read_csr")
    try:
        for c, plan in enumerate(chunks):
            mine = np.asarray([j[0] for j in plan], np.int32)
            ranges = np.asarray([[j[1], j[2]] for j in plan], np.int64).reshape(-1, 2)
            n_shared = sum(1 for j in plan if j[3])
            keep = np.asarray([not (j[3] and rank != 0) for j in plan], bool)      # ray-split jobs are replicated: count once
            solve = _native.DualSolve(ctx, d_scene.native, d_em.native, mine, active[mine] if mine.size else np.zeros((0, n_surf), np.uint8),
                                      table, mine.copy(), ids[mine], min_sid[mine], side(mp), side(sp), bool(sp["discrete"]),
                                      ray_range=ranges if world > 1 else None)
            try:
                if any_shared and c == 0:
                    # split-phase: trace, sum the iteration tallies of the ray-split jobs (both sides) over the ranks, fold
                    from . import dist as D
                    done, chunk = 0, first
                    while done < limit:
                        for _ in range(chunk):
                            solve.enqueue_trace()
                            if n_shared and mine.size:
                                solve.matrix_part.allreduce_iter_tallies(n_shared)
                                solve.sky_part.allreduce_iter_tallies(n_shared)
                            solve.matrix_part.enqueue_fold()
                            solve.sky_part.enqueue_fold()
                        done += chunk
                        running = (solve.matrix_part.poll() + solve.sky_part.poll()) if mine.size else 0
                        if D.max_over_ranks(float(running)) <= 0:
                            break
                        chunk = min(4, limit - done)
                elif mine.size and limit > 0:
                    running = solve.step(first)
                    done = first
                    while running > 0 and done < limit:
                        chunk = min(4, limit - done)
                        running = solve.step(chunk)
                        done += chunk
                if mine.size:
                    for k, (part, out) in enumerate(((solve.matrix_part, out_m), (solve.sky_part, out_s))):
                        if blocks:
                            i, r = part.read_counters()
                            blocks[k].add_solve(part, keep)
                        elif csr_single and k == 0:
                            i, r = part.read_counters()
                            out[0] = part.read_csr()
                        else:
                            t, i, r = part.read_block()
                            out[0][mine[keep]] = t[keep]
                        out[1][mine[keep]], out[2][mine[keep]] = i.astype(np.int64)[keep], r[keep]
            finally:
                solve.close()
        if blocks:
            for k, out in enumerate((out_m, out_s)):
                blocks[k].allreduce()
                out[0] = blocks[k].download(copy=True)          # two blocks share the context's staging area: copy
        if world > 1:
            from .dist import allreduce_sum_
            if blocks:
                allreduce_sum_([out_m[1], out_m[2], out_s[1], out_s[2]])
            else:
                allreduce_sum_([*out_m, *out_s], device=getattr(ctx, "device", 0))
    finally:
        for blk in blocks or ():
            blk.close()
    return tuple(out_m), tuple(out_s)


def view_factor_matrix_and_sky(meshes: List[Mesh], *, matrix_params: MatrixParams, sky_params: SkyParams,
                               prepared: Optional[PreparedSolver] = None):
    """Scene view factors and sky view factors from the same per-iteration ray samples (reference
    main.py:1209-1686).  Both sides converge independently and use, per emitter and iteration, exactly the rays the
    separate solves would use, so the results equal ``view_factor_matrix`` + ``view_factor_to_tregenza_sky``
    (the reference documents the same equivalence, main.py:1231-1234).  Every ray is traced ONCE by the dual kernel
    (closest receiver hit + any-hit flag); under torch.distributed each rank does so for its share of the emitters
    and the integer tallies are summed once at the end.
    ``enforce_reciprocity_rowsum`` is not applied here (main.py never does in this function)."""
    if not isinstance(matrix_params, MatrixParams):
        raise TypeError("matrix_params must be a MatrixParams instance")
    if not isinstance(sky_params, SkyParams):
        raise TypeError("sky_params must be a SkyParams instance")
    if not outside_workflow_shareable(matrix_params, sky_params):
        raise ValueError("matrix_params and sky_params are not compatible for shared tracing")
    solver = _ensure_prepared(meshes, prepared)
    mh: dict = {}
    sh: dict = {}
    if len(meshes) > 0:
        mh["precomputed"], sh["precomputed"] = _shared_ray_solve(meshes, matrix_params, sky_params, solver)
    vf_scene = view_factor_matrix(meshes, matrix_params, prepared=solver, _hook=mh)
    sky_vf = view_factor_to_tregenza_sky(meshes, sky_params, prepared=solver, _hook=sh)
    n = len(meshes)
    for i, (name_e, _, _) in enumerate(meshes):
        m_it, s_it = int(mh["iters"][i]), int(sh["iters"][i])
        traced = max(m_it, s_it)
        _log(f"({i+1}/{n}) [{name_e}] traced {traced} iter, {traced * int(mh['n_once'][i]):,} rays -> "
             f"{(mh['elapsed'] + sh['elapsed']) / max(1, n):0.3f}s  (scene={m_it} iter, sky={s_it} iter, "
             f"BVH={mh['label']}, device={mh['device']})")
    return vf_scene, sky_vf


__all__ = ["view_factor_matrix", "view_factor", "view_factor_to_tregenza_sky", "view_factor_matrix_and_sky",
           "outside_workflow_shareable"]
