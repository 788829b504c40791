#!/usr/bin/env python3
"""Benchmark of the Monte-Carlo view-factor hot path on the synthetic 1M-triangle urban scene (BASELINE.json
config #5, SURVEY.md 8d "C5").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* = one Monte-Carlo iteration of every emitter of the scene: 2001 emitters, 239 026 176 rays, each ray
generated (QMC sampler), traced to its closest hit through the wide BVH and tallied per receiver x {front, back},
followed by the on-device statistics/convergence pass.  Prints ONE JSON line (rank 0).

  value     closest-hit Grays/s of the whole job, inputs resident in HBM: K steps enqueued back to back, CUDA events on
            the launching streams around all of them, max over ranks; the L2 is evicted before every step (160 MB
            scratch write on the trace stream, inside the timed region)
  e2e       the same metric through the public API ``view_factor_matrix`` (C ABI underneath) with HOST buffers:
            every call starts from the mesh list alone -- vertices + faces uploaded, triangle/emitter records and the
            BVH built on the GPU, 40 iterations, tally download and result assembly
  roofline  memory roofline of the dominant kernel (rsk_trace_kernel<matrix,bvh>): algorithmic bytes per ray
            (reference data layout, counted by the oracle's instrumented replay) x rays / kernel time
  parity    results independent of the GPU count: crc32 of the rank-summed C5 tally block after 2 fixed iterations,
            equal to the block of an unsharded solve (N > 1) and to the committed single-GPU crc; reference goldens
            through the sharded public API; per-ray agreement with the oracle on the CPU sample (N = 1)
  cpu_baseline      the CPU oracle port (oracle/, C + OpenMP, all host threads) on a bounded sample of the workload
  reference_baselines   (N = 1) the UNMODIFIED reference from baseline/_ref on the same box: its Numba CPU kernels and its
            Numba-CUDA path (baseline/numba_baselines.py, separate process)
  --impl reference  times the reference's own Numba CPU kernels (baseline/_ref) on the bounded sample, all host threads
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "closest-hit Grays/s (1M-triangle urban scene)"
UNIT = "Grays/s"
WORKLOAD = "C5 synthetic urban block: 2001 meshes, 1,026,048 triangles, samples=4 rays=64 (239,026,176 rays/iteration)"
SAMPLE_STRIDE = 16          # CPU sample: every 16th emitter (126 emitters incl. the ground) ...
SAMPLE_RAYS = 16384         # ... first 16384 rays of each  (= 2,064,384 rays per CPU step)
PARITY_ITERS = 2            # fixed iterations of the sharded-parity solve (crc of the rank-summed tally block)
FALLBACK_BYTES_PER_RAY = 5400.0   # SURVEY.md 8d, used only if the live replay is not run (N>1) and no file exists


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to stdout) must not add to it:
# file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a private copy of the real stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU, sampled under load through NVML in-process (nvidia_ml_py).

    One GPU: clock and throttle reasons every 50 ms DURING the timed region, first sample as soon as the K steps are
    enqueued (measured harmless: 134 samples at 5 ms spacing leave the C5 step at 66.2 ms).

    Several GPUs: NVML queries during an NCCL-coupled step sequence stall the job.  Measured at 2 GPUs (5 steps of
    33.2 ms, gpurun_out/r2w*): clock + reasons every 5 ms -> 40.8 ms per step, reasons alone -> 35.8, clock alone -> 34.2,
    no sampling -> 33.2; at 8 GPUs 20 samples turned 8.56 ms steps into 9.81 ms.  So inside a multi-GPU timed region only
    the cheap query runs -- the SM clock, once at its start (again after every full second) -- and the throttle reasons (with another clock
    reading) are taken under the same load immediately before it, while the warm-up steps execute (``adjacent``).  ``RSK_BENCH_SAMPLE_MS`` overrides the interval ("off": no
    sampling); ``nvidia-smi -lms`` (the recipe's command) is the fall-back when NVML cannot be loaded -- it needs ~1 s
    to print its first row and saw nothing of the 43 ms region of the first round-2 lines at 8 GPUs."""

    REASONS = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))

    def __init__(self, device: int, world: int = 1):
        self.device = device
        self.multi = world > 1
        self.sm, self.sm_adjacent, self.mx, self.reasons = [], [], 0, set()
        self.proc = self.nv = self.handle = self.thread = None
        self.running = False
        self.source = None
        mode = os.environ.get("RSK_BENCH_SAMPLE_MS", "1000" if self.multi else "50")
        self.off = mode == "off"
        self.interval = 0.05 if self.off else max(0.001, float(mode) * 1e-3)
        if self.off:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.mx = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.source = "nvml"
        except Exception:
            self.nv = None

    def sample(self, reasons: bool = True, adjacent: bool = False):
        nv = self.nv
        if nv is None:
            return
        try:
            (self.sm_adjacent if adjacent else self.sm).append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
            if reasons:
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, const in self.REASONS:
                    if bits & getattr(nv, const):
                        self.reasons.add(name)
        except Exception:
            pass

    def sample_adjacent(self):
        """Multi-GPU only: clock + reasons while work enqueued just outside the timed region executes."""
        if self.multi and not self.off:
            self.sample(reasons=True, adjacent=True)

    def _loop(self):
        while self.running:
            time.sleep(self.interval)
            if self.running:
                self.sample(reasons=not self.multi)

    def start(self):
        if self.off:
            return
        if self.nv is not None:
            self.running = True
            self.sample(reasons=not self.multi)
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(int(r[0]))
                self.mx = max(self.mx, int(r[1]))
            except Exception:
                continue
            for col, (name, _) in enumerate(self.REASONS, start=2):
                if len(r) > col and r[col].lower().startswith("active"):
                    self.reasons.add(name)

    def pause(self):
        """End of the timed region: stop the thread, keep what was collected (no sample here: the GPU is idle again)."""
        if self.thread is not None:
            self.running = False
            self.thread.join(timeout=1.0)
            self.thread = None

    def stop(self):
        self.pause()
        if self.proc:
            self.proc.terminate()
        sm = self.sm
        out = {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons),
               "samples": len(sm), "source": self.source}
        if self.multi and self.nv is not None:
            out["adjacent"] = {"sm_mhz": self.sm_adjacent, "when": "while the warm-up steps right before the timed region execute (same "
                               "load); the throttle reasons come from this sample -- NVML's reasons query stalls NCCL-coupled "
                               "steps (bench.py ClockSampler), inside the region only the SM clock is read"}
        return out


def build_scene(side: int):
    from raystrack_b200 import synthetic
    return synthetic.urban_block(side)


# --------------------------------------------------------------------------------------------- CPU (oracle) legs

class CpuSample:
    """Bounded sample of the workload for the CPU oracle: SAMPLE_RAYS first rays of every SAMPLE_STRIDE-th emitter."""

    def __init__(self, meshes, samples, rays, seed):
        from oracle import oracle as O
        self.O = O
        t = time.time()
        self.solver = O.OracleSolver(meshes)
        self.scene = self.solver.scene(True)                       # reference BVH (utils/bvh.py), built in Python
        self.ems = self.solver.emitters(samples, rays, False)
        self.centers, self.extents = self.solver.bounds()
        self.emit = list(range(0, len(meshes), SAMPLE_STRIDE))
        self.counts = [min(SAMPLE_RAYS, self.ems[i].n_rays_once) for i in self.emit]
        O.halton_dims(max(self.counts))
        self.masks = [O.surface_mask(i, self.ems[i], self.centers, self.extents) for i in self.emit]
        self.seed = seed
        self.prep_s = time.time() - t
        self.rays_per_step = int(sum(self.counts))

    def step(self, itr: int, stats=None, keep=None):
        O = self.O
        for i, n, act in zip(self.emit, self.counts, self.masks):
            cpg, cpd = O.rotation(self.seed, i, itr)
            o, d = O.build_rays(self.ems[i], cpg, cpd, count=n)
            hs, fr = O.trace_firsthit(self.scene, o, d, act, i, 0, stats=stats)
            if keep is not None:
                keep[i] = (hs, fr)

    def describe(self):
        return (f"every {SAMPLE_STRIDE}th emitter ({len(self.emit)} incl. ground), first {SAMPLE_RAYS} rays of each = "
                f"{self.rays_per_step} rays/step, reference BVH (median split, leaf 8), closest hit")


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: its Numba kernels build_rays +
    trace_cpu_bvh_firsthit with its own preparation, imported unmodified from baseline/_ref (kind "reference"); the C
    port under oracle/ only if the reference cannot be imported there (kind "port").  All host threads in either case
    -- torchrun exports OMP_NUM_THREADS=1, which is overridden here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "baseline"))
    import numba_baselines as NB
    threads = NB.prepare_environment(host_threads())          # before numba / OpenMP start
    meshes = build_scene(args.side)
    why = NB.reference_available()
    if why is None:
        try:
            cpu = NB.NumbaCpuSample(meshes, args.samples, args.rays, args.seed)
            kind, cores = "reference", cpu.threads
        except Exception as e:      # noqa: BLE001
            why = f"{type(e).__name__}: {e}"[:200]
    if why is not None:
        from oracle import oracle as O
        O.set_num_threads(threads)
        cpu = CpuSample(meshes, args.samples, args.rays, args.seed)
        kind, cores = "port", O.num_threads()
    log(f"[reference] kind={kind} prep {cpu.prep_s:.1f}s, {cores} threads, {cpu.rays_per_step} rays/step" + (f" (reference unavailable: {why})" if why else ""))
    for w in range(max(args.warmup, 1)):                       # the first step also JIT-compiles the Numba kernels
        cpu.step(w)
    t = time.perf_counter()
    for k in range(args.steps):
        cpu.step(args.warmup + k)
    dt = time.perf_counter() - t
    v = cpu.rays_per_step * args.steps / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (f64 ray generation)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": cpu.describe()},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu.describe()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if why:
        line["cpu_baseline"]["reference_unavailable"] = why
    emit(line)


# --------------------------------------------------------------------------------------------- GPU arm

L2_FLUSH_BYTES = 160 * 1024 * 1024      # > the 126 MB L2: written before every trace launch, on its stream, inside the timed region


def _time_steps(ctx, solve, steps):
    """CUDA-event time (ms per step) of ``steps`` iterations enqueued back to back (``solve.step(steps)``: pipelined
    over the context's two streams, one host round trip at the end); the L2 is evicted before every trace launch
    (Context.set_l2_flush is on while bench.py measures)."""
    ctx.synchronize()
    ctx.timer_start()
    solve.step(steps)
    return ctx.timer_stop() / steps


def _max_abs_diff(res, gold):
    err = 0.0
    for name, row in gold.items():
        for key in set(row) | set(res[name]):
            err = max(err, abs(res[name].get(key, 0.0) - row.get(key, 0.0)))
    return err


def parity_block(ctx, rank, world, sc, em, active, n_once, table, args):
    """Results independent of the GPU count, checked inside the driver-run bench at every N:
    the rank-summed C5 tally block after PARITY_ITERS fixed iterations through the sharded solve driver -- its crc32,
    whether it equals the block an unsharded solve of the same process produces (N > 1), and whether the crc equals the
    committed single-GPU value; golden solves of the reference through the sharded public API; a ray-split scene."""
    import zlib
    from raystrack_b200 import MatrixParams, main as M, synthetic, view_factor_matrix
    sys.path.insert(0, str(ROOT / "tests"))
    from scenes import scene_for
    K = PARITY_ITERS
    n = active.shape[0]
    ids = np.arange(n, dtype=np.int32)
    kw = dict(max_iters=K, min_iters=K, interval=1, tol_mode="stderr", tol=0.0, emit_sid=ids, min_sid=np.zeros(n, np.int32))
    tallies, iters, totals = M._solve_sharded(ctx, sc, em, list(range(n)), n_once, active, table, **kw)
    tallies = np.array(tallies, copy=True)                     # the result may be a view of the pinned staging area
    out = {"iters": K, "tally_crc32": f"{zlib.crc32(tallies.tobytes()) & 0xffffffff:08x}", "tally_sum": int(tallies.sum()),
           "rays": int(totals.sum())}
    ref_file = ROOT / "profiles" / "c5_tally_crc_r2.json"
    if ref_file.exists() and args.side == 20 and args.samples == 4 and args.rays == 64 and args.seed == 1:
        want = json.loads(ref_file.read_text())
        if int(want.get("iters", -1)) == K:
            out["crc_equals_committed_n1"] = out["tally_crc32"] == want["tally_crc32"]
    if world > 1:
        M._DIST_OVERRIDE = (0, 1)                              # the same solve, unsharded, on this rank's GPU
        try:
            t1, i1, r1 = M._solve_sharded(ctx, sc, em, list(range(n)), n_once, active, table, **kw)
            same = bool(np.array_equal(t1, tallies) and np.array_equal(i1, iters) and np.array_equal(r1, totals))
        finally:
            M._DIST_OVERRIDE = None
        flags = np.array([1 if same else 0], np.int64)
        from raystrack_b200 import dist as D
        D.allreduce_sum_([flags])
        out["equals_n1"] = bool(int(flags[0]) == world)
    # the reference's own results through the sharded public API
    gold = json.loads((ROOT / "tests" / "golden" / "solves.json").read_text())
    worst = 0.0
    cases = ("C2_canyon_ex01", "U3_urban_matrix_bvh")
    for case in cases:
        worst = max(worst, _max_abs_diff(view_factor_matrix(scene_for(case), MatrixParams(**gold[case]["params"])), gold[case]["result"]))
    out["golden_max_abs_err"] = worst
    out["golden_cases"] = list(cases)
    # one oversized emitter (the ground) -> ray-split path with its per-iteration all-reduce
    meshes = synthetic.urban_block(4, 4, 8, 0)
    prm = MatrixParams(samples=4, rays=32, seed=2, bvh="builtin", reciprocity=False, max_iters=12, min_iters=3, tol=5e-4)
    res = view_factor_matrix(meshes, prm)
    if world > 1:
        from raystrack_b200.prepared import PreparedSolver
        n4 = [int(e.n_cells * 32) for e in PreparedSolver(meshes).get_emitters(samples=4, rays=32, flip_faces=False)]
        out["ray_split_jobs_per_rank"] = sum(1 for j in M.plan_shards(list(range(len(meshes))), n4, world)[rank] if j[3])
        M._DIST_OVERRIDE = (0, 1)
        try:
            same = view_factor_matrix(meshes, prm) == res
        finally:
            M._DIST_OVERRIDE = None
        flags = np.array([1 if same else 0], np.int64)
        D.allreduce_sum_([flags])
        out["ray_split_equals_n1"] = bool(int(flags[0]) == world)
    return out


def terrain_block(ctx, with_cpu: bool):
    """A second workload for tree quality: terrain + small objects (synthetic.terrain_with_objects, 1 053 352 triangles of
    0.03 ... 50 m, non-planar emitters), closest-hit throughput of one iteration of all 748 emitters and -- against the
    oracle (median-split reference tree) on a CPU sample -- per-ray agreement."""
    from raystrack_b200 import _native, main as M, synthetic
    from raystrack_b200.prepared import PreparedSolver
    meshes = synthetic.terrain_with_objects()
    samples, rays, seed = 1, 16, 1
    ps = PreparedSolver(meshes)
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx)
    em = ps.get_device_emitters(samples=samples, rays=rays, flip_faces=False, ctx=ctx)
    ems = ps.get_emitter_summaries(samples=samples, rays=rays, flip_faces=False, ctx=ctx)
    n = len(meshes)
    active = M._surface_masks(ems, *ps.get_mesh_bounds())
    n_once = [int(e.n_cells * rays) for e in ems]
    ids = np.arange(n, dtype=np.int32)
    solve = _native.Solve(ctx, sc.native, em.native, ids, active, M._rotation_table(seed, n, 64), ids.copy(), max_iters=64, min_iters=64,
                          interval=1, tol_mode="stderr", tol=0.0, emit_sid=ids, min_sid=np.zeros(n, np.int32))
    solve.step(2)
    ms = _time_steps(ctx, solve, 10)
    solve.close()
    info = sc.info()
    out = {"workload": f"terrain + small objects: {n} meshes, {ps.total_faces} triangles, samples={samples} rays={rays} "
                       f"({int(sum(n_once))} rays/iteration)", "value": sum(n_once) / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms,
           "bvh": f"{info['n_nodes']} wide nodes, depth {info['depth']}, built in {info['build_us']/1e3:.1f} ms"}
    if with_cpu:
        from oracle import oracle as O
        t = time.time()
        S = O.OracleSolver(meshes)
        scene = S.scene(True)
        centers, extents = S.bounds()
        agree = tot = 0
        emit = list(range(0, 36, 7)) + list(range(36, n, 53))      # terrain tiles, boxes, spheres, slabs
        for i in emit:
            oe = O.prepare_emitters([meshes[i]], samples, rays, False)[0]
            cnt = min(oe.n_rays_once, 16384)
            cpg, cpd = O.rotation(seed, i, 1)
            act = O.surface_mask(i, oe, centers, extents)
            ro, rd = O.build_rays(oe, cpg, cpd, count=cnt)
            rh, rf = O.trace_firsthit(scene, ro, rd, act, i, 0)
            _, _, hit, front = _native.trace_rays(ctx, sc.native, em.native, i, act, i, 0, np.concatenate([cpg, cpd]), mode=0, n_rays=cnt,
                                                  want_rays=False)
            agree += int(np.sum((hit == rh) & (front == rf)))
            tot += cnt
        out["parity"] = {"per_ray_agreement": agree / tot, "rays_compared": tot, "emitters": len(emit),
                         "against": "oracle (reference median-split BVH, C port)", "oracle_s": round(time.time() - t, 1)}
    return out


def secondary_block(ctx, sc, em, active, n_once, table, rays_per_step):
    """The other kernels of the path on the same scene and step definition (one iteration of every emitter), CUDA-event
    timed over 5 back-to-back steps, L2 evicted before every trace: discrete-sky any-hit, dual (closest hit + any-hit flag from one walk), and the API-default
    reciprocity=True schedule (emitter i ignores meshes j <= i; the last emitter has no receivers)."""
    from raystrack_b200 import _native
    n = active.shape[0]
    ids = np.arange(n, dtype=np.int32)
    zeros = np.zeros(n, np.int32)
    out = {}
    common = dict(max_iters=64, min_iters=64, interval=1, tol_mode="stderr", tol=0.0)

    def run(name, solve, rays):
        solve.step(2)
        ms = _time_steps(ctx, solve, 5)
        out[name] = {"value": rays / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms}
        solve.close()

    run("sky_discrete_any_hit", _native.Solve(ctx, sc.native, em.native, ids, active, table, ids.copy(), sky=True, discrete=True, **common),
        rays_per_step)
    side = dict(max_iters=64, min_iters=64, tol=0.0, tol_mode="stderr", interval=1)
    run("dual_matrix_and_sky", _native.DualSolve(ctx, sc.native, em.native, ids, active, table, ids.copy(), ids, zeros, side, side, True),
        rays_per_step)
    has_recv = np.asarray([bool(active[i, i + 1:].any()) for i in range(n)], bool)
    rid = ids[has_recv]
    run("matrix_reciprocity_schedule", _native.Solve(ctx, sc.native, em.native, rid, active[rid], table, rid.copy(), emit_sid=rid,
                                                     min_sid=rid + 1, **common), int(sum(n_once[i] for i in rid)))
    return out


def run_ours(args):
    import torch
    from raystrack_b200 import MatrixParams, _native, dist as D, main as M, view_factor_matrix
    from raystrack_b200.prepared import PreparedSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, world = D.init_from_env("nccl") if world > 1 else (0, 1)     # torchrun's group: rendezvous + the driver's NCCL check
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if _native.device_count() <= 0:
        raise RuntimeError("bench.py needs a B200: no CUDA device visible")
    # N > 1: the context that owns the library's own NCCL communicator (rsk_comm_init; id handed over through the group)
    ctx = M._context() if world > 1 else _native.Context.for_device(local)
    if world > 1 and not D.native_comm_active():
        raise RuntimeError(f"library communicator unavailable: {D._NATIVE_FAILED}")

    meshes = build_scene(args.side)
    ps = PreparedSolver(meshes)
    t = time.time()
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx)            # raw meshes up, records + BVH built on the GPU
    em = ps.get_device_emitters(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx)
    ems = ps.get_emitter_summaries(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx)
    ctx.synchronize()
    prep_s = time.time() - t
    # the same once more on the warm context (pool, Halton tables and jitter grids in place): what every later call pays
    t = time.time()
    ps2 = PreparedSolver(meshes)
    ps2.get_device_scene(use_bvh=True, ctx=ctx)
    ps2.get_device_emitters(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx)
    ps2.get_emitter_summaries(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx)
    ctx.synchronize()
    ps2.clear_device_cache()
    del ps2
    t = time.time()                                             # ... and a third time, into the blocks the second pass freed
    ps3 = PreparedSolver(meshes)
    ps3.get_device_scene(use_bvh=True, ctx=ctx)
    ps3.get_device_emitters(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx)
    ps3.get_emitter_summaries(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx)
    ctx.synchronize()
    prep_warm_s = time.time() - t
    ps3.clear_device_cache()
    del ps3
    geometry_bytes = ps._geometry(ctx).h2d_bytes
    info = sc.info()
    n = len(meshes)
    centers, extents = ps.get_mesh_bounds()
    active = M._surface_masks(ems, centers, extents)
    n_once = [int(e.n_cells * args.rays) for e in ems]
    rays_per_step = int(sum(n_once))
    total_iters = max(args.warmup, 3) + 2 * args.steps + 2
    table = M._rotation_table(args.seed, n, max(total_iters, 64))
    # the plan the public call would make: emitters weighted by rays x measured cost per ray (N > 1)
    cost = None
    if world > 1:
        ids_all = np.arange(n, dtype=np.int32)
        cost = M._emitter_cost_per_ray(ctx, sc, em, list(range(n)), n_once, active, table, ids_all, np.zeros(n, np.int32), rank, world)
    plans = M.plan_shards(list(range(n)), n_once, world, cost_per_ray=cost)
    plan = plans[rank]
    ids = np.asarray([j[0] for j in plan], np.int32)
    ranges = np.asarray([[j[1], j[2]] for j in plan], np.int64).reshape(-1, 2)
    n_shared = sum(1 for j in plan if j[3])
    solve = _native.Solve(ctx, sc.native, em.native, ids, active[ids], table, ids.copy(), max_iters=total_iters,
                          min_iters=total_iters, interval=1, tol_mode="stderr", tol=0.0,
                          emit_sid=ids, min_sid=np.zeros(len(ids), np.int32), ray_range=ranges)
    ctx.set_l2_flush(L2_FLUSH_BYTES)        # from here on every trace launch is preceded by a 160 MB scratch write

    def one_step(trace_only_timer=None):
        if trace_only_timer is not None:
            ctx.timer_start()
        solve.enqueue_trace()
        if trace_only_timer is not None:
            trace_only_timer.append(ctx.timer_stop())
        if world > 1 and n_shared:
            solve.allreduce_iter_tallies(n_shared)              # NCCL on the iteration's stream, between trace and fold
        solve.enqueue_fold()

    sampler = ClockSampler(local, world) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        one_step()
    if sampler:
        sampler.sample_adjacent()                # multi-GPU: clock + throttle reasons while the warm-up steps execute
    ctx.synchronize()
    if world > 1:
        D.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    # the timed region: exactly K steps, enqueued back to back (iterations are pipelined over two streams inside the
    # library; statistics still fold in iteration order), CUDA events on the context's streams around all of them.
    # The clock sampler starts once everything is enqueued and runs while this thread waits for the GPU (ClockSampler
    # explains what it may query inside a multi-GPU region).
    ctx.timer_start()
    for _ in range(args.steps):
        one_step()
    if sampler:
        sampler.start()
    local_ms = ctx.timer_stop()
    if sampler:
        sampler.pause()
    launches = ctx.launch_count() - launches0
    torch.cuda.synchronize()
    if world > 1:
        D.barrier()
    total_ms = D.max_over_ranks(float(local_ms))
    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e9

    # dominant kernel alone (same stream, CUDA events around the trace launch only)
    trace_ms = []
    for _ in range(args.steps):
        ctx.synchronize()
        one_step(trace_ms)
    ctx.synchronize()
    clocks = sampler.stop() if sampler else None
    my_rays = int(sum(j[2] - j[1] for j in plan))
    trace_avg_ms = float(np.mean(trace_ms))
    solve.close()
    ctx.set_l2_flush(0)                     # the public-API calls below run as a user's would

    # ---- e2e through the public API with host buffers: the BASELINE config-#5 call (fixed iteration count so every
    # implementation traces identical rays).  Every timed call starts from the caller's mesh list alone: flattening,
    # upload, device-side preparation, BVH build, masks, the iterations, tally download and the result dict.
    old_log = M._log
    M._log = lambda msg: None

    def timed_call(iters):
        prm = MatrixParams(samples=args.samples, rays=args.rays, seed=args.seed, bvh="builtin", reciprocity=False,
                           max_iters=iters, min_iters=iters, tol=0.0)
        torch.cuda.synchronize()
        if world > 1:
            D.barrier()
        t0 = time.perf_counter()
        res = view_factor_matrix(meshes, prm)                  # the caller holds the result: freeing 600 k floats is not part of the call
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        phases.append({k: round(1e3 * v, 1) for k, v in M.LAST_TIMING.items()})
        del res
        return D.max_over_ranks(dt)

    e2e_times, e2e_single, phases = [], [], []
    plan_note = None
    parity = secondary = None
    try:
        if args.e2e_steps > 0:
            # warm-up of the public path: one call of each kind.  The first calls of a process grow the context's
            # memory pool (blocks freed on one stream are not reusable on another until the streams meet), which
            # costs 50-150 ms once (scripts/overlap_ab.py prints the per-call phases)
            timed_call(1)
            timed_call(args.e2e_iters)
            e2e_times = [timed_call(args.e2e_iters) for _ in range(args.e2e_steps)]
            cut, n_todo = M.LAST_PLAN.get("first_part", 0), M.LAST_PLAN.get("emitters", 0)
            plan_note = (f"emitters [0, {cut}) of {n_todo} solved first, their result rows built on a worker thread while the "
                         f"remaining {n_todo - cut} are traced (main._overlap_split)") if cut else None
            e2e_single = [timed_call(1) for _ in range(3)]
        if not args.no_parity:
            parity = parity_block(ctx, rank, world, sc, em, active, n_once, table, args)
        if world == 1 and not args.no_secondary:
            ctx.set_l2_flush(L2_FLUSH_BYTES)
            secondary = secondary_block(ctx, sc, em, active, n_once, table, rays_per_step)
            secondary["terrain_scene"] = terrain_block(ctx, with_cpu=not args.no_cpu)
            ctx.set_l2_flush(0)
    finally:
        M._log = old_log
    e2e_value = rays_per_step * args.e2e_iters / float(np.mean(e2e_times)) / 1e9 if e2e_times else None
    e2e_single_value = rays_per_step / float(np.mean(e2e_single)) / 1e9 if e2e_single else None
    h2d = geometry_bytes + n * n + table.nbytes                # vertices + faces + offsets, surf_active, rotations
    d2h = n * 2 * n * 8 + n * 12                               # rank-summed int64 tally block + iteration/ray counters
    comm = ctx.comm_info() if world > 1 else None

    if world > 1:
        D.barrier()
        D.shutdown_native()
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (f64 ray generation)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "one Monte-Carlo iteration of all 2001 emitters", "rays_per_step": rays_per_step,
                       "bvh": f"GPU LBVH -> 8-wide quantised, {info['n_nodes']} nodes, depth {info['depth']}, built in {info['build_us']/1e3:.1f} ms",
                       "l2": "evicted before every iteration: 160 MB scratch write on the trace stream, inside the timed region", "sharding": f"emitters over {world} GPU(s), {n_shared} ray-split" + ("" if cost is None else ", loads weighted by measured cost per ray"),
                       "collectives": None if comm is None else f"librsk_b200 rsk_comm (NCCL {comm['nccl_version']}, {comm['nranks']} ranks) on the kernel stream",
                       "upload_prepare_build_s": round(prep_s, 3), "upload_prepare_build_warm_s": round(prep_warm_s, 4)},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": f"view_factor_matrix(meshes, MatrixParams(samples=4, rays=64, bvh='builtin', reciprocity=False, "
                            f"min_iters=max_iters={args.e2e_iters}, tol=0)) from the mesh list alone: upload of vertices+faces, "
                            f"device-side preparation, GPU BVH build, {args.e2e_iters} iterations, tally download, result dict",
                    "ms_per_step": 1e3 * float(np.mean(e2e_times)) if e2e_times else None,
                    "ms_each_call": [round(1e3 * t, 1) for t in e2e_times],
                    "phases_ms_last_call": phases[args.e2e_steps + 1] if len(phases) > args.e2e_steps + 1 else None,
                    "warmup_calls": 2,
                    "two_part_solve": plan_note,
                    "single_iteration_call": {"value": e2e_single_value, "unit": UNIT,
                                              "ms": 1e3 * float(np.mean(e2e_single)) if e2e_single else None}}}
    if parity is not None:
        line["parity"] = parity
    if secondary is not None:
        line["secondary"] = secondary

    # ---- CPU oracle (C port, all host threads): baseline + algorithmic bytes per ray; per-ray parity on the sample (N=1)
    bpr_file = ROOT / "profiles" / "c5_bytes_per_ray.json"
    bytes_per_ray, bpr_src = FALLBACK_BYTES_PER_RAY, "SURVEY.md 8d"
    if bpr_file.exists():
        bytes_per_ray = float(json.loads(bpr_file.read_text())["bytes_per_ray"])
        bpr_src = "profiles/c5_bytes_per_ray.json"
    if not args.no_cpu:
        from oracle import oracle as O
        O.set_num_threads(host_threads())                        # torchrun exports OMP_NUM_THREADS=1
        cpu = CpuSample(meshes, args.samples, args.rays, args.seed)
        cpu.step(0)                                              # warm-up
        stats = np.zeros(4, np.int64)
        keep = {}
        t = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t < 10.0 and reps < 12):
            cpu.step(1 + reps, stats=stats if reps == 0 else None, keep=keep if reps == 0 else None)
            reps += 1
        dt = time.perf_counter() - t
        cpu_v = cpu.rays_per_step * reps / dt / 1e9
        n_in, n_leaf, n_tri_t, n_skip = (stats / cpu.rays_per_step).tolist()
        bytes_per_ray = 60 * n_in + 8 * n_leaf + 53 * n_tri_t + 5 * n_skip
        bpr_src = "oracle replay on the CPU sample (this run)"
        line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": cpu.describe(),
                                "visits_per_ray": {"inner": n_in, "leaf": n_leaf, "tri": n_tri_t, "skipped": n_skip}}
        if world == 1:
            # per-ray parity of the GPU path on the same sample (iteration 1), through the C-ABI per-ray hook
            agree = tot = 0
            for i, cnt in zip(cpu.emit, cpu.counts):
                cpg, cpd = O.rotation(args.seed, i, 1)
                _, _, hit, front = _native.trace_rays(ctx, sc.native, em.native, i, active[i], i, 0, np.concatenate([cpg, cpd]),
                                                      mode=0, n_rays=cnt, want_rays=False)
                hs, fr = keep[i]
                agree += int(np.sum((hit == hs) & (front == fr)))
                tot += cnt
            line.setdefault("parity", {}).update({"per_ray_agreement": agree / tot, "rays_compared": tot})
    # ---- the metric's second half: VF max abs error vs the reference, on the reference's own example config (C2:
    # examples/ex01 street canyon, golden result generated by the reference itself: tests/golden/solves.json)
    if world == 1:
        try:
            from raystrack_b200 import synthetic
            gold = json.loads((ROOT / "tests" / "golden" / "solves.json").read_text())["C2_canyon_ex01"]
            M._log = lambda msg: None
            res = view_factor_matrix(synthetic.street_canyon(), MatrixParams(**gold["params"]))
            M._log = old_log
            line["vf_max_abs_err"] = {"value": _max_abs_diff(res, gold["result"]),
                                      "config": "C2 street canyon, ex01 params (11 meshes, reciprocity, stderr tol 1e-4)",
                                      "against": "reference CPU result (tests/golden/solves.json), tolerance 1e-4"}
        except Exception as e:      # noqa: BLE001
            line["vf_max_abs_err"] = {"value": None, "error": str(e)[:200]}
    elif parity is not None:
        line["vf_max_abs_err"] = {"value": parity.get("golden_max_abs_err"), "config": "sharded: " + ", ".join(parity.get("golden_cases", [])),
                                  "against": "reference CPU results (tests/golden/solves.json), tolerance 1e-4"}
    peak, peak_src = measured_peaks()
    achieved = my_rays * bytes_per_ray / (trace_avg_ms * 1e-3) / 1e9
    line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "kernel": "rsk_trace_kernel<matrix,bvh>", "kernel_ms": trace_avg_ms, "bytes_per_ray": bytes_per_ray,
                        "bytes_per_ray_source": bpr_src, "peak_source": peak_src,
                        "note": "SURVEY 8(d) yardstick: algorithmic bytes of the REFERENCE's 2-wide layout over the HBM peak; the 80 MB scene is "
                                "L2-resident, so frac > 1 is expected and says nothing about saturation -- what bounds the kernel is in "
                                "issue / l1 / l2 below (ncu capture of the same kernel and scene)"}
    ncu_csv = ROOT / "profiles" / "trace_r2_ncu_raw.csv"
    if ncu_csv.exists():
        try:
            sys.path.insert(0, str(ROOT / "scripts"))
            import ncu_extract
            prof = ncu_extract.summary(ncu_csv, "rsk_trace_kernel", rays_per_step)
            line["roofline"].update({"traffic": prof["dram"]["bytes"], "issue": prof["issue"], "l1": prof["l1"], "l2": prof["l2"],
                                     "ncu_kernel_ms": prof["kernel_ms"], "registers_per_thread": prof["registers_per_thread"],
                                     "warps_active_pct": prof["warps_active_pct"],
                                     "source": "profiles/trace_r2_ncu_raw.csv via scripts/ncu_extract.py (ncu --set full, one launch over the "
                                               "whole C5 iteration: compare ncu_kernel_ms with kernel_ms at N=1)"})
        except Exception as e:      # noqa: BLE001
            line["roofline"]["ncu_error"] = str(e)[:200]
    # ---- the unmodified reference on this box, next to the numbers above (N=1): Numba CPU kernels and Numba-CUDA path
    if world == 1 and not args.no_numba:
        line["reference_baselines"] = run_numba_baselines(args)
        nb = line["reference_baselines"]
        for key in ("numba_cpu", "numba_cuda"):
            v = (nb.get(key) or {}).get("Grays_per_s")
            if v:
                nb[key]["speedup_of_value"] = value / v
    emit(line)


def run_numba_baselines(args) -> dict:
    """baseline/numba_baselines.py in a process of its own (Numba-CUDA brings its own CUDA context and JIT): the
    reference's Numba CPU kernels on the CPU sample and its view_factor_matrix(device='gpu') over the whole scene."""
    cmd = [sys.executable, str(ROOT / "baseline" / "numba_baselines.py"), "--side", str(args.side), "--iters", "2", "--cpu-seconds", "6"]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    t = time.time()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=args.numba_timeout, env=env)
        last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        out = json.loads(last[-1]) if last else {"unavailable": f"no output (rc {r.returncode}): {r.stderr[-300:]}"}
    except subprocess.TimeoutExpired:
        out = {"unavailable": f"timed out after {args.numba_timeout} s"}
    except Exception as e:      # noqa: BLE001
        out = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    out["wall_s"] = round(time.time() - t, 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=20, help="buildings per side (20 = the 1M-triangle C5 scene)")
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--rays", type=int, default=64)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=2, help="timed public-API calls (0 = skip the e2e leg)")
    ap.add_argument("--e2e-iters", type=int, default=40, help="iterations per public-API call (C5: min_iters=max_iters=40)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU oracle legs (profiling runs)")
    ap.add_argument("--no-numba", action="store_true", help="skip the reference's Numba CPU / Numba-CUDA legs (N=1)")
    ap.add_argument("--numba-timeout", type=int, default=420)
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-parity block")
    ap.add_argument("--no-secondary", action="store_true", help="skip the sky / dual / reciprocity kernel timings (N=1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
