#!/usr/bin/env python3
"""Reported baselines: the UNMODIFIED reference (pip-installed into baseline/_ref, git-ignored, travels to the GPU box)
on the C5 scene.  Nothing of the product is on this path except the synthetic scene generator (mesh lists).

  * Numba CPU path: the reference's own ``build_rays`` + ``trace_cpu_bvh_firsthit`` (utils/ray_builder.py:25-94,
    utils/cpu_trace.py:120-277) with the reference's own preparation (``prepare_scene`` / ``build_bvh`` /
    ``prepare_emitters``) on bench.py's bounded sample, all host threads.  ``NumbaCpuSample`` is what
    ``bench.py --impl reference`` times.
  * Numba-CUDA path: the reference's ``view_factor_matrix(device="gpu")`` (main.py:1025) on one B200 over the whole
    scene, fixed iteration count, JIT warm-up and preparation excluded (PreparedSolver reused).

CLI (run by bench.py as a subprocess at N=1, and by hand):
    python baseline/numba_baselines.py [--side 20] [--iters 1] [--cpu-seconds 8] [--no-cuda] [--out FILE]
prints one JSON object on the last stdout line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
SAMPLE_STRIDE = 16          # keep equal to bench.py: every 16th emitter ...
SAMPLE_RAYS = 16384         # ... first 16384 rays of each


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def prepare_environment(threads: int | None = None) -> int:
    """Must run before numba is imported: torchrun exports OMP_NUM_THREADS=1, which would pin Numba's OpenMP
    threading layer (and the oracle's OpenMP loops) to one thread."""
    n = int(threads or host_threads())
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["NUMBA_NUM_THREADS"] = str(n)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
    return n


def reference_available() -> str | None:
    """None when the reference can be imported from baseline/_ref, else the reason."""
    if not (REF / "raystrack").exists():
        return "baseline/_ref/raystrack missing (pip install --target baseline/_ref of the reference not done)"
    try:
        import numba  # noqa: F401
    except Exception as e:      # noqa: BLE001
        return f"numba not importable: {e}"
    return None


def import_reference():
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    if str(ROOT) not in sys.path:
        sys.path.insert(1, str(ROOT))
    import raystrack
    assert str(REF) in raystrack.__file__, raystrack.__file__
    import raystrack.main as ref_main
    ref_main._log = lambda msg: None
    return raystrack, ref_main


class NumbaCpuSample:
    """The reference's CPU kernels on the bounded sample: every SAMPLE_STRIDE-th emitter, first SAMPLE_RAYS rays.
    Preparation is the reference's own: ``prepare_scene(meshes, use_bvh=True)`` (median-split BVH, 20 s at C5) and
    ``prepare_emitters`` on the sampled meshes only (per-mesh function; the unsampled meshes' emitter records are
    never read)."""

    def __init__(self, meshes, samples: int, rays: int, seed: int):
        import numpy as np
        _, ref_main = import_reference()
        import numba
        from raystrack.utils import cpu_trace, prepared as ref_prepared, ray_builder
        self.np, self.ref_main, self.cpu_trace, self.ray_builder = np, ref_main, cpu_trace, ray_builder
        self.rays, self.seed = int(rays), int(seed)
        t = time.time()
        self.emit = list(range(0, len(meshes), SAMPLE_STRIDE))
        self.ems = ref_prepared.prepare_emitters([meshes[i] for i in self.emit], samples=samples, rays=rays, flip_faces=False)
        self.scene = ref_prepared.prepare_scene(meshes, use_bvh=True)
        solver = ref_prepared.PreparedSolver(meshes)
        self.centers, self.extents = solver.get_mesh_bounds()
        self.counts = [min(SAMPLE_RAYS, em.n_cells * self.rays) for em in self.ems]
        self.masks = [ref_main._build_emitter_surface_mask(i, em, self.centers, self.extents) for i, em in zip(self.emit, self.ems)]
        self.prep_s = time.time() - t
        self.rays_per_step = int(sum(self.counts))
        self.threads = int(numba.config.NUMBA_NUM_THREADS)
        self.numba_version = numba.__version__

    def step(self, itr: int) -> int:
        np, sc = self.np, self.scene
        for i, em, n, act in zip(self.emit, self.ems, self.counts, self.masks):
            rng = np.random.default_rng(self.seed + i + itr)                # main.py:1810-1812
            cpg = rng.random(2, dtype=np.float32)
            cpd = rng.random(5, dtype=np.float32)
            o = np.empty((n, 3), np.float32)
            d = np.empty_like(o)
            self.ray_builder.build_rays(em.u_grid, em.v_grid, em.halton_tri[:n], em.halton_u[:n], em.halton_v[:n], em.halton_r1[:n],
                                        em.halton_r2[:n], em.cdf, em.tri_a, em.tri_e1, em.tri_e2, em.tri_u, em.tri_v, em.tri_n,
                                        em.tri_origin_eps, self.rays, o, d, cpg, cpd)
            hs = np.empty(n, np.int32)
            fr = np.empty(n, np.uint8)
            self.cpu_trace.trace_cpu_bvh_firsthit(o, d, sc.v0, sc.e1, sc.e2, sc.normals, sc.sid, act, sc.bb_min, sc.bb_max, sc.left,
                                                  sc.right, sc.start, sc.count, i, 0, hs, fr)
        return self.rays_per_step

    def describe(self) -> str:
        return (f"every {SAMPLE_STRIDE}th emitter ({len(self.emit)} incl. ground), first {SAMPLE_RAYS} rays of each = "
                f"{self.rays_per_step} rays/step; reference prepare_scene/build_bvh/prepare_emitters + build_rays + "
                f"trace_cpu_bvh_firsthit (Numba {self.numba_version}, {self.threads} threads)")


def numba_cuda_leg(meshes, samples: int, rays: int, seed: int, iters: int) -> dict:
    """The reference's public call on its Numba-CUDA path over the whole scene."""
    raystrack, _ = import_reference()
    from numba import cuda
    if not cuda.is_available():
        return {"unavailable": "numba.cuda.is_available() is False"}
    out: dict = {}
    t = time.time()
    ps = raystrack.PreparedSolver(meshes)
    ems = ps.get_emitters(samples=samples, rays=rays, flip_faces=False)
    ps.get_scene(use_bvh=True)
    out["reference_prep_s"] = round(time.time() - t, 1)
    rays_iter = int(sum(e.n_cells * rays for e in ems))

    def call(n_it):
        prm = raystrack.MatrixParams(samples=samples, rays=rays, seed=seed, bvh="builtin", device="gpu", reciprocity=False,
                                     max_iters=n_it, min_iters=n_it, tol=0.0)
        t0 = time.perf_counter()
        raystrack.view_factor_matrix(meshes, prm, prepared=ps)
        cuda.synchronize()
        return time.perf_counter() - t0

    out["first_call_s"] = round(call(1), 1)                    # JIT compile + uploads
    dt = call(iters)
    out.update({"Grays_per_s": rays_iter * iters / dt / 1e9, "seconds": dt, "rays": rays_iter * iters,
                "call": f"view_factor_matrix(device='gpu', bvh='builtin', reciprocity=False, max_iters=min_iters={iters}, tol=0), "
                        f"PreparedSolver reused, whole scene"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=20)
    ap.add_argument("--iters", type=int, default=1, help="iterations of the timed Numba-CUDA call")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    threads = prepare_environment()
    why = reference_available()
    if why:
        print(json.dumps({"unavailable": why}))
        return
    sys.path.insert(1, str(ROOT))
    from raystrack_b200 import synthetic
    out: dict = {"scene_side": args.side, "host_threads": threads}
    meshes = synthetic.urban_block(args.side)
    out["n_meshes"] = len(meshes)
    if not args.no_cpu:
        try:
            cpu = NumbaCpuSample(meshes, 4, 64, 1)
            cpu.step(0)                                         # JIT + warm-up
            t = time.perf_counter()
            reps = n_rays = 0
            while reps < 3 or time.perf_counter() - t < args.cpu_seconds:
                n_rays += cpu.step(1 + reps)
                reps += 1
            dt = time.perf_counter() - t
            out["numba_cpu"] = {"Grays_per_s": n_rays / dt / 1e9, "threads": cpu.threads, "rays": n_rays, "steps": reps,
                                "prep_s": round(cpu.prep_s, 1), "sample": cpu.describe()}
        except Exception as e:      # noqa: BLE001
            out["numba_cpu"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        print("numba_cpu", out["numba_cpu"], file=sys.stderr, flush=True)
    if not args.no_cuda:
        try:
            out["numba_cuda"] = numba_cuda_leg(meshes, 4, 64, 1, args.iters)
        except Exception as e:      # noqa: BLE001
            out["numba_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        print("numba_cuda", out["numba_cuda"], file=sys.stderr, flush=True)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
