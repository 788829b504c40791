#!/usr/bin/env python3
"""Reported baselines: the UNMODIFIED reference (pip-installed into baseline/_ref, git-ignored) on the C5 scene.

  * Numba-CUDA path: reference view_factor_matrix(device="gpu") on one B200, fixed iteration count, JIT warm-up and
    preparation excluded (PreparedSolver reused), rays/s over the whole solve call
  * Numba CPU path: reference build_rays + trace_cpu_bvh_firsthit on bench.py's bounded sample, all host threads

Writes profiles/numba_baselines_r1.json (only on a box that has baseline/_ref and a GPU).  Nothing of the product
is on this path."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.path.insert(0, str(REF))
sys.path.insert(1, str(ROOT))


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    out = {"scene_side": side, "iters": iters}
    if not (REF / "raystrack").exists():
        print(json.dumps({"unavailable": "baseline/_ref missing"}))
        return
    import numba
    import raystrack
    import raystrack.main as ref_main
    from raystrack import MatrixParams, PreparedSolver, view_factor_matrix
    from raystrack.utils import cpu_trace, ray_builder
    from raystrack_b200 import synthetic

    assert str(REF) in raystrack.__file__, raystrack.__file__
    ref_main._log = lambda msg: None
    meshes = synthetic.urban_block(side)
    out["n_meshes"] = len(meshes)
    t = time.time()
    ps = PreparedSolver(meshes)
    ems = ps.get_emitters(samples=4, rays=64, flip_faces=False)
    scene = ps.get_scene(use_bvh=True)
    out["reference_prep_s"] = round(time.time() - t, 1)
    out["numba"] = numba.__version__
    rays_iter = int(sum(e.n_cells * 64 for e in ems))
    print("prep done", out, flush=True)

    # ---- Numba CPU path on the bounded sample (bench.py: every 16th emitter, first 16384 rays)
    centers, extents = ps.get_mesh_bounds()
    emit = list(range(0, len(meshes), 16))

    def cpu_step(itr):
        n_tot = 0
        for i in emit:
            em = ems[i]
            n = min(16384, em.n_cells * 64)
            rng = np.random.default_rng(1 + i + itr)
            cpg = rng.random(2, dtype=np.float32)
            cpd = rng.random(5, dtype=np.float32)
            o = np.empty((n, 3), np.float32)
            d = np.empty_like(o)
            ray_builder.build_rays(em.u_grid, em.v_grid, em.halton_tri[:n], em.halton_u[:n], em.halton_v[:n], em.halton_r1[:n],
                                   em.halton_r2[:n], em.cdf, em.tri_a, em.tri_e1, em.tri_e2, em.tri_u, em.tri_v, em.tri_n,
                                   em.tri_origin_eps, 64, o, d, cpg, cpd)
            act = ref_main._build_emitter_surface_mask(i, em, centers, extents)
            hs = np.empty(n, np.int32)
            fr = np.empty(n, np.uint8)
            cpu_trace.trace_cpu_bvh_firsthit(o, d, scene.v0, scene.e1, scene.e2, scene.normals, scene.sid, act, scene.bb_min,
                                             scene.bb_max, scene.left, scene.right, scene.start, scene.count, i, 0, hs, fr)
            n_tot += n
        return n_tot

    cpu_step(0)     # JIT + warm-up
    t = time.perf_counter()
    reps, n_rays = 0, 0
    while reps < 3 or time.perf_counter() - t < 10:
        n_rays += cpu_step(1 + reps)
        reps += 1
    dt = time.perf_counter() - t
    out["numba_cpu"] = {"Grays_per_s": n_rays / dt / 1e9, "threads": numba.config.NUMBA_NUM_THREADS, "rays": n_rays,
                        "sample": "every 16th emitter, first 16384 rays, build_rays + trace_cpu_bvh_firsthit"}
    print("cpu", out["numba_cpu"], flush=True)

    # ---- Numba-CUDA path through the reference's public API
    try:
        from numba import cuda
        if not cuda.is_available():
            raise RuntimeError("numba.cuda.is_available() is False")
        prm = MatrixParams(samples=4, rays=64, seed=1, bvh="builtin", device="gpu", reciprocity=False, max_iters=1, min_iters=1, tol=0.0)
        t = time.perf_counter()
        view_factor_matrix(meshes, prm, prepared=ps)                # JIT compile + uploads
        out["numba_cuda_first_call_s"] = round(time.perf_counter() - t, 1)
        prm = MatrixParams(samples=4, rays=64, seed=1, bvh="builtin", device="gpu", reciprocity=False, max_iters=iters, min_iters=iters, tol=0.0)
        t = time.perf_counter()
        view_factor_matrix(meshes, prm, prepared=ps)
        cuda.synchronize()
        dt = time.perf_counter() - t
        out["numba_cuda"] = {"Grays_per_s": rays_iter * iters / dt / 1e9, "seconds": dt, "rays": rays_iter * iters,
                             "call": f"view_factor_matrix(device='gpu', max_iters=min_iters={iters}), PreparedSolver reused"}
    except Exception as e:      # noqa: BLE001
        out["numba_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    print("cuda", out["numba_cuda"], flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "numba_baselines_r1.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
