#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by IMPORTING THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden.py

Outputs (committed):
  stage_vectors.npz   per-stage vectors: Halton tables, emitter preparation, rays, BVH arrays, per-ray hits,
                      hit masks, Tregenza patch ids  (reference functions called directly)
  solves.json         whole-solve results + per-emitter iteration counts for the BASELINE configs C1-C4,
                      two validation cases and the sky variants (reference public API, device="cpu")
  solves_extra.json   further parameter combinations (row-sum enforcement, delta sky, flipped enclosure with BVH, tilted meshes)
  workflow.json       view_factor_outside_workflow / view_factor_matrix_and_sky results (reference api.py, main.py:1209)
  shipped.json        the result files the reference ships (examples/*.json, validation/results/*), verbatim data
  large_rays.npz      per-ray closest hits of the reference's BVH tracer on a 127 488-triangle urban block (7x7 buildings,
                      16x16 face grids; reference tree: 32 767 nodes) for a wall, a roof, an inner wall and the ground
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
REF = Path("/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REPO))

import raystrack  # noqa: E402  (the reference)
import raystrack.main as ref_main  # noqa: E402
from raystrack import MatrixParams, SkyParams, view_factor_matrix, view_factor_to_tregenza_sky  # noqa: E402
from raystrack.utils import cpu_trace, halton, prepared, ray_builder  # noqa: E402
from raystrack.utils.bvh import build_bvh  # noqa: E402
from raystrack.io import load_meshes_json  # noqa: E402

from raystrack_b200 import synthetic  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_logged(fn, meshes, params):
    logs = []
    old = ref_main._log
    ref_main._log = logs.append
    try:
        out = fn(meshes, params)
    finally:
        ref_main._log = old
    iters = {}
    pat = re.compile(r"\[\s*(?P<name>[^\]]+?)\s*\]\s+(?P<iters>\d+)\s+iter")
    for m in logs:
        mm = pat.search(m)
        if mm:
            iters[mm.group("name")] = int(mm.group("iters"))
    out = {k: {kk: float(vv) for kk, vv in row.items()} for k, row in out.items()}
    return out, iters


def stage_vectors():
    z = {}
    # --- Halton (utils/halton.py)
    dims = halton.cached_halton_dims(70000)
    z["halton_dims_head"] = np.stack([d[:4096] for d in dims])
    z["halton_dims_tail"] = np.stack([d[-512:] for d in dims])
    big = 45158400  # C5 ground emitter: 840^2 * 64
    probe = np.array([0, 1, 2, 12345, 999999, 4194303, 4194304, 33554431, big - 1], np.int64)
    z["halton_probe_idx"] = probe
    z["halton_probe_val"] = np.array(
        [[np.float32(halton._halton(int(i) + 1, b)) for i in probe] for b in (5, 2, 3, 7, 11)], np.float32)
    for g in (4, 16, 26):
        u, v = halton.cached_halton(g)
        z[f"grid_u_{g}"] = u
        z[f"grid_v_{g}"] = v

    # --- emitter preparation (utils/prepared.py) on tilted + canyon meshes
    for tag, meshes, flip in (("tilted", synthetic.tilted_pair(), False), ("tiltedflip", synthetic.tilted_pair(), True),
                              ("canyon", synthetic.street_canyon(), False)):
        ems = prepared.prepare_emitters(meshes, samples=16, rays=8, flip_faces=flip)
        for i, em in enumerate(ems):
            for f in ("tri_a", "tri_e1", "tri_e2", "tri_u", "tri_v", "tri_n", "tri_origin_eps", "cdf",
                      "plane_origin", "plane_normal"):
                z[f"em_{tag}_{i}_{f}"] = getattr(em, f)
            z[f"em_{tag}_{i}_scalars"] = np.array([em.total_area, em.plane_tol, float(em.plane_is_planar), em.g], np.float64)

    # --- rays (utils/ray_builder.py) for the tilted pair and one canyon wall
    def rays_for(meshes, idx, samples, rays, seed_sum, flip=False):
        em = prepared.prepare_emitters(meshes, samples=samples, rays=rays, flip_faces=flip)[idx]
        rng = np.random.default_rng(seed_sum)
        cpg = rng.random(2, dtype=np.float32)
        cpd = rng.random(5, dtype=np.float32)
        n = em.n_cells * rays
        o = np.empty((n, 3), np.float32)
        d = np.empty_like(o)
        ray_builder.build_rays(em.u_grid, em.v_grid, em.halton_tri, em.halton_u, em.halton_v, em.halton_r1, em.halton_r2,
                               em.cdf, em.tri_a, em.tri_e1, em.tri_e2, em.tri_u, em.tri_v, em.tri_n, em.tri_origin_eps,
                               rays, o, d, cpg, cpd)
        return em, cpg, cpd, o, d

    em, cpg, cpd, o, d = rays_for(synthetic.tilted_pair(), 0, 64, 16, 7)
    z["rays_tilted_cp"] = np.concatenate([cpg, cpd])
    z["rays_tilted_orig"] = o
    z["rays_tilted_dir"] = d
    em, cpg, cpd, o, d = rays_for(synthetic.street_canyon(), 10, 16, 128, 11)   # road, 36^2*128 rays
    z["rays_road_cp"] = np.concatenate([cpg, cpd])
    z["rays_road_sha"] = np.frombuffer(bytes.fromhex(sha(o) + sha(d)), np.uint8)
    z["rays_road_orig_head"] = o[:2048]
    z["rays_road_dir_head"] = d[:2048]

    # --- BVH + per-ray hits on a mini urban block (3x3 buildings, 4x4 face grids, 8x8 ground: 848 triangles)
    meshes = synthetic.urban_block(n_side=3, face_grid=4, ground_grid=8, seed=0)
    scene_b = prepared.prepare_scene(meshes, use_bvh=True)
    scene_f = prepared.prepare_scene(meshes, use_bvh=False)
    for f in ("bb_min", "bb_max", "left", "right", "start", "count", "sid"):
        z[f"bvh_{f}"] = getattr(scene_b, f)
    z["bvh_v0_sha"] = np.frombuffer(bytes.fromhex(sha(scene_b.v0)), np.uint8)
    solver = prepared.PreparedSolver(meshes)
    centers, extents = solver.get_mesh_bounds()
    ems = solver.get_emitters(samples=4, rays=16, flip_faces=False)
    n_surf = len(meshes)
    for idx, recip in ((0, False), (7, True), (n_surf - 1, False), (22, False)):
        em = ems[idx]
        rng = np.random.default_rng(100 + idx)
        cpg = rng.random(2, dtype=np.float32)
        cpd = rng.random(5, dtype=np.float32)
        n = em.n_cells * 16
        o = np.empty((n, 3), np.float32)
        d = np.empty_like(o)
        ray_builder.build_rays(em.u_grid, em.v_grid, em.halton_tri, em.halton_u, em.halton_v, em.halton_r1, em.halton_r2,
                               em.cdf, em.tri_a, em.tri_e1, em.tri_e2, em.tri_u, em.tri_v, em.tri_n, em.tri_origin_eps,
                               16, o, d, cpg, cpd)
        active = ref_main._build_emitter_surface_mask(idx, em, centers, extents)
        emit_sid, min_sid = ref_main._matrix_skip(idx, recip)
        hs = np.empty(n, np.int32)
        fr = np.empty(n, np.uint8)
        cpu_trace.trace_cpu_bvh_firsthit(o, d, scene_b.v0, scene_b.e1, scene_b.e2, scene_b.normals, scene_b.sid, active,
                                         scene_b.bb_min, scene_b.bb_max, scene_b.left, scene_b.right, scene_b.start,
                                         scene_b.count, emit_sid, min_sid, hs, fr)
        hs2 = np.empty(n, np.int32)
        fr2 = np.empty(n, np.uint8)
        cpu_trace.trace_cpu_firsthit(o, d, scene_f.v0, scene_f.e1, scene_f.e2, scene_f.normals, scene_f.sid, active,
                                     emit_sid, min_sid, hs2, fr2)
        hm = np.empty(n, np.uint8)
        cpu_trace.trace_cpu_bvh_hitmask(o, d, scene_b.v0, scene_b.e1, scene_b.e2, scene_b.sid, active,
                                        scene_b.bb_min, scene_b.bb_max, scene_b.left, scene_b.right, scene_b.start,
                                        scene_b.count, idx, 0, hm)
        hm2 = np.empty(n, np.uint8)
        cpu_trace.trace_cpu_hitmask(o, d, scene_f.v0, scene_f.e1, scene_f.e2, scene_f.sid, active, idx, 0, hm2)
        cs, cf, ca = (np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.uint8))
        cpu_trace.trace_cpu_bvh_combined(o, d, scene_b.v0, scene_b.e1, scene_b.e2, scene_b.normals, scene_b.sid, active,
                                         scene_b.bb_min, scene_b.bb_max, scene_b.left, scene_b.right, scene_b.start,
                                         scene_b.count, idx, min_sid, cs, cf, ca)
        counts = np.empty(145, np.int64)
        cpu_trace.bin_tregenza_cpu(d, hm, counts)
        k = f"urb_{idx}_{int(recip)}"
        z[k + "_cp"] = np.concatenate([cpg, cpd])
        z[k + "_active"] = active
        z[k + "_hit_bvh"] = hs
        z[k + "_front_bvh"] = fr
        z[k + "_hit_brute"] = hs2
        z[k + "_front_brute"] = fr2
        z[k + "_mask_bvh"] = hm
        z[k + "_mask_brute"] = hm2
        z[k + "_comb_hit"] = cs
        z[k + "_comb_front"] = cf
        z[k + "_comb_any"] = ca
        z[k + "_tregenza"] = counts
        z[k + "_upward"] = np.array([cpu_trace.count_upward_misses_cpu(d, hm)], np.int64)

    # --- Tregenza patch ids on a deterministic direction set (incl. ring / azimuth boundaries)
    rng = np.random.default_rng(5)
    dirs = rng.normal(size=(20000, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    extra = []
    for el in range(0, 91, 6):
        for az in range(0, 360, 3):
            e, a = np.radians(el), np.radians(az)
            extra.append([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)])
    dirs = np.concatenate([dirs, np.asarray(extra, np.float32)]).astype(np.float32)
    import numba as nb

    @nb.njit
    def patch_ids(dd):
        out = np.empty(dd.shape[0], np.int32)
        for i in range(dd.shape[0]):
            out[i] = cpu_trace._tregenza_patch_id(dd[i, 0], dd[i, 1], dd[i, 2])
        return out

    z["treg_dirs"] = dirs
    z["treg_ids"] = patch_ids(dirs)
    np.savez_compressed(HERE / "stage_vectors.npz", **z)
    print("stage_vectors.npz:", len(z), "arrays")


LARGE_SCENE = dict(n_side=7, face_grid=16, ground_grid=32, seed=0)     # 246 meshes, 127 488 triangles
LARGE_EMITTERS = (0, 4, 122, 245)                                      # edge wall, roof, inner wall, ground
LARGE_RAYS = 8192


def large_rays():
    """A real-size tree pinned to the reference: trace_cpu_bvh_firsthit (utils/cpu_trace.py:120-277) through the
    reference's own median-split BVH (utils/bvh.py) over 127 488 triangles, first LARGE_RAYS rays of four emitters."""
    meshes = synthetic.urban_block(**LARGE_SCENE)
    scene = prepared.prepare_scene(meshes, use_bvh=True)
    centers, extents = prepared.PreparedSolver(meshes).get_mesh_bounds()
    z = {"n_tri": np.array([scene.v0.shape[0]], np.int64), "n_nodes": np.array([scene.left.shape[0]], np.int64)}
    samples, rays = 4, 16
    for idx in LARGE_EMITTERS:
        em = prepared.prepare_emitters([meshes[idx]], samples=samples, rays=rays, flip_faces=False)[0]
        rng = np.random.default_rng(500 + idx)
        cpg = rng.random(2, dtype=np.float32)
        cpd = rng.random(5, dtype=np.float32)
        n = min(LARGE_RAYS, em.n_cells * rays)
        o = np.empty((n, 3), np.float32)
        d = np.empty_like(o)
        ray_builder.build_rays(em.u_grid, em.v_grid, em.halton_tri[:n], em.halton_u[:n], em.halton_v[:n], em.halton_r1[:n],
                               em.halton_r2[:n], em.cdf, em.tri_a, em.tri_e1, em.tri_e2, em.tri_u, em.tri_v, em.tri_n,
                               em.tri_origin_eps, rays, o, d, cpg, cpd)
        active = ref_main._build_emitter_surface_mask(idx, em, centers, extents)
        hs = np.empty(n, np.int32)
        fr = np.empty(n, np.uint8)
        cpu_trace.trace_cpu_bvh_firsthit(o, d, scene.v0, scene.e1, scene.e2, scene.normals, scene.sid, active, scene.bb_min,
                                         scene.bb_max, scene.left, scene.right, scene.start, scene.count, idx, 0, hs, fr)
        hm = np.empty(n, np.uint8)
        cpu_trace.trace_cpu_bvh_hitmask(o, d, scene.v0, scene.e1, scene.e2, scene.sid, active, scene.bb_min, scene.bb_max,
                                        scene.left, scene.right, scene.start, scene.count, idx, 0, hm)
        z[f"e{idx}_cp"] = np.concatenate([cpg, cpd])
        z[f"e{idx}_active"] = active.astype(np.uint8)
        z[f"e{idx}_hit"] = hs.astype(np.int16)
        z[f"e{idx}_front"] = fr
        z[f"e{idx}_mask"] = hm
        z[f"e{idx}_rays_sha"] = np.frombuffer(bytes.fromhex(sha(o) + sha(d)), np.uint8)
        print(f"large emitter {idx}: {n} rays, hit fraction {float(np.mean(hs >= 0)):.3f}, receivers {len(np.unique(hs))}")
    # --- a zero-area emitter (collinear corners): all-zero QMC tables, cdf of ones (utils/prepared.py:278-287)
    Vz = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]], np.float32)
    Fz = np.array([[0, 1, 2], [1, 2, 3]], np.int32)
    em = prepared.prepare_emitters([("line", Vz, Fz)], samples=4, rays=8, flip_faces=False)[0]
    assert em.total_area <= 0.0 and em.g == 4 and not em.halton_tri.any()
    rng = np.random.default_rng(77)
    cpg = rng.random(2, dtype=np.float32)
    cpd = rng.random(5, dtype=np.float32)
    n = em.n_cells * 8
    o = np.empty((n, 3), np.float32)
    d = np.empty_like(o)
    ray_builder.build_rays(em.u_grid, em.v_grid, em.halton_tri, em.halton_u, em.halton_v, em.halton_r1, em.halton_r2, em.cdf, em.tri_a,
                           em.tri_e1, em.tri_e2, em.tri_u, em.tri_v, em.tri_n, em.tri_origin_eps, 8, o, d, cpg, cpd)
    z["zero_area_cp"] = np.concatenate([cpg, cpd])
    z["zero_area_orig"] = o
    z["zero_area_dir"] = d
    np.savez_compressed(HERE / "large_rays.npz", **z)
    print("large_rays.npz", (HERE / "large_rays.npz").stat().st_size, "bytes")


def solves():
    out = {}
    canyon = load_meshes_json(str(REF / "examples" / "street_canyon.json"))
    mine = synthetic.street_canyon()
    assert all(a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) for a, b in zip(canyon, mine))
    cube = synthetic.unit_cube_enclosure()

    def add(name, fn, meshes, params):
        res, iters = run_logged(fn, meshes, params)
        out[name] = {"params": params.as_dict(), "result": res, "iters": iters}
        print(name, "done", iters)

    add("C1_readme_squares", view_factor_matrix, synthetic.parallel_unit_squares(),
        MatrixParams(samples=256, rays=256, bvh="builtin", reciprocity=True, device="cpu"))
    add("C2_canyon_ex01", view_factor_matrix, canyon,
        MatrixParams(samples=16, rays=128, seed=1, bvh="auto", device="cpu", max_iters=200, tol=1e-4, tol_mode="stderr",
                     min_iters=10, reciprocity=True, enforce_reciprocity_rowsum=False, cuda_async=True))
    add("C2b_canyon_delta_norecip", view_factor_matrix, canyon,
        MatrixParams(samples=8, rays=64, seed=3, bvh="builtin", device="cpu", max_iters=60, tol=2e-4, tol_mode="delta",
                     min_iters=5, reciprocity=False))
    add("C3_canyon_sky_discrete", view_factor_to_tregenza_sky, canyon,
        SkyParams(samples=32, rays=256, discrete=True, device="cpu"))
    add("C3b_canyon_sky_merged", view_factor_to_tregenza_sky, canyon,
        SkyParams(samples=16, rays=64, discrete=False, device="cpu", seed=5, min_iters=8))
    add("C4_cube_ex04", view_factor_matrix, cube,
        MatrixParams(samples=16, rays=128, seed=42, bvh="auto", device="cpu", flip_faces=True, reciprocity=False,
                     enforce_reciprocity_rowsum=False, max_iters=1000, tol=1e-3, tol_mode="stderr", min_iters=10,
                     cuda_async=True))
    add("V06_canyon_view3d", view_factor_matrix, canyon,
        MatrixParams(samples=8, rays=512, seed=31, bvh="builtin", device="cpu", cuda_async=False, gpu_raygen=False,
                     max_iters=500, min_iters=40, tol=1e-4, tol_mode="stderr", convergence_interval=1, reciprocity=False))
    # validation 04 geometry (validate_04_patch_to_disc.py): restated small variant is produced by tests; here the mini urban
    urb = synthetic.urban_block(n_side=3, face_grid=4, ground_grid=8, seed=0)
    add("U3_urban_matrix_bvh", view_factor_matrix, urb,
        MatrixParams(samples=2, rays=32, seed=9, bvh="builtin", device="cpu", max_iters=12, min_iters=12, tol=0.0,
                     reciprocity=False))
    add("U3_urban_matrix_recip", view_factor_matrix, urb,
        MatrixParams(samples=2, rays=32, seed=9, bvh="auto", device="cpu", max_iters=30, min_iters=5, tol=2e-3,
                     reciprocity=True))
    add("U3_urban_sky", view_factor_to_tregenza_sky, urb,
        SkyParams(samples=2, rays=32, seed=9, bvh="builtin", device="cpu", max_iters=10, min_iters=10, tol=0.0, discrete=True))
    (HERE / "solves.json").write_text(json.dumps(out, indent=1, sort_keys=True))


def solves_extra():
    """More parameter combinations of the public solves (written to solves_extra.json; same format as solves.json)."""
    out = {}
    canyon = synthetic.street_canyon()
    cube = synthetic.unit_cube_enclosure()
    urb = synthetic.urban_block(n_side=3, face_grid=4, ground_grid=8, seed=0)
    tilted = synthetic.tilted_pair()

    def add(name, fn, meshes, params):
        res, iters = run_logged(fn, meshes, params)
        out[name] = {"params": params.as_dict(), "result": res, "iters": iters}
        print(name, "done", iters)

    add("X1_canyon_rowsum", view_factor_matrix, canyon,
        MatrixParams(samples=8, rays=64, seed=11, bvh="off", device="cpu", max_iters=40, min_iters=6, tol=5e-4, reciprocity=True,
                     enforce_reciprocity_rowsum=True))
    add("X2_canyon_sky_delta", view_factor_to_tregenza_sky, canyon,
        SkyParams(samples=8, rays=64, seed=4, bvh="builtin", device="cpu", max_iters=50, min_iters=4, tol=1e-3, tol_mode="delta",
                  discrete=True))
    add("X3_cube_flip_bvh", view_factor_matrix, cube,
        MatrixParams(samples=16, rays=64, seed=2, bvh="builtin", device="cpu", flip_faces=True, reciprocity=True, max_iters=60,
                     min_iters=5, tol=2e-3))
    add("X4_urban_delta_recip", view_factor_matrix, urb,
        MatrixParams(samples=2, rays=32, seed=5, bvh="builtin", device="cpu", max_iters=20, min_iters=3, tol=2e-3, tol_mode="delta",
                     reciprocity=True))
    add("X5_tilted_matrix", view_factor_matrix, tilted,
        MatrixParams(samples=32, rays=64, seed=8, bvh="builtin", device="cpu", max_iters=30, min_iters=5, tol=1e-3, reciprocity=False))
    add("X5_tilted_sky", view_factor_to_tregenza_sky, tilted,
        SkyParams(samples=32, rays=64, seed=8, bvh="off", device="cpu", max_iters=12, min_iters=4, tol=1e-3, discrete=False))
    # BASELINE config #4 "at high ray count to convergence tolerance": the ex04 cube with 16 384 rays per iteration, run
    # until the replicate standard error of every entry is <= 1e-4 (hundreds of iterations)
    add("X3_cube_converged_1e-4", view_factor_matrix, cube,
        MatrixParams(samples=64, rays=256, seed=42, bvh="auto", device="cpu", flip_faces=True, reciprocity=False, max_iters=4000,
                     min_iters=10, tol=1e-4))
    (HERE / "solves_extra.json").write_text(json.dumps(out, indent=1, sort_keys=True))


def workflow():
    """view_factor_outside_workflow (api.py) and view_factor_matrix_and_sky (main.py:1209) on the canyon."""
    from raystrack import view_factor_outside_workflow
    from raystrack.main import view_factor_matrix_and_sky
    canyon = load_meshes_json(str(REF / "examples" / "street_canyon.json"))
    out = {}

    def clean(d):
        return {k: {kk: float(vv) for kk, vv in row.items()} for k, row in d.items()}

    cases = {
        "W1_shared_recip_merged": (dict(samples=8, rays=64, seed=4, bvh="off", device="cpu", max_iters=40, min_iters=6, tol=5e-4, reciprocity=True),
                                   dict(samples=8, rays=64, seed=4, bvh="off", device="cpu", max_iters=30, min_iters=5, tol=5e-4, discrete=False)),
        "W2_shared_discrete_rowsum": (dict(samples=8, rays=64, seed=4, bvh="builtin", device="cpu", max_iters=25, min_iters=6, tol=1e-3,
                                           reciprocity=True, enforce_reciprocity_rowsum=True),
                                      dict(samples=8, rays=64, seed=4, bvh="builtin", device="cpu", max_iters=12, min_iters=12, tol=0.0, discrete=True)),
        "W3_separate_norecip": (dict(samples=8, rays=64, seed=4, bvh="off", device="cpu", max_iters=20, min_iters=6, tol=1e-3, reciprocity=False),
                                dict(samples=6, rays=32, seed=9, bvh="off", device="cpu", max_iters=15, min_iters=5, tol=1e-3, discrete=False)),
    }
    old = ref_main._log
    ref_main._log = lambda m: None
    try:
        for name, (mp, sp) in cases.items():
            vf, sky, rest = view_factor_outside_workflow(canyon, matrix_params=MatrixParams(**mp), sky_params=SkyParams(**sp))
            out[name] = {"matrix_params": MatrixParams(**mp).as_dict(), "sky_params": SkyParams(**sp).as_dict(),
                         "vf_scene": clean(vf), "sky_vf": clean(sky), "rest_vf": clean(rest)}
            print(name, "done")
        mp, sp = cases["W1_shared_recip_merged"]
        vf, sky = view_factor_matrix_and_sky(canyon[:1], matrix_params=MatrixParams(**mp), sky_params=SkyParams(**sp))
        out["W4_single_mesh_shared"] = {"matrix_params": MatrixParams(**mp).as_dict(), "sky_params": SkyParams(**sp).as_dict(),
                                        "vf_scene": clean(vf), "sky_vf": clean(sky)}
    finally:
        ref_main._log = old
    (HERE / "workflow.json").write_text(json.dumps(out, indent=1, sort_keys=True))


def shipped():
    out = {}
    for rel in ("examples/vf_matrix.json", "examples/inside_vf_matrix.json",
                "validation/results/06_canyon_view3d_raystrack_raw.json",
                "validation/view3d_reference/canyon_view3d_base.json"):
        out[rel] = json.loads((REF / rel).read_text())
    txt = {}
    for p in sorted((REF / "validation" / "results").glob("0[1-5]_*.txt")):
        body = p.read_text()
        val = float(re.search(r"raystrack:\s+([0-9.]+)", body).group(1))
        ana = float(re.search(r"analytical:\s+([0-9.]+)", body).group(1))
        its = {m.group(1): int(m.group(2)) for m in re.finditer(r"^\s{4}(\S+): (\d+)$", body, re.M)}
        txt[p.stem] = {"raystrack": val, "analytical": ana, "iterations": its}
    out["validation_results_txt"] = txt
    (HERE / "shipped.json").write_text(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    what = sys.argv[1:] or ["stage", "solves", "extra", "shipped", "workflow", "large"]
    if "stage" in what:
        stage_vectors()
    if "solves" in what:
        solves()
    if "extra" in what:
        solves_extra()
    if "shipped" in what:
        shipped()
    if "workflow" in what:
        workflow()
    if "large" in what:
        large_rays()
