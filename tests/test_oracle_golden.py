"""Pin the CPU oracle (oracle/) to the reference: every stage against vectors produced by the reference
itself (tests/golden/make_golden.py) and whole solves against the reference's results and shipped files."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from raystrack_b200 import synthetic
from scenes import URBAN_RAY_CASES, scene_for


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_halton_tables_bit_exact(stage):
    dims = O.halton_dims(70000)
    for r in range(5):
        assert np.array_equal(dims[r][:4096], stage["halton_dims_head"][r])
        assert np.array_equal(dims[r][-512:], stage["halton_dims_tail"][r])
    for g in (4, 16, 26):
        u, v = O.halton_grid(g)
        assert np.array_equal(u, stage[f"grid_u_{g}"]) and np.array_equal(v, stage[f"grid_v_{g}"])


@pytest.mark.parametrize("tag,flip", [("tilted", False), ("tiltedflip", True), ("canyon", False)])
def test_emitter_preparation_bit_exact(stage, tag, flip):
    meshes = synthetic.street_canyon() if tag == "canyon" else synthetic.tilted_pair()
    for i, em in enumerate(O.prepare_emitters(meshes, 16, 8, flip)):
        for f, mine in (("tri_a", em.tri_a), ("tri_e1", em.tri_e1), ("tri_e2", em.tri_e2), ("tri_u", em.tri_u),
                        ("tri_v", em.tri_v), ("tri_n", em.tri_n), ("tri_origin_eps", em.tri_eps), ("cdf", em.cdf),
                        ("plane_origin", em.plane_origin), ("plane_normal", em.plane_normal)):
            assert np.array_equal(mine, stage[f"em_{tag}_{i}_{f}"]), (tag, i, f)
        sc = stage[f"em_{tag}_{i}_scalars"]
        assert (sc[0], sc[1], bool(sc[2]), int(sc[3])) == (em.total_area, em.plane_tol, em.planar, em.g)


def test_rays_bit_exact(stage):
    em = O.prepare_emitters(synthetic.tilted_pair(), 64, 16, False)[0]
    cp = stage["rays_tilted_cp"]
    o, d = O.build_rays(em, cp[:2], cp[2:])
    assert np.array_equal(o, stage["rays_tilted_orig"]) and np.array_equal(d, stage["rays_tilted_dir"])
    em = O.prepare_emitters(synthetic.street_canyon(), 16, 128, False)[10]
    cp = stage["rays_road_cp"]
    o, d = O.build_rays(em, cp[:2], cp[2:])
    assert _sha(o) + _sha(d) == bytes(stage["rays_road_sha"]).hex()


def test_bvh_arrays_bit_exact(stage):
    sb = O.prepare_scene(synthetic.urban_block(3, 4, 8, 0), True)
    for f, mine in zip(("bb_min", "bb_max", "left", "right", "start", "count"), sb.bvh):
        assert np.array_equal(mine, stage["bvh_" + f]), f
    assert np.array_equal(sb.sid, stage["bvh_sid"]) and _sha(sb.v0) == bytes(stage["bvh_v0_sha"]).hex()


@pytest.mark.parametrize("idx,recip", URBAN_RAY_CASES)
def test_per_ray_hits_bit_exact(stage, idx, recip):
    meshes = synthetic.urban_block(3, 4, 8, 0)
    S = O.OracleSolver(meshes)
    em = S.emitters(4, 16, False)[idx]
    c, e = S.bounds()
    k = f"urb_{idx}_{int(recip)}"
    cp = stage[k + "_cp"]
    o, d = O.build_rays(em, cp[:2], cp[2:])
    act = O.surface_mask(idx, em, c, e)
    assert np.array_equal(act, stage[k + "_active"])
    es, ms = (idx, idx + 1) if recip else (idx, 0)
    for scene, suffix in ((S.scene(True), "bvh"), (S.scene(False), "brute")):
        hs, fr = O.trace_firsthit(scene, o, d, act, es, ms)
        assert np.array_equal(hs, stage[f"{k}_hit_{suffix}"]) and np.array_equal(fr, stage[f"{k}_front_{suffix}"])
        assert np.array_equal(O.trace_hitmask(scene, o, d, act, idx, 0), stage[f"{k}_mask_{suffix}"])
    cs, cf, ca = O.trace_combined(S.scene(True), o, d, act, idx, ms)
    assert np.array_equal(cs, stage[k + "_comb_hit"]) and np.array_equal(cf, stage[k + "_comb_front"])
    assert np.array_equal(ca, stage[k + "_comb_any"])
    hm = stage[k + "_mask_bvh"]
    assert np.array_equal(O.bin_tregenza(d, hm), stage[k + "_tregenza"])
    assert O.count_upward_misses(d, hm) == int(stage[k + "_upward"][0])


def test_tregenza_patch_ids_bit_exact(stage):
    ids = np.array([O.tregenza_patch_id(*r) for r in stage["treg_dirs"]])
    assert np.array_equal(ids, stage["treg_ids"])


# whole solves: identical keys/iteration counts; values differ by at most a few rays (float near-ties: the
# reference is compiled with fastmath, the oracle evaluates in source order)
@pytest.mark.parametrize("case", ["C1_readme_squares", "C2_canyon_ex01", "C2b_canyon_delta_norecip",
                                  "C3_canyon_sky_discrete", "C3b_canyon_sky_merged", "C4_cube_ex04",
                                  "U3_urban_matrix_recip", "U3_urban_sky",
                                  "X2_canyon_sky_delta", "X3_cube_flip_bvh", "X4_urban_delta_recip", "X5_tilted_matrix", "X5_tilted_sky",
                                  "X3_cube_converged_1e-4"])
def test_whole_solve_matches_reference(solves, case):
    g = solves[case]
    S = O.OracleSolver(scene_for(case))
    pe = {}
    p = g["params"]
    res = S.view_factor_to_tregenza_sky(per_emitter=pe, **p) if "discrete" in p else S.view_factor_matrix(per_emitter=pe, **p)
    assert {k: v["iters"] for k, v in pe.items()} == g["iters"]
    for name, row in g["result"].items():
        # a key exists as soon as ONE ray hits that receiver side, so a near-tie ray may add or drop a key whose value
        # is a single ray's weight; every value (absent = 0) must agree to 1e-5
        for key in set(res[name]) | set(row):
            assert abs(res[name].get(key, 0.0) - row.get(key, 0.0)) <= 1e-5, (name, key)


def test_shipped_result_files(solves, shipped):
    """The reference's own shipped outputs are reproduced by the reference run recorded in solves.json
    (bit-identical), which the oracle in turn reproduces (test above)."""
    pairs = (("C2_canyon_ex01", "examples/vf_matrix.json"), ("C4_cube_ex04", "examples/inside_vf_matrix.json"),
             ("V06_canyon_view3d", "validation/results/06_canyon_view3d_raystrack_raw.json"))
    for case, rel in pairs:
        ship = shipped[rel]
        res = solves[case]["result"]
        for name, row in ship.items():
            assert {k for k, v in res[name].items() if v != 0.0} == set(row)
            for key, val in row.items():
                assert res[name][key] == val


LARGE_EMITTERS = (0, 4, 122, 245)


def test_oracle_matches_reference_on_a_127k_triangle_tree():
    """tests/golden/large_rays.npz: the reference's trace_cpu_bvh_firsthit / _hitmask through its own median-split BVH over
    127 488 triangles (7x7 urban block).  The oracle builds the same tree and must reproduce every ray."""
    from pathlib import Path
    from raystrack_b200 import synthetic
    z = np.load(Path(__file__).resolve().parent / "golden" / "large_rays.npz")
    meshes = synthetic.urban_block(7, 16, 32, 0)
    S = O.OracleSolver(meshes)
    scene = S.scene(True)
    assert scene.v0.shape[0] == int(z["n_tri"][0]) and scene.bvh[2].shape[0] == int(z["n_nodes"][0])
    for idx in LARGE_EMITTERS:
        em = O.prepare_emitters([meshes[idx]], 4, 16, False)[0]
        cp = z[f"e{idx}_cp"]
        n = z[f"e{idx}_hit"].shape[0]
        o, d = O.build_rays(em, cp[:2], cp[2:], count=n)
        act = z[f"e{idx}_active"]
        hit, front = O.trace_firsthit(scene, o, d, act, idx, 0)
        assert np.array_equal(hit, z[f"e{idx}_hit"].astype(np.int32)) and np.array_equal(front, z[f"e{idx}_front"])
        assert np.array_equal(O.trace_hitmask(scene, o, d, act, idx, 0), z[f"e{idx}_mask"])


ZERO_AREA_MESH = ("line", np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]], np.float32), np.array([[0, 1, 2], [1, 2, 3]], np.int32))


def test_zero_area_emitter_uses_all_zero_tables():
    """utils/prepared.py:278-287: a zero-area mesh keeps g = 4, a cdf of ones and all-zero jitter / Halton tables, so only
    the Cranley-Patterson offsets move its rays (reference rays in tests/golden/large_rays.npz)."""
    from pathlib import Path
    from raystrack_b200 import prepared as P
    z = np.load(Path(__file__).resolve().parent / "golden" / "large_rays.npz")
    em = O.prepare_emitters([ZERO_AREA_MESH], 4, 8, False)[0]
    cp = z["zero_area_cp"]
    o, d = O.build_rays(em, cp[:2], cp[2:])
    assert np.array_equal(o, z["zero_area_orig"], equal_nan=True) and np.array_equal(d, z["zero_area_dir"], equal_nan=True)
    host = P.prepare_emitters([ZERO_AREA_MESH], samples=4, rays=8, flip_faces=False)[0]
    assert host.g == 4 and not host.u_grid.any() and not host.halton_tri.any() and not host.halton_r2.any()
    assert np.array_equal(host.cdf, np.ones(2, np.float32))
