"""Host-side logic of the product (no GPU): preparation, masks, rotations, shard planning, JSON I/O, params --
each against the oracle / the reference-generated golden vectors."""
import json

import numpy as np
import pytest

from oracle import oracle as O
from raystrack_b200 import MatrixParams, SkyParams, synthetic
from raystrack_b200 import io as rio
from raystrack_b200 import main as M
from raystrack_b200 import prepared as P


@pytest.mark.parametrize("tag,flip", [("tilted", False), ("tiltedflip", True), ("canyon", False)])
def test_emitter_preparation_matches_reference(stage, tag, flip):
    meshes = synthetic.street_canyon() if tag == "canyon" else synthetic.tilted_pair()
    for i, em in enumerate(P.prepare_emitters(meshes, samples=16, rays=8, flip_faces=flip)):
        for f in ("tri_a", "tri_e1", "tri_e2", "tri_u", "tri_v", "tri_n", "tri_origin_eps", "cdf", "plane_origin", "plane_normal"):
            assert np.array_equal(getattr(em, f), stage[f"em_{tag}_{i}_{f}"]), (tag, i, f)
        sc = stage[f"em_{tag}_{i}_scalars"]
        assert (sc[0], sc[1], bool(sc[2]), int(sc[3])) == (em.total_area, em.plane_tol, em.plane_is_planar, em.g)


def test_vectorised_frames_equal_per_triangle_frames():
    rng = np.random.default_rng(3)
    n = rng.normal(size=(5000, 3)).astype(np.float32)
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    n[:5] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, 0, 0]]
    u, v = P._triangle_frames(n)
    ou, ov = O.triangle_frames(n)
    assert np.array_equal(u, ou) and np.array_equal(v, ov)


def test_scene_matches_oracle_in_mesh_order():
    meshes = synthetic.urban_block(3, 4, 8, 0)
    mine = P.prepare_scene(meshes, use_bvh=True)
    ref = O.prepare_scene(meshes, use_bvh=False)
    for f in ("v0", "e1", "e2", "normals", "sid"):
        assert np.array_equal(getattr(mine, f), getattr(ref, f))
    assert mine.use_bvh and mine.bb_min is None


def test_host_halton_properties_match_reference(stage):
    em = P.prepare_emitters(synthetic.street_canyon(), samples=16, rays=8, flip_faces=False)[0]
    assert em.g == 26
    assert np.array_equal(em.u_grid, stage["grid_u_26"]) and np.array_equal(em.v_grid, stage["grid_v_26"])
    n = em.n_cells * 8
    for r, name in enumerate(("halton_tri", "halton_u", "halton_v", "halton_r1", "halton_r2")):
        assert np.array_equal(getattr(em, name)[:4096], stage["halton_dims_head"][r][:min(n, 4096)])


def test_surface_masks_and_receivers_match_oracle():
    for meshes in (synthetic.urban_block(3, 4, 8, 0), synthetic.street_canyon(), synthetic.tilted_pair()):
        ps = P.PreparedSolver(meshes)
        ems = ps.get_emitters(samples=4, rays=8, flip_faces=False)
        c, e = ps.get_mesh_bounds()
        oc, oe = O.mesh_bounds(meshes)
        assert np.array_equal(c, oc) and np.array_equal(e, oe)
        masks = M._surface_masks(ems, c, e)
        oems = O.prepare_emitters(meshes, 4, 8, False)
        for i in range(len(meshes)):
            assert np.array_equal(masks[i], O.surface_mask(i, oems[i], oc, oe))


def test_rotation_table_matches_reference_rng():
    table = M._rotation_table(7, 5, 9)
    assert table.shape == (14, 7)
    for i in range(5):
        for it in range(9):
            cpg, cpd = O.rotation(7, i, it)
            assert np.array_equal(table[i + it, :2], cpg) and np.array_equal(table[i + it, 2:], cpd)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plan_shards_covers_every_ray_once(world):
    rng = np.random.default_rng(world)
    n_once = [int(x) for x in rng.integers(2048, 200000, 60)] + [45158400, 9000000]
    todo = [i for i in range(len(n_once)) if i % 7 != 3]
    plans = M.plan_shards(todo, n_once, world)
    assert len(plans) == world
    cover = {}
    for r, plan in enumerate(plans):
        shared_part = [j for j in plan if j[3]]
        assert plan[: len(shared_part)] == shared_part            # shared jobs first ...
        assert [j[0] for j in shared_part] == [j[0] for j in plans[0] if j[3]]   # ... same order on every rank
        assert len({j[0] for j in plan}) == len(plan)              # an emitter appears once per rank
        for i, b, e, shared in plan:
            assert 0 <= b <= e <= n_once[i]
            assert b % M.TILE_RAYS == 0
            cover.setdefault(i, []).append((b, e))
    assert sorted(cover) == sorted(todo)
    for i, parts in cover.items():
        parts.sort()
        assert parts[0][0] == 0 and parts[-1][1] == n_once[i]
        assert all(parts[k][1] == parts[k + 1][0] for k in range(len(parts) - 1))
    loads = [sum(e - b for _, b, e, _ in plan) for plan in plans]
    assert max(loads) <= 1.05 * (sum(loads) / world) + 250000


@pytest.mark.parametrize("world", [1, 2, 5])
def test_plan_shards_without_splitting_assigns_whole_emitters(world):
    """The shared-ray (dual) solve shards whole emitters only: every emitter on exactly one rank, full ray range."""
    n_once = [4096 * k for k in (3, 1, 40, 7, 7, 2, 900, 5)]
    plans = M.plan_shards(list(range(len(n_once))), n_once, world, allow_split=False)
    jobs = [j for plan in plans for j in plan]
    assert sorted(j[0] for j in jobs) == list(range(len(n_once)))
    assert all(b == 0 and e == n_once[i] and not shared for i, b, e, shared in jobs)


def test_select_bvh_and_errors():
    assert M._select_bvh("auto", 511) is False and M._select_bvh("auto", 512) is True
    assert M._select_bvh("builtin", 1) is True and M._select_bvh("off", 10 ** 6) is False and M._select_bvh(None, 600) is True
    with pytest.raises(ValueError):
        M._select_bvh("fast", 10)
    with pytest.raises(TypeError):
        M._ensure_prepared([], object())


def test_params_roundtrip_and_defaults():
    p = MatrixParams()
    assert (p.samples, p.rays, p.seed, p.bvh, p.device, p.max_iters, p.tol, p.tol_mode, p.min_iters, p.convergence_interval,
            p.reciprocity, p.enforce_reciprocity_rowsum, p.flip_faces) == (16, 128, 1, "auto", "auto", 100, 1e-4, "stderr", 5, 1, True, False, False)
    assert MatrixParams.from_dict(p.as_dict()) == p
    s = SkyParams(discrete=True)
    assert SkyParams.from_dict(s.as_dict()) == s and "reciprocity" not in s.as_dict()


def test_save_vf_matrix_json(tmp_path):
    vf = {"b": {"a_front": 0.25, "a_back": 0.5, "c_front": 0.0}, "a": {"b_front": np.float64(0.125)}}
    path = rio.save_vf_matrix_json(vf, str(tmp_path / "sub" / "out"))
    assert path.endswith("out.json")
    text = open(path).read()
    assert text == json.dumps({"a": {"b_front": 0.125}, "b": {"a_back": 0.5, "a_front": 0.25}}, indent=2, sort_keys=True)
    assert rio.load_vf_matrix_json(path) == {"a": {"b_front": 0.125}, "b": {"a_back": 0.5, "a_front": 0.25}}
    path2 = rio.save_vf_matrix_json([vf, {"b": {"d_back": 1.0}}], str(tmp_path / "m.json"), strip_dir=True)
    assert json.load(open(path2)) == {"a": {"b": 0.125}, "b": {"a": 0.75, "d": 1.0}}
    with pytest.raises(TypeError):
        rio.save_vf_matrix_json({"a": {"b": "x"}}, str(tmp_path / "bad.json"))
    with pytest.raises(TypeError):
        rio.save_vf_matrix_json(3, str(tmp_path / "bad.json"))


def test_meshes_json_roundtrip(tmp_path):
    meshes = synthetic.street_canyon()
    path = rio.save_meshes_json(meshes, str(tmp_path / "geo"))
    back = rio.load_meshes_json(path)
    assert [m[0] for m in back] == [m[0] for m in meshes]
    assert all(np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) for a, b in zip(meshes, back))


def test_reciprocity_write_back_split():
    from raystrack_b200 import reciprocity as R
    res = {"a": {"b_front": 0.2, "b_back": 0.2}, "b": {"a_back": 0.1}}
    Fp = np.array([[0.0, 0.8], [0.3, 0.0]])
    R._write_back(res, ["a", "b"], Fp)
    assert res["a"] == {"b_front": 0.4, "b_back": 0.4} and res["b"] == {"a_back": pytest.approx(0.3)}
    assert np.array_equal(R._totals_matrix(res, ["a", "b"]), np.array([[0.0, 0.8], [0.3, 0.0]]))


# --------------------------------------------------------------------------- host side of the device-side preparation

def _emulated_summary(meshes, flip):
    """What rsk_prepare_meshes_kernel returns (float64 statistics against the first triangle), computed with NumPy from
    the host-prepared arrays; lets the host half of the device path be tested without a GPU."""
    from raystrack_b200 import _native
    from raystrack_b200.prepared import prepare_emitters
    ems = prepare_emitters(meshes, samples=4, rays=4, flip_faces=flip)
    S = np.zeros(len(meshes), _native.MESH_SUMMARY_DTYPE)
    for i, e in enumerate(ems):
        if e.tri_a.shape[0] == 0:
            continue
        n = e.tri_n.astype(np.float64)
        n0, org = n[0], e.tri_a[0]
        worst = mag = 0.0
        for p in (e.tri_a, e.tri_a + e.tri_e1, e.tri_a + e.tri_e2):
            d = (p - org).astype(np.float64) * n0
            worst = max(worst, float(np.abs(d.sum(1)).max()))
            mag = max(mag, float(np.abs(d).sum(1).max()))
        S[i] = (e.total_area, org, e.tri_n[0], e.tri_origin_eps.max(), 0, (n @ n0).min(), worst, mag)
    return ems, S


@pytest.mark.parametrize("flip", [False, True])
def test_summaries_from_device_reproduce_the_plane_records(flip):
    from raystrack_b200 import synthetic
    from raystrack_b200.prepared import device_plane_verdict, summaries_from_device
    rng = np.random.default_rng(5)
    meshes = synthetic.urban_block(2, 4, 4, 0) + synthetic.tilted_pair() + synthetic.unit_cube_enclosure()
    meshes.append(("empty", np.zeros((0, 3)), np.zeros((0, 3), np.int64)))
    meshes.append(("needle", np.array([[0, 0, 0], [1e-9, 0, 0], [0, 1e-9, 0]], np.float32), np.array([[0, 1, 2]])))
    for i in range(24):                                     # tilted grids near and far from the origin, some slightly bent
        scale = 3.0 if i % 2 else 600.0
        name, V, F = synthetic.quad_grid(f"t{i}", tuple(rng.uniform(-scale, scale, 3)), tuple(rng.uniform(-2, 2, 3)),
                                         tuple(rng.uniform(-2, 2, 3)), 6)
        if i % 3 == 0:
            V = V + (rng.standard_normal(V.shape) * 10.0 ** rng.uniform(-8, -4)).astype(V.dtype)
        meshes.append((name, V, F))
    ems, S = _emulated_summary(meshes, flip)
    got = summaries_from_device(S, np.asarray([e.g for e in ems]), meshes, samples=4, rays=4, flip_faces=flip)
    assert len(got) == len(ems)
    for s, e in zip(got, ems):
        assert s.plane_is_planar == e.plane_is_planar and s.plane_tol == e.plane_tol
        assert s.total_area == e.total_area and s.g == e.g and s.n_cells == e.n_cells
        assert np.array_equal(s.plane_origin.view(np.uint32), e.plane_origin.view(np.uint32))
        assert np.array_equal(s.plane_normal.view(np.uint32), e.plane_normal.view(np.uint32))
    verdicts = [device_plane_verdict(row, e.plane_tol) for row, e in zip(S, ems) if e.tri_a.shape[0]]
    decided = [(v, e.plane_is_planar) for v, e in zip(verdicts, [e for e in ems if e.tri_a.shape[0]]) if v is not None]
    assert decided and all(v == h for v, h in decided)       # the statistics never decide wrongly ...
    assert any(v is None for v in verdicts)                  # ... and the fallback is exercised


def test_flatten_meshes_layout_and_index_range():
    from raystrack_b200.prepared import flatten_meshes
    a = ("a", np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float64), np.array([[0, 1, 2], [1, 3, -1]], np.int64))
    b = ("b", np.zeros((0, 3)), np.zeros((0, 3), np.int64))
    c = ("c", [[0, 0, 1], [1, 0, 1], [0, 1, 1]], [[0, 1, 2]])                         # plain lists, as load_meshes_json gives
    verts, vo, faces, to = flatten_meshes([a, b, c])
    assert verts.dtype == np.float32 and faces.dtype == np.int32 and verts.flags.c_contiguous and faces.flags.c_contiguous
    assert vo.tolist() == [0, 4, 4, 7] and to.tolist() == [0, 2, 2, 3]
    assert faces.tolist() == [[0, 1, 2], [1, 3, -1], [0, 1, 2]] and np.array_equal(verts[4:], np.asarray(c[1], np.float32))
    assert [x.shape for x in flatten_meshes([])] == [(0, 3), (1,), (0, 3), (1,)]
    with pytest.raises(IndexError):
        flatten_meshes([("big", np.zeros((3, 3)), np.array([[0, 1, 2 ** 40]]))])


def test_prepared_solver_mesh_bounds_match_per_mesh_loop():
    from raystrack_b200 import synthetic
    from raystrack_b200.prepared import PreparedSolver
    for meshes in (synthetic.urban_block(2, 3, 4, 1), [("e", np.zeros((0, 3)), np.zeros((0, 3), np.int64))] + synthetic.tilted_pair()):
        centers, extents = PreparedSolver(meshes).get_mesh_bounds()
        for i, (_, V, _) in enumerate(meshes):
            v = np.asarray(V, np.float32)
            if v.size == 0:
                assert not centers[i].any() and not extents[i].any()
                continue
            lo, hi = v.min(0), v.max(0)
            assert np.array_equal(centers[i], (0.5 * (lo + hi)).astype(np.float32)) and np.array_equal(extents[i], (0.5 * (hi - lo)).astype(np.float32))


def test_plan_chunks_respects_budget_and_keeps_shared_jobs_together(monkeypatch):
    """_plan_chunks: every job in exactly one chunk, order preserved, at most budget/(48*bins) jobs per chunk -- except
    that all ray-split (shared) jobs stay in the first chunk, because every rank must step them in lockstep."""
    plan = [(0, 0, 8192, True), (3, 0, 4096, True), (1, 0, 100, False), (2, 0, 100, False), (4, 0, 100, False), (5, 0, 100, False), (6, 0, 100, False)]
    n_hist = 100                                                    # 4800 bytes of solve state per job
    monkeypatch.setenv("RSK_SOLVE_MEMORY_MB", str(3 * 4800 / (1 << 20)))      # three jobs per chunk
    chunks = M._plan_chunks(plan, n_hist)
    assert [j for c in chunks for j in c] == plan
    assert all(len(c) <= 3 for c in chunks) and len(chunks) == 3
    monkeypatch.setenv("RSK_SOLVE_MEMORY_MB", str(1 * 4800 / (1 << 20)))      # one job per chunk, but two shared jobs
    chunks = M._plan_chunks(plan, n_hist)
    assert chunks[0] == plan[:2] and all(len(c) == 1 for c in chunks[1:]) and [j for c in chunks for j in c] == plan
    monkeypatch.delenv("RSK_SOLVE_MEMORY_MB")
    assert M._plan_chunks(plan, n_hist) == [plan] and M._plan_chunks([], n_hist) == [[]]


def test_csr_from_dense_matches_the_dense_row_loop():
    """`_csr_from_dense` (the host form of rsk_solve_csr): non-zero bins per row, columns ascending, F = hits / rays."""
    rng = np.random.default_rng(3)
    t = np.zeros((7, 12), np.int64)
    m = rng.random(t.shape) < 0.3
    t[m] = rng.integers(1, 10**9, int(m.sum()))
    t[4] = 0                                                       # an emitter that never ran: total 0, no bins
    totals = np.array([5, 7, 11, 13, 0, 17, 10**12], np.int64)
    row_ptr, cols, vals = M._csr_from_dense(t, totals)
    assert row_ptr.dtype == np.int64 and cols.dtype == np.int32 and vals.dtype == np.float64
    assert row_ptr[0] == 0 and row_ptr[-1] == int(m.sum()) - int(m[4].sum())
    for i in range(t.shape[0]):
        nz = np.flatnonzero(t[i])
        assert cols[row_ptr[i]:row_ptr[i + 1]].tolist() == nz.tolist()
        assert vals[row_ptr[i]:row_ptr[i + 1]].tolist() == [int(t[i, j]) / float(totals[i]) for j in nz]
    empty = M._csr_from_dense(np.zeros((0, 4), np.int64), np.zeros(0, np.int64))
    assert empty[0].tolist() == [0] and empty[1].size == 0


def test_overlap_split_rule(monkeypatch):
    """main._overlap_split: large solves are cut where the suffix still holds 15 % of the rays; small ones never."""
    from raystrack_b200 import main as M
    n = 2001
    n_once = [97_000] * (n - 1) + [45_000_000]                    # C5: 2000 facades and roofs + the ground
    todo = list(range(n))
    assert M._overlap_split(todo, n_once, 40) == n - 1            # the ground alone is 19 % of the rays
    assert M._overlap_split(todo, n_once, 1) == 0                 # 0.24 G rays certain: not worth a second solve
    assert M._overlap_split(todo[:100], n_once, 1000) == 0        # few emitters
    even = [1_000_000] * 1000
    k = M._overlap_split(list(range(1000)), even, 5)
    assert k == 850                                               # uniform scene: the last 15 % of the emitters
    head_heavy = [50_000_000] + [10_000] * 999
    assert M._overlap_split(list(range(1000)), head_heavy, 40) == 0   # the suffix would hold most rows: no cut
    monkeypatch.setenv("RSK_OVERLAP_ASSEMBLY", "0")
    assert M._overlap_split(todo, n_once, 40) == 0


def test_rotation_rows_equal_numpy_generators():
    """main._rotation_rows: the vectorised SeedSequence + PCG64 restatement (_pcg.py) gives the rows of
    ``default_rng(seed + s)`` bit for bit; seeds outside its range use the generators themselves."""
    import pytest
    from raystrack_b200 import main as M, _pcg
    rng = np.random.default_rng(2024)
    seeds = [0, 1, 11, 31, 2**31 - 3, 2**32 - 40] + rng.integers(0, 2**32 - 64, 60).tolist()
    total = 0
    for seed in seeds:
        rows = 40 if seed else 2600
        got = _pcg.rotation_rows(int(seed), rows)
        want = M._rotation_rows_generators(int(seed), rows)
        assert got.dtype == np.float32 and np.array_equal(got, want), seed
        total += rows
    assert total >= 5000
    M._rotation_rows.cache_clear()
    big = M._rotation_rows(2**32 - 3, 8)                          # crosses 2**32: two-word entropy -> generator loop
    assert np.array_equal(big, M._rotation_rows_generators(2**32 - 3, 8)) and not big.flags.writeable
    assert np.array_equal(M._rotation_table(7, 11, 100), M._rotation_rows_generators(7, 111))
    with pytest.raises(ValueError):
        M._rotation_rows(-5, 4)
    M._rotation_rows.cache_clear()
