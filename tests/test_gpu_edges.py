"""Edge cases of the CUDA path on a B200: tiny / degenerate scenes, the global-atomic tally fallback, forced
brute force on a BVH-sized scene, rerun determinism and PreparedSolver reuse -- each checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import raystrack_b200
    from raystrack_b200 import _native
    if _native.device_count() <= 0:
        pytest.fail("no CUDA device visible")
    import raystrack_b200.main as M
    M._log = lambda m: None
    return raystrack_b200


def _worst(a, b):
    w = 0.0
    for name in set(a) | set(b):
        for key in set(a.get(name, {})) | set(b.get(name, {})):
            w = max(w, abs(a.get(name, {}).get(key, 0.0) - b.get(name, {}).get(key, 0.0)))
    return w


def test_tiny_scene_forced_bvh(rb):
    """Two single-triangle meshes with bvh='builtin': the builder's <=3-triangle root path."""
    from oracle import oracle as O
    t1 = ("t1", np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.int32))
    t2 = ("t2", np.array([[0, 0, 1], [0, 1, 1], [1, 0, 1]], np.float32), np.array([[0, 1, 2]], np.int32))
    p = dict(samples=64, rays=32, seed=3, bvh="builtin", max_iters=8, min_iters=8, tol=0.0, reciprocity=False)
    got = rb.view_factor_matrix([t1, t2], rb.MatrixParams(**p))
    want = O.OracleSolver([t1, t2]).view_factor_matrix(**p)
    assert got["t1"] and _worst(got, want) <= 1e-5


def test_degenerate_triangles_and_nonplanar_emitter(rb):
    """A mesh with a zero-area triangle and a folded (non-planar) emitter: no NaNs, same tallies as the oracle."""
    from oracle import oracle as O
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0.4], [0, 1, 0], [0.5, 0.5, 0.5]], np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3], [4, 4, 4]], np.int32)            # last one is degenerate
    lid = ("lid", np.array([[-1, -1, 2], [2, -1, 2], [2, 2, 2], [-1, 2, 2]], np.float32), np.array([[0, 2, 1], [0, 3, 2]], np.int32))
    meshes = [("fold", V, F), lid]
    p = dict(samples=32, rays=16, seed=5, bvh="off", max_iters=6, min_iters=6, tol=0.0, reciprocity=False)
    got = rb.view_factor_matrix(meshes, rb.MatrixParams(**p))
    want = O.OracleSolver(meshes).view_factor_matrix(**p)
    assert all(np.isfinite(v) for row in got.values() for v in row.values())
    assert _worst(got, want) <= 1e-5


def test_brute_force_equals_bvh_on_large_scene(rb):
    """Closest hits do not depend on the acceleration structure: bvh='off' and bvh='builtin' give the same tallies
    on a 10k-triangle urban block (up to float near-ties), and both match the oracle."""
    from oracle import oracle as O
    from raystrack_b200 import synthetic
    meshes = synthetic.urban_block(4, 8, 8, 1)
    p = dict(samples=1, rays=8, seed=2, max_iters=3, min_iters=3, tol=0.0, reciprocity=False)
    a = rb.view_factor_matrix(meshes, rb.MatrixParams(bvh="off", **p))
    b = rb.view_factor_matrix(meshes, rb.MatrixParams(bvh="builtin", **p))
    assert _worst(a, b) <= 2e-5
    want = O.OracleSolver(meshes).view_factor_matrix(bvh="builtin", **p)
    assert _worst(b, want) <= 2e-5


def test_many_surfaces_global_tally_path(rb):
    """30 000 single-quad meshes: the receiver histogram (2 x 30 000 x 4 B = 240 KB) no longer fits in shared memory, so
    the kernel tallies with warp-aggregated global atomics.  Four emitters are solved through the C ABI and compared
    with the oracle's per-ray results."""
    from oracle import oracle as O
    from raystrack_b200 import _native
    from raystrack_b200.main import _rotation_table, _surface_masks
    from raystrack_b200.prepared import PreparedSolver
    rng = np.random.default_rng(0)
    n = 30000
    base = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    F2 = np.array([[0, 2, 1], [0, 3, 2]], np.int32)
    gx, gy = np.meshgrid(np.arange(150), np.arange(100), indexing="ij")
    meshes = []
    for k in range(n // 2):
        off = np.array([1.5 * gx.flat[k], 1.5 * gy.flat[k], 0.0], np.float32)
        meshes.append((f"f{k}", base + off, F))
        meshes.append((f"c{k}", base * np.array([1.2, 1.2, 1], np.float32) + off + np.array([0, 0, 1 + rng.uniform(0, 1)], np.float32), F2))
    ctx = _native.Context.for_device(0)
    ps = PreparedSolver(meshes)
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx).native
    em = ps.get_device_emitters(samples=64, rays=32, flip_faces=False, ctx=ctx).native
    ems = ps.get_emitters(samples=64, rays=32, flip_faces=False)
    ids = np.array([0, 777, 15000, 29998], np.int32)
    centers, extents = ps.get_mesh_bounds()
    active = _surface_masks([ems[i] for i in ids], centers, extents)          # rows for the chosen emitters only
    active = np.ones((len(ids), n), np.uint8)
    for r, i in enumerate(ids):
        active[r, i] = 0
    table = _rotation_table(4, n, 3)
    solve = _native.Solve(ctx, sc, em, ids, active, table, ids.copy(), max_iters=3, min_iters=3, interval=1, tol_mode="stderr",
                          tol=0.0, emit_sid=ids, min_sid=np.zeros(len(ids), np.int32))
    assert solve.step(3) == 0
    tallies, iters, totals = solve.read_block()
    solve.close()
    assert list(iters) == [3, 3, 3, 3]
    S = O.OracleSolver(meshes)
    oscene = S.scene(True)
    oems = S.emitters(64, 32, False)
    for r, i in enumerate(ids):
        want = np.zeros(2 * n, np.int64)
        for it in range(3):
            cpg, cpd = O.rotation(4, int(i), it)
            o, d = O.build_rays(oems[i], cpg, cpd)
            hs, fr = O.trace_firsthit(oscene, o, d, active[r], int(i), 0)
            hit = hs >= 0
            np.add.at(want, 2 * hs[hit] + (1 - fr[hit].astype(np.int64)), 1)
        assert np.abs(tallies[r] - want).sum() <= 2, (i, np.abs(tallies[r] - want).sum())
        assert tallies[r].sum() > 0


def test_rerun_is_deterministic_and_prepared_reuse(rb):
    """Integer tallies: two runs give bit-identical dictionaries; a PreparedSolver reused with a new seed gives the
    same result as a fresh solve with that seed (examples/ex05_prepared_seed_compare.py)."""
    from raystrack_b200 import synthetic
    meshes = synthetic.street_canyon()
    ps = rb.PreparedSolver(meshes)
    p1 = rb.MatrixParams(samples=8, rays=32, seed=1, max_iters=30, min_iters=5, tol=1e-3)
    a = rb.view_factor_matrix(meshes, p1, prepared=ps)
    b = rb.view_factor_matrix(meshes, p1, prepared=ps)
    assert a == b
    p2 = rb.MatrixParams(samples=8, rays=32, seed=7, max_iters=30, min_iters=5, tol=1e-3)
    c = rb.view_factor_matrix(meshes, p2, prepared=ps)
    d = rb.view_factor_matrix(meshes, p2)
    assert c == d and c != a
    sender = rb.view_factor(meshes[0], meshes[1:], p1)
    assert set(sender) == {"east_side_0"} and sender["east_side_0"] == a["east_side_0"]


def test_convergence_interval_and_max_iters_edge(rb):
    """convergence_interval > 1 delays the stop to the next checkpoint (reference GPU schedule, main.py:392-416);
    max_iters=0 returns empty rows."""
    from raystrack_b200 import synthetic
    import raystrack_b200.main as M
    meshes = synthetic.street_canyon()
    logs = []
    M._log = logs.append
    try:
        rb.view_factor_matrix(meshes, rb.MatrixParams(samples=8, rays=32, seed=1, max_iters=60, min_iters=5, tol=1e-3, convergence_interval=1))
        it1 = [int(l.split("]")[1].split()[0]) for l in logs]
        logs.clear()
        rb.view_factor_matrix(meshes, rb.MatrixParams(samples=8, rays=32, seed=1, max_iters=60, min_iters=5, tol=1e-3, convergence_interval=7))
        it7 = [int(l.split("]")[1].split()[0]) for l in logs]
    finally:
        M._log = lambda m: None
    for x, y in zip(it1, it7):
        if x == 0:
            assert y == 0
        else:
            assert y >= x and (y - 5) % 7 == 0 or y == 60
    empty = rb.view_factor_matrix(meshes, rb.MatrixParams(samples=8, rays=32, max_iters=0))
    assert all(row == {} for row in empty.values())


def test_chunked_solve_equals_single_solve(rb, monkeypatch):
    """RSK_SOLVE_MEMORY_MB bounds the per-solve device state; emitters are then solved in chunks.  Emitters are
    independent, so the result must not change by a single bit."""
    from raystrack_b200 import synthetic
    meshes = synthetic.urban_block(3, 4, 8, 0)
    p = rb.MatrixParams(samples=2, rays=16, seed=3, bvh="builtin", max_iters=20, min_iters=4, tol=2e-3, reciprocity=True)
    sp = rb.SkyParams(samples=2, rays=16, seed=3, bvh="builtin", max_iters=9, min_iters=4, tol=1e-3, discrete=True)
    import raystrack_b200.main as M
    whole = rb.view_factor_matrix(meshes, p)
    whole_sky = rb.view_factor_to_tregenza_sky(meshes, sp)
    whole_both = M.view_factor_matrix_and_sky(meshes, matrix_params=p, sky_params=sp)
    monkeypatch.setenv("RSK_SOLVE_MEMORY_MB", "0.02")            # ~5 emitters per chunk
    assert rb.view_factor_matrix(meshes, p) == whole
    assert rb.view_factor_to_tregenza_sky(meshes, sp) == whole_sky
    assert M.view_factor_matrix_and_sky(meshes, matrix_params=p, sky_params=sp) == whole_both      # dual solve in chunks


@pytest.mark.parametrize("offset", [(0.0, 0.0, 0.0), (512345.0, 4112233.0, 250.0)])
def test_georeferenced_coordinates_per_ray(rb, offset):
    """City models often carry projected (UTM-like) coordinates of 10^5..10^6 m: float32 vertices then have ~0.03-0.5 m
    of resolution and the quantised BVH boxes must stay conservative.  Rays (bit-equal) and per-ray closest hits are
    compared with the oracle on the same translated scene, through the BVH and by brute force."""
    from oracle import oracle as O
    from raystrack_b200 import _native, synthetic
    from raystrack_b200.prepared import PreparedSolver
    off = np.asarray(offset, np.float64)
    meshes = [(name, np.asarray(V, np.float64) + off, F) for name, V, F in synthetic.urban_block(3, 4, 8, 0)]
    ctx = _native.Context.for_device(0)
    S = O.OracleSolver(meshes)
    oem = S.emitters(4, 16, False)
    centers, extents = S.bounds()
    for use_bvh in (True, False):
        ps = PreparedSolver(meshes)
        sc = ps.get_device_scene(use_bvh=use_bvh, ctx=ctx).native
        em = ps.get_device_emitters(samples=4, rays=16, flip_faces=False, ctx=ctx).native
        agree = total = 0
        for e in (0, 7, 22, 45):
            cpg, cpd = O.rotation(3, e, 1)
            act = O.surface_mask(e, oem[e], centers, extents)
            o, d, hit, front = _native.trace_rays(ctx, sc, em, e, act, e, 0, np.concatenate([cpg, cpd]), mode=0)
            ro, rd = O.build_rays(oem[e], cpg, cpd)
            assert np.array_equal(o, ro) and np.array_equal(d, rd)
            rh, rf = O.trace_firsthit(S.scene(use_bvh), ro, rd, act, e, 0)
            agree += int(np.sum((hit == rh) & (front == rf)))
            total += hit.shape[0]
        assert agree / total >= 0.9999, (use_bvh, agree, total)
        ps.clear_device_cache()


@pytest.mark.parametrize("seed", range(8))
def test_random_triangle_soups_per_ray(rb, seed):
    """Unstructured geometry: meshes of random, mutually intersecting triangles with very different sizes (incl.
    slivers and a zero-area triangle).  Device preparation, the LBVH/wide-BVH builder and both traversal modes
    must still agree with the oracle ray by ray (closest hit and any-hit), through the BVH and by brute force."""
    from oracle import oracle as O
    from raystrack_b200 import _native
    from raystrack_b200.prepared import PreparedSolver
    rng = np.random.default_rng(1000 + seed)
    meshes = []
    for m in range(7):
        nt = int(rng.integers(1, 90))
        centre = rng.uniform(-4, 4, 3)
        size = 10.0 ** rng.uniform(-1.5, 0.8)
        V = centre + rng.standard_normal((nt * 3, 3)) * size
        if m == 2 and nt > 1:
            V[3:6] = V[3]                                                  # zero-area triangle
        if m == 4 and nt > 2:
            V[7] = V[6] + 1e-4 * (V[8] - V[6])                             # sliver
        meshes.append((f"soup{m}", V, np.arange(nt * 3, dtype=np.int64).reshape(nt, 3)))
    ctx = _native.Context.for_device(0)
    S = O.OracleSolver(meshes)
    oem = S.emitters(8, 16, False)
    centers, extents = S.bounds()
    agree = total = 0
    for use_bvh in (True, False):
        ps = PreparedSolver(meshes)
        sc = ps.get_device_scene(use_bvh=use_bvh, ctx=ctx).native
        em = ps.get_device_emitters(samples=8, rays=16, flip_faces=False, ctx=ctx).native
        for e in (0, 3, 6):
            cpg, cpd = O.rotation(seed, e, 2)
            act = O.surface_mask(e, oem[e], centers, extents)
            ro, rd = O.build_rays(oem[e], cpg, cpd)
            o, d, hit, front = _native.trace_rays(ctx, sc, em, e, act, e, 0, np.concatenate([cpg, cpd]), mode=0)
            assert np.array_equal(o, ro) and np.array_equal(d, rd)
            rh, rf = O.trace_firsthit(S.scene(use_bvh), ro, rd, act, e, 0)
            agree += int(np.sum((hit == rh) & (front == rf)))
            _, _, occl, _ = _native.trace_rays(ctx, sc, em, e, act, e, 0, np.concatenate([cpg, cpd]), mode=1, want_rays=False)
            rmask = O.trace_hitmask(S.scene(use_bvh), ro, rd, act, e, 0)
            agree += int(np.sum(occl.astype(bool) == np.asarray(rmask).astype(bool)))
            total += 2 * hit.shape[0]
        ps.clear_device_cache()
    assert total > 0 and agree / total >= 0.9999, (agree, total)


def test_device_csr_rows_equal_dense_block(rb):
    """rsk_solve_csr / rsk_tally_block_csr: the compressed result rows built on the device equal the dense tally block
    divided on the host (same int64 -> float64 conversions, one IEEE division), bit for bit."""
    from raystrack_b200 import _native, main as M, synthetic
    from raystrack_b200.prepared import PreparedSolver
    meshes = synthetic.urban_block(3, 4, 8, 0)
    ctx = _native.Context.for_device(0)
    ps = PreparedSolver(meshes)
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx)
    em = ps.get_device_emitters(samples=2, rays=16, flip_faces=False, ctx=ctx)
    ems = ps.get_emitter_summaries(samples=2, rays=16, flip_faces=False, ctx=ctx)
    n = len(meshes)
    active = M._surface_masks(ems, *ps.get_mesh_bounds())
    ids = np.arange(n, dtype=np.int32)
    solve = _native.Solve(ctx, sc.native, em.native, ids, active, M._rotation_table(5, n, 6), ids.copy(), max_iters=6, min_iters=6,
                          interval=1, tol_mode="stderr", tol=0.0, emit_sid=ids, min_sid=np.zeros(n, np.int32))
    try:
        solve.step(6)
        dense, iters, totals = solve.read_block()
        view, iters2, totals2 = solve.read_block_view()
        assert np.array_equal(view, dense) and np.array_equal(iters, iters2) and np.array_equal(totals, totals2)
        want = M._csr_from_dense(dense, totals)
        got = solve.read_csr()
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and np.array_equal(a, b)
        assert got[0][-1] > 100
        # the same through a tally block: rows scattered by emitter id, jobs with keep == 0 left out
        blk = _native.TallyBlock(ctx, n, 2 * n)
        keep = np.ones(n, np.uint8)
        keep[3] = 0
        blk.add_solve(solve, keep)
        blk.allreduce()                                              # no communicator: a no-op
        dense_k = dense.copy()
        dense_k[3] = 0
        assert np.array_equal(blk.download(copy=True), dense_k)
        for a, b in zip(blk.read_csr(totals), M._csr_from_dense(dense_k, totals)):
            assert np.array_equal(a, b)
        blk.close()
    finally:
        solve.close()


def test_terrain_with_small_objects_per_ray_and_solve(rb):
    """Strongly varying triangle sizes (0.85 m terrain, centimetre spheres, 50 m slabs; non-planar emitters): per-ray closest
    hits vs the oracle through the GPU-built tree, and a short whole solve."""
    from oracle import oracle as O
    from raystrack_b200 import _native, synthetic
    from raystrack_b200.prepared import PreparedSolver
    meshes = synthetic.terrain_with_objects(2, 60, 40, 20, 2, 1)          # 66 meshes, 33 764 triangles
    ctx = _native.Context.for_device(0)
    ps = PreparedSolver(meshes)
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx).native
    em = ps.get_device_emitters(samples=1, rays=16, flip_faces=False, ctx=ctx).native
    S = O.OracleSolver(meshes)
    oem = S.emitters(1, 16, False)
    centers, extents = S.bounds()
    scene = S.scene(True)
    for idx in (0, 3, 4, 30, 44, 50, 64):                                  # terrain tiles, boxes, spheres, a slab
        cpg, cpd = O.rotation(7, idx, 2)
        act = O.surface_mask(idx, oem[idx], centers, extents)
        n = min(oem[idx].n_rays_once, 16384)
        o, d, hit, front = _native.trace_rays(ctx, sc, em, idx, act, idx, 0, np.concatenate([cpg, cpd]), mode=0, n_rays=n)
        ro, rd = O.build_rays(oem[idx], cpg, cpd, count=n)
        assert np.array_equal(o, ro) and np.array_equal(d, rd)
        rh, rf = O.trace_firsthit(scene, ro, rd, act, idx, 0)
        assert float(np.mean((hit == rh) & (front == rf))) >= 0.9999, idx
    params = dict(samples=1, rays=16, seed=4, bvh="builtin", max_iters=4, min_iters=4, tol=0.0, reciprocity=False)
    got = rb.view_factor_matrix(meshes, rb.MatrixParams(**params))
    assert _worst(got, S.view_factor_matrix(**params)) <= 1e-4


def test_zero_area_and_empty_emitters(rb):
    """A zero-area emitter shoots the reference's rays (all-zero QMC tables, prepared.py:278-287) on both preparation
    routes; a mesh without triangles is never traced and does not disturb its neighbours."""
    from pathlib import Path
    from raystrack_b200 import _native, synthetic
    from raystrack_b200.prepared import PreparedSolver
    z = np.load(Path(__file__).resolve().parent / "golden" / "large_rays.npz")
    line = ("line", np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]], np.float32), np.array([[0, 1, 2], [1, 2, 3]], np.int32))
    sq = synthetic.parallel_unit_squares()
    meshes = [line, sq[0], sq[1]]
    ctx = _native.Context.for_device(0)
    for host in (False, True):
        ps = PreparedSolver(meshes)
        if host:
            ps._host_prepare_forced = lambda: True
        sc = ps.get_device_scene(use_bvh=False, ctx=ctx).native
        em = ps.get_device_emitters(samples=4, rays=8, flip_faces=False, ctx=ctx).native
        o, d, _, _ = _native.trace_rays(ctx, sc, em, 0, np.ones(3, np.uint8), 0, 0, z["zero_area_cp"], mode=0)
        assert np.array_equal(o, z["zero_area_orig"], equal_nan=True) and np.array_equal(d, z["zero_area_dir"], equal_nan=True), host
    empty = ("nothing", np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32))
    p = rb.MatrixParams(samples=16, rays=32, seed=3, bvh="off", max_iters=6, min_iters=6, tol=0.0, reciprocity=False)
    with_empty = rb.view_factor_matrix([sq[0], empty, sq[1]], p)
    assert with_empty["nothing"] == {} and all("nothing" not in k for row in with_empty.values() for k in row)
    assert with_empty["A"] == rb.view_factor_matrix(sq, p)["A"]
    sky = rb.view_factor_to_tregenza_sky([sq[0], empty, sq[1]], rb.SkyParams(samples=16, rays=32, seed=3, bvh="off", max_iters=6, min_iters=6))
    assert sky["nothing"] == {"Sky": 0.0}


def test_pipelined_and_sequential_stepping_agree(rb, tmp_path):
    """Iterations are pipelined over two streams (a job may be traced once more before its stop decision is known; the
    statistics kernel drops those tallies).  RSK_PIPELINE=0 (read once per process, hence a child process) steps strictly
    one iteration after the other: results and iteration counts must be identical, for all three solve kinds."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    code = (
        "import json, sys\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "import raystrack_b200 as rb\n"
        "from raystrack_b200 import main as M, synthetic\n"
        "logs = []\n"
        "M._log = logs.append\n"
        "meshes = synthetic.urban_block(3, 4, 8, 0)\n"
        "p = rb.MatrixParams(samples=2, rays=16, seed=3, bvh='builtin', max_iters=80, min_iters=3, tol=4e-4, reciprocity=True)\n"
        "sp = rb.SkyParams(samples=2, rays=16, seed=3, bvh='builtin', max_iters=60, min_iters=3, tol=3e-4, discrete=True)\n"
        "out = {'m': rb.view_factor_matrix(meshes, p), 's': rb.view_factor_to_tregenza_sky(meshes, sp),\n"
        "       'd': M.view_factor_matrix_and_sky(meshes, matrix_params=p, sky_params=sp), 'logs': [l.split('->')[0] for l in logs]}\n"
        "print('RESULT' + json.dumps(out, sort_keys=True, default=float))\n")
    outs = {}
    for mode in ("1", "0"):
        import os
        env = dict(os.environ, RSK_PIPELINE=mode)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[mode] = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1]
    assert outs["1"] == outs["0"]
    logs = json.loads(outs["1"][6:])["logs"]
    iters = {int(ln.split("]")[1].split("iter")[0]) for ln in logs if "iter" in ln and "traced" not in ln}
    assert len(iters) >= 4 and max(iters) >= 8, iters          # emitters stop at different iterations: the stop path is exercised


@pytest.mark.parametrize("reciprocity", [False, True])
def test_two_part_solve_with_overlapped_rows_equals_one_solve(rb, monkeypatch, reciprocity):
    """Large solves are cut in two (main._overlap_split) so that the rows of the first part are assembled on a worker
    thread while the second part is traced: same dictionaries, same key order (rows and reciprocity fill-ins are
    written in emitter order either way), same progress lines apart from the time shares, and a converging solve stops
    every emitter at the same iteration."""
    import raystrack_b200.main as M
    from raystrack_b200 import synthetic
    meshes = synthetic.urban_block(3, face_grid=4, ground_grid=8)              # 46 meshes; the ground holds the most rays
    prm = rb.MatrixParams(samples=8, rays=16, seed=5, bvh="builtin", reciprocity=reciprocity, max_iters=12, min_iters=3, tol=2e-3)
    lines = {}

    def run(on: bool):
        monkeypatch.setenv("RSK_OVERLAP_ASSEMBLY", "1" if on else "0")
        got = []
        monkeypatch.setattr(M, "_log", got.append)
        out = rb.view_factor_matrix(meshes, prm)
        lines[on] = [ln.split(" -> ")[0] for ln in got]
        return out

    one = run(False)
    monkeypatch.setattr(M, "OVERLAP_MIN_EMITTERS", 8)
    monkeypatch.setattr(M, "OVERLAP_MIN_RAYS", 0.0)
    monkeypatch.setenv("RSK_OVERLAP_ASSEMBLY", "1")
    todo = list(range(len(meshes)))
    n_once = [1000] * (len(meshes) - 1) + [20000]
    cut = M._overlap_split(todo, n_once, 1)
    assert 0 < cut < len(meshes) and sum(n_once[cut:]) >= 0.15 * sum(n_once) > sum(n_once[cut + 1:])
    two = run(True)
    assert "assemble" in M.LAST_TIMING
    assert one == two
    for name in one:
        assert list(one[name]) == list(two[name])
    assert lines[False] == lines[True] and len(lines[True]) == len(meshes)
    monkeypatch.setattr(M, "_log", lambda m: None)
