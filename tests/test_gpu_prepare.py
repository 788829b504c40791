"""Device-side preparation (csrc/rsk_prepare.cu) against the host preparation, bit for bit.

The host arrays of ``prepare_scene`` / ``prepare_emitters`` are themselves pinned to the reference's
``utils/prepared.py`` by tests/test_oracle_golden.py and tests/test_host_logic.py; here the records the GPU computes
from raw vertices and faces must equal them in every bit, and whole solves must not depend on which side prepared."""
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


def _soup(seed: int, n_tri: int, scale: float = 5.0):
    """Random triangle soup with shared vertices, a few degenerate triangles and negative (from-the-end) indices."""
    rng = np.random.default_rng(seed)
    V = (rng.standard_normal((max(3, n_tri // 2 + 3), 3)) * scale).astype(np.float64)
    F = rng.integers(0, V.shape[0], (n_tri, 3)).astype(np.int64)
    if n_tri > 4:
        F[1] = F[1, 0]                      # zero-area triangle
        F[3, 2] = -1                        # NumPy-style negative index
    return V, F


def _scenes():
    from raystrack_b200 import synthetic
    soup = [(f"soup{i}", *_soup(i, n)) for i, n in enumerate((1, 7, 8, 9, 127, 128, 129, 1000, 4099))]
    axis = np.array([[0, 0, 0], [1e-9, 0, 0], [0, 1e-9, 0]], np.float32)          # normal ~ 0: frame fallback branch
    xnorm = ("xnormal", np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [0, 1, 1]], np.float32), np.array([[0, 1, 2], [1, 3, 2]], np.int32))
    tiny = ("tiny", axis, np.array([[0, 1, 2]], np.int32))
    big = synthetic.quad_grid("big", (0.0, 0.0, 0.0), (37.0, 3.0, 1.0), (-2.0, 41.0, 5.0), 150)        # 45 000 triangles, tilted
    return {
        "canyon": synthetic.street_canyon(),
        "cube": synthetic.unit_cube_enclosure(),
        "tilted": synthetic.tilted_pair(),
        "urban": synthetic.urban_block(3, 4, 8, 0),
        "soup": soup + [xnorm, tiny],
        "big": [big, synthetic.quad_grid("lid", (0.0, 0.0, 30.0), (40.0, 0.0, 0.0), (0.0, 40.0, 0.0), 3)],
    }


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", ["canyon", "cube", "tilted", "urban", "soup", "big"])
@pytest.mark.parametrize("flip", [False, True])
def test_device_records_equal_host_arrays(rb, name, flip):
    from raystrack_b200 import _native
    from raystrack_b200.prepared import PreparedSolver, flatten_meshes, summaries_from_device
    meshes = _scenes()[name]
    samples, rays = 7, 5
    host = PreparedSolver(meshes)
    ems = host.get_emitters(samples=samples, rays=rays, flip_faces=flip)
    hs = host.get_scene(use_bvh=False)
    ctx = _native.Context.for_device(0)
    geo = _native.DeviceGeometry(ctx, *flatten_meshes(meshes))
    d_em, summary = _native.DeviceEmitters.from_geometry(geo, samples, rays, flip)
    rec, cdf = d_em.download_records(geo.n_tri)
    cat = lambda f: np.concatenate([getattr(e, f) for e in ems], axis=0)
    assert np.array_equal(_bits(rec[:, 0:3]), _bits(cat("tri_a")))
    assert np.array_equal(_bits(rec[:, 3]), _bits(cat("tri_origin_eps")))
    assert np.array_equal(_bits(rec[:, 4:7]), _bits(cat("tri_e1")))
    assert np.array_equal(_bits(rec[:, 8:11]), _bits(cat("tri_e2")))
    assert np.array_equal(_bits(rec[:, [7, 11, 15]]), _bits(cat("tri_n")))
    assert np.array_equal(_bits(rec[:, 12:15]), _bits(cat("tri_u")))
    assert np.array_equal(_bits(rec[:, 16:19]), _bits(cat("tri_v")))
    assert np.array_equal(_bits(cdf), _bits(cat("cdf")))
    assert np.array_equal(d_em.g, np.asarray([e.g for e in ems], np.int32))
    # per-mesh by-products: area, grid side and the complete plane record
    got = summaries_from_device(summary, d_em.g, meshes, samples=samples, rays=rays, flip_faces=flip)
    for s, e in zip(got, ems):
        assert s.total_area == e.total_area and s.g == e.g and s.plane_tol == e.plane_tol
        assert s.plane_is_planar == e.plane_is_planar
        assert np.array_equal(_bits(s.plane_origin), _bits(e.plane_origin)) and np.array_equal(_bits(s.plane_normal), _bits(e.plane_normal))
    # scene records (input order without a BVH)
    d_sc = _native.DeviceScene.from_geometry(geo, False)
    tri, nrm = d_sc.download_triangles()
    assert np.array_equal(_bits(tri[:, 0:3]), _bits(hs.v0)) and np.array_equal(_bits(tri[:, 4:7]), _bits(hs.e1))
    assert np.array_equal(_bits(tri[:, 8:11]), _bits(hs.e2)) and np.array_equal(_bits(nrm[:, 0:3]), _bits(hs.normals))
    assert np.array_equal(tri[:, 3].view(np.int32), hs.sid) and np.array_equal(nrm[:, 3].view(np.int32), hs.sid)
    for h in (d_sc, d_em, geo):
        h.close()


def test_device_planarity_verdicts_are_never_wrong(rb):
    """The device decides planarity only when its float64 statistics are clear of the thresholds; whenever it
    decides, the verdict is the host's.  Axis-aligned walls are decided on the device ("planar"), tilted grids far
    from the origin too ("not planar": float32 vertex rounding bends them); tilted grids near the origin deviate from
    their plane by about the tolerance itself -- the reference's own verdict is rounding noise there -- and go to the
    host."""
    from raystrack_b200 import _native, synthetic
    from raystrack_b200.prepared import device_plane_verdict, flatten_meshes, prepare_emitters
    rng = np.random.default_rng(11)
    meshes = [synthetic.quad_grid("wall", (2.0, 0.0, 0.0), (0.0, 3.0, 0.0), (0.0, 0.0, 4.0), 6)]
    for i in range(40):
        scale = 3.0 if i % 2 else 600.0
        p0 = rng.uniform(-scale, scale, 3)
        du, dv = rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3)
        name, V, F = synthetic.quad_grid(f"tilt{i}", tuple(p0), tuple(du), tuple(dv), 10)
        if i % 3 == 0:
            V = V + (rng.standard_normal(V.shape) * 10.0 ** rng.uniform(-8, -4)).astype(V.dtype)      # slightly bent
        meshes.append((name, V, F))
    ctx = _native.Context.for_device(0)
    geo = _native.DeviceGeometry(ctx, *flatten_meshes(meshes))
    d_em, summary = _native.DeviceEmitters.from_geometry(geo, 4, 4, False)
    host = prepare_emitters(meshes, samples=4, rays=4, flip_faces=False)
    verdicts = [device_plane_verdict(row, e.plane_tol) for row, e in zip(summary, host)]
    assert verdicts[0] is True and host[0].plane_is_planar
    decided = [(v, e.plane_is_planar) for v, e in zip(verdicts, host) if v is not None]
    assert all(v == h for v, h in decided)
    assert any(v is None for v in verdicts) and any(v is False for v in verdicts)      # both routes are exercised
    from raystrack_b200.prepared import summaries_from_device
    final = summaries_from_device(summary, d_em.g, meshes, samples=4, rays=4, flip_faces=False)
    assert [s.plane_is_planar for s in final] == [e.plane_is_planar for e in host]
    d_em.close()
    geo.close()


def test_solves_do_not_depend_on_where_preparation_ran(rb, monkeypatch):
    from raystrack_b200 import synthetic
    for meshes, kw in ((synthetic.street_canyon(), dict(samples=16, rays=32, bvh="off")),
                       (synthetic.unit_cube_enclosure(), dict(samples=16, rays=16, bvh="builtin", flip_faces=True, reciprocity=False)),
                       (synthetic.urban_block(3, 4, 8, 0), dict(samples=2, rays=16, bvh="builtin"))):
        p = rb.MatrixParams(seed=4, max_iters=12, min_iters=4, tol=1e-3, **kw)
        sp = rb.SkyParams(samples=kw["samples"], rays=kw["rays"], seed=4, bvh=kw["bvh"], max_iters=6, min_iters=3, discrete=True)
        monkeypatch.setenv("RSK_HOST_PREPARE", "1")
        host_m, host_s = rb.view_factor_matrix(meshes, p), rb.view_factor_to_tregenza_sky(meshes, sp)
        monkeypatch.delenv("RSK_HOST_PREPARE")
        assert rb.view_factor_matrix(meshes, p) == host_m
        assert rb.view_factor_to_tregenza_sky(meshes, sp) == host_s


def test_prepared_solver_uses_device_preparation_by_default(rb):
    """No host emitter arrays are built on the default path; asking for them afterwards still works and agrees."""
    from raystrack_b200 import synthetic
    meshes = synthetic.urban_block(2, 4, 4, 1)
    ps = rb.PreparedSolver(meshes)
    p = rb.MatrixParams(samples=4, rays=8, seed=1, bvh="builtin", max_iters=5, min_iters=5, tol=0.0)
    first = rb.view_factor_matrix(meshes, p, prepared=ps)
    assert not ps._emitter_cache and not ps._scene_cache and ps._geometry_cache
    ems = ps.get_emitters(samples=4, rays=8, flip_faces=False)
    sums = ps.get_emitter_summaries(samples=4, rays=8, flip_faces=False)
    assert [e.g for e in ems] == [s.g for s in sums] and [e.total_area for e in ems] == [s.total_area for s in sums]
    ps.clear_device_cache()
    assert rb.view_factor_matrix(meshes, p, prepared=ps) == first          # now through the host arrays


def test_out_of_range_face_index_raises_index_error(rb):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    bad = [("a", V, np.array([[0, 1, 3]], np.int32)), ("b", V + 1.0, np.array([[0, 1, 2]], np.int32))]
    with pytest.raises(IndexError):
        rb.view_factor_matrix(bad, rb.MatrixParams(samples=4, rays=4, max_iters=2, min_iters=1))


@pytest.mark.parametrize("name", ["canyon", "cube", "tilted", "urban", "soup"])
def test_device_surface_masks_equal_host_masks(rb, name):
    """rsk_surface_masks against main._surface_masks (itself pinned to the oracle's per-emitter loop on the host)."""
    from raystrack_b200 import _native, main as M
    from raystrack_b200.prepared import PreparedSolver
    meshes = _scenes()[name]
    ps = PreparedSolver(meshes)
    ems = ps.get_emitters(samples=4, rays=4, flip_faces=False)
    centers, extents = ps.get_mesh_bounds()
    host = M._surface_masks(ems, centers, extents)
    ctx = _native.Context.for_device(0)
    dev = M._cached_masks(PreparedSolver(meshes), ems, centers, extents, False, ctx)
    assert dev.dtype == np.uint8 and np.array_equal(dev, host)
