"""Outside-workflow post-processing (raystrack_b200/api.py, reciprocity.py) against the reference's results, on CPU:
the raw matrix/sky inputs come from the oracle (pinned to the reference), the product code does the rest."""
import numpy as np
import pytest

from oracle import oracle as O
from raystrack_b200 import MatrixParams, SkyParams, api, synthetic
from raystrack_b200 import reciprocity as R


def _oracle_raw(mp, sp):
    S = O.OracleSolver(synthetic.street_canyon())
    m = {k: v for k, v in mp.items() if k != "enforce_reciprocity_rowsum"}
    return S.view_factor_matrix(**m), S.view_factor_to_tregenza_sky(**sp)


def _max_diff(a, b):
    worst = 0.0
    for name in set(a) | set(b):
        ra, rb = a.get(name, {}), b.get(name, {})
        for key in set(ra) | set(rb):
            worst = max(worst, abs(ra.get(key, 0.0) - rb.get(key, 0.0)))
    return worst


@pytest.mark.parametrize("case", ["W1_shared_recip_merged", "W3_separate_norecip"])
def test_workflow_postprocessing_matches_reference(workflow_golden, monkeypatch, case):
    g = workflow_golden[case]
    vf_raw, sky_raw = _oracle_raw(g["matrix_params"], g["sky_params"])
    monkeypatch.setattr(api, "view_factor_matrix_and_sky", lambda meshes, **kw: (vf_raw, sky_raw))
    monkeypatch.setattr(api, "view_factor_matrix", lambda meshes, **kw: vf_raw)
    monkeypatch.setattr(api, "view_factor_to_tregenza_sky", lambda meshes, **kw: sky_raw)
    vf, sky, rest = api.view_factor_outside_workflow(synthetic.street_canyon(), matrix_params=MatrixParams(**g["matrix_params"]),
                                                     sky_params=SkyParams(**g["sky_params"]))
    assert _max_diff(vf, g["vf_scene"]) <= 1e-5
    assert _max_diff(sky, g["sky_vf"]) <= 1e-5
    assert _max_diff(rest, g["rest_vf"]) <= 2e-5
    for name in vf:
        assert set(vf[name]) == set(g["vf_scene"][name])
        assert abs(sum(vf[name].values()) + sum(sky[name].values()) + rest[name]["Rest"] - 1.0) <= 1e-5 or rest[name]["Rest"] == 0.0


def test_reciprocity_only_enforces_symmetry():
    meshes = synthetic.street_canyon()
    S = O.OracleSolver(meshes)
    vf = S.view_factor_matrix(samples=8, rays=32, seed=2, bvh="off", max_iters=8, min_iters=8, tol=0.0, reciprocity=False)
    before = {k: dict(v) for k, v in vf.items()}
    R.enforce_reciprocity_only(vf, meshes)
    names = [m[0] for m in meshes]
    A = R._mesh_areas(meshes)
    F = R._totals_matrix(vf, names)
    assert np.allclose(A[:, None] * F, (A[:, None] * F).T, atol=1e-12)
    F0 = R._totals_matrix(before, names)
    assert np.allclose(A[:, None] * F, 0.5 * (A[:, None] * F0 + (A[:, None] * F0).T), atol=1e-12)


def test_shareable_rule_and_errors():
    from raystrack_b200 import main as M
    assert M.outside_workflow_shareable(MatrixParams(), SkyParams())
    assert not M.outside_workflow_shareable(MatrixParams(flip_faces=True), SkyParams())
    assert not M.outside_workflow_shareable(MatrixParams(samples=8), SkyParams())
    with pytest.raises(TypeError):
        api.view_factor_outside_workflow([], matrix_params=SkyParams(), sky_params=SkyParams())
    with pytest.raises(TypeError):
        M.view_factor_matrix_and_sky([], matrix_params=MatrixParams(), sky_params=MatrixParams())
    with pytest.raises(ValueError):
        M.view_factor_matrix_and_sky([], matrix_params=MatrixParams(samples=8), sky_params=SkyParams())
