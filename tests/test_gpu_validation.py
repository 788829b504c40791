"""The reference's analytic validation suite (validation/validate_01..05, common_validation.run_raystrack settings:
bvh="builtin", device="cpu", tol=1e-4 stderr, min_iters=40, max_iters=500, seed=11, reciprocity=False) on the CUDA
path: the closed form within 1e-4 (the reference's pass criterion), and the reference's own printed value and
per-emitter iteration counts (validation/results/0N_*.txt) reproduced."""
import re

import pytest

from validation_shapes import analytic_cases

pytestmark = pytest.mark.gpu
CASES = analytic_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_validation_case(shipped, case):
    import raystrack_b200 as rb
    import raystrack_b200.main as M
    key, meshes, samples, rays, emitter, receiver, exact = case
    logs = []
    old = M._log
    M._log = logs.append
    try:
        vf = rb.view_factor_matrix(meshes, rb.MatrixParams(samples=samples, rays=rays, seed=11, bvh="builtin", device="cpu", cuda_async=False,
                                                           gpu_raygen=False, max_iters=500, min_iters=40, tol=1e-4, tol_mode="stderr",
                                                           convergence_interval=1, reciprocity=False))
    finally:
        M._log = old
    value = float(vf[emitter].get(f"{receiver}_front", 0.0))
    ref = shipped["validation_results_txt"][key]
    assert abs(exact - ref["analytical"]) < 1e-9
    assert abs(value - exact) <= 1e-4                                   # validation/common_validation.py:231-232
    assert abs(value - ref["raystrack"]) <= 2e-6                        # the reference's own result (printed to 1e-10)
    pat = re.compile(r"\[\s*(?P<name>[^\]]+?)\s*\]\s+(?P<iters>\d+)\s+iter")
    iters = {m.group("name"): int(m.group("iters")) for m in map(pat.search, logs) if m}
    assert iters == ref["iterations"]
