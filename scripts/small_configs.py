"""Wall time of the reference's own small configurations (C1-C4, validation case 06, the ex03
workflow) through the public API on one B200, next to the reference timings quoted there.  Results must also equal the
goldens (tests do that; here only a max |dF| is printed for orientation).

    python scripts/small_configs.py > profiles/small_configs_r1.json
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import raystrack_b200 as rb                              # noqa: E402
from raystrack_b200 import main as M, synthetic          # noqa: E402
from scenes import scene_for                             # noqa: E402

REFERENCE_CPU_S = {      # BASELINE.md section 2 (8-core build container, Numba CPU path)
    "C1_readme_squares": "14.0 s incl. JIT (5.77 M rays)", "C2_canyon_ex01": "2.7 s warm (20.2 M rays)",
    "C3_canyon_sky_discrete": "5.5 s incl. JIT (19.9 M rays)", "C4_cube_ex04": "0.06 s",
    "V06_canyon_view3d": "15.8 s incl. JIT (73 M rays)", "ex03_workflow": "189 s (1.62e9 rays)",
}


def worst(a, b):
    return max((abs(a.get(n, {}).get(k, 0.0) - b.get(n, {}).get(k, 0.0)) for n in set(a) | set(b)
                for k in set(a.get(n, {})) | set(b.get(n, {}))), default=0.0)


if __name__ == "__main__":
    logs = []
    M._log = logs.append
    gold = json.loads((ROOT / "tests" / "golden" / "solves.json").read_text())
    out = {}
    for case in ("C1_readme_squares", "C2_canyon_ex01", "C3_canyon_sky_discrete", "C4_cube_ex04", "V06_canyon_view3d"):
        g = gold[case]
        meshes = scene_for(case)
        sky = "discrete" in g["params"]
        call = (lambda: rb.view_factor_to_tregenza_sky(meshes, rb.SkyParams(**g["params"]))) if sky else \
               (lambda: rb.view_factor_matrix(meshes, rb.MatrixParams(**g["params"])))
        call()                                                      # context, QMC tables
        times = []
        for _ in range(5):
            logs.clear()
            t = time.perf_counter()
            res = call()
            times.append(time.perf_counter() - t)
        rays = sum(int(l.split(" iter, ")[1].split(" rays")[0].replace(",", "")) for l in logs if " iter, " in l)
        out[case] = {"ms": round(1e3 * sorted(times)[2], 2), "rays": rays, "max_abs_diff_vs_reference": worst(res, g["result"]),
                     "reference_cpu": REFERENCE_CPU_S[case]}
    # ex03 workflow parameters (examples/ex03: shared matrix + sky solve, tol 1e-5, up to 500 iterations)
    meshes = synthetic.street_canyon()
    mp = rb.MatrixParams(samples=32, rays=256, seed=7, bvh="builtin", max_iters=500, min_iters=5, tol=1e-5, tol_mode="stderr",
                         enforce_reciprocity_rowsum=False, reciprocity=True)
    sp = rb.SkyParams(samples=32, rays=256, seed=7, bvh="builtin", max_iters=500, min_iters=5, tol=1e-5, tol_mode="stderr", discrete=False)
    rb.view_factor_outside_workflow(meshes, matrix_params=mp, sky_params=sp)
    times = []
    for _ in range(3):
        logs.clear()
        t = time.perf_counter()
        rb.view_factor_outside_workflow(meshes, matrix_params=mp, sky_params=sp)
        times.append(time.perf_counter() - t)
    rays = sum(int(l.split(" iter, ")[1].split(" rays")[0].replace(",", "")) for l in logs if "traced" in l)
    out["ex03_workflow"] = {"ms": round(1e3 * sorted(times)[1], 2), "rays": rays, "reference_cpu": REFERENCE_CPU_S["ex03_workflow"],
                            "note": "street canyon, samples=32 rays=256 seed=7 tol=1e-5 max_iters=500 (examples/ex03_workflow.py), shared-ray (dual) solve"}
    print(json.dumps(out, indent=1))
