"""cProfile of one public-API call on the C5 scene (host overheads around the kernels)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import raystrack_b200 as rb                      # noqa: E402
from raystrack_b200 import main as M, synthetic  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
meshes = synthetic.urban_block(20)
M._log = lambda m: None
p = rb.MatrixParams(samples=4, rays=64, seed=1, bvh="builtin", reciprocity=False, max_iters=iters, min_iters=iters, tol=0.0)
t = time.perf_counter(); rb.view_factor_matrix(meshes, p); print("first call (context, QMC tables)", time.perf_counter() - t)
for rep in range(3):
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    rb.view_factor_matrix(meshes, p)                 # from the bare mesh list, as bench.py's e2e leg does
    pr.disable()
    print(f"--- call {rep + 2}: {time.perf_counter() - t:.3f}s  phases {({k: round(1e3 * v, 1) for k, v in M.LAST_TIMING.items()})}")
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
