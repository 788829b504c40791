"""cProfile of the public call on C5 (host side of the e2e number): where the milliseconds outside the kernels go."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from raystrack_b200 import MatrixParams, main as M, synthetic, view_factor_matrix     # noqa: E402

M._log = lambda msg: None
import os                                                                             # noqa: E402
if int(os.environ.get("WORLD_SIZE", "1")) > 1:                                        # torchrun: profile the sharded path on rank 0
    from raystrack_b200 import dist as D
    D.init_from_env("nccl")
meshes = synthetic.urban_block(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
prm = MatrixParams(samples=4, rays=64, seed=1, bvh="builtin", reciprocity=False, max_iters=int(os.environ.get('ITERS', '1')), min_iters=int(os.environ.get('ITERS', '1')), tol=0.0)
for _ in range(2):
    t = time.perf_counter()
    view_factor_matrix(meshes, prm)
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"call: {1e3 * (time.perf_counter() - t):.1f} ms", {k: round(1e3 * v, 1) for k, v in M.LAST_TIMING.items()})
pr = cProfile.Profile()
pr.enable()
view_factor_matrix(meshes, prm)
pr.disable()
if int(os.environ.get("RANK", "0")) == 0:
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
