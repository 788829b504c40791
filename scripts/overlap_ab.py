"""A/B of the overlapped row assembly (main._overlap_split) on the C5 public call, any GPU count (torchrun): alternates
RSK_OVERLAP_ASSEMBLY=0/1 inside one process group, wall time per call = max over ranks."""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from raystrack_b200 import MatrixParams, dist as D, main as M, synthetic, view_factor_matrix     # noqa: E402

M._log = lambda msg: None
rank, world = 0, 1
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    rank, world = D.init_from_env("nccl")
meshes = synthetic.urban_block(int(os.environ.get("SIDE", "20")))
iters = int(os.environ.get("ITERS", "40"))
prm = MatrixParams(samples=4, rays=64, seed=1, bvh="builtin", reciprocity=False, max_iters=iters, min_iters=iters, tol=0.0)
view_factor_matrix(meshes, prm)                     # warm-up
times = {"0": [], "1": []}
phases = {}
ref = None
for rep in range(int(os.environ.get("REPS", "3"))):
    for ov in ("0", "1"):
        os.environ["RSK_OVERLAP_ASSEMBLY"] = ov
        if world > 1:
            D.barrier()
        t = time.perf_counter()
        res = view_factor_matrix(meshes, prm)
        dt = D.max_over_ranks(time.perf_counter() - t)
        times[ov].append(1e3 * dt)
        phases[ov] = {k: round(1e3 * v, 1) for k, v in M.LAST_TIMING.items()}
        if rep == 0:
            phases["first" + ov] = phases[ov]
        if ref is None:
            ref = res
        elif res != ref:
            raise SystemExit("results differ between the variants")
        del res
if rank == 0:
    for ov in ("0", "1"):
        print(f"overlap={ov} world={world}: {np.round(times[ov], 1).tolist()} ms, mean {np.mean(times[ov]):.1f}", phases[ov])
        print("   first call of this variant:", phases["first" + ov])
    print("identical results: True")
if world > 1:
    D.barrier()
    D.shutdown_native()
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
