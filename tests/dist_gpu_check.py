"""Multi-GPU parity check, launched by torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_gpu_check.py

Every rank solves its shard (whole emitters + ray slices of oversized emitters), tallies are all-reduced over NCCL,
and every rank must end with the single-GPU / reference result: integer tallies make it independent of N."""
import json
import os
import sys
from pathlib import Path


ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import raystrack_b200 as rb  # noqa: E402
from raystrack_b200 import dist as D, main as M, synthetic  # noqa: E402
from scenes import scene_for  # noqa: E402


def main():
    rank, world = D.init_from_env("nccl")
    M._log = lambda msg: None
    solves = json.loads((ROOT / "tests" / "golden" / "solves.json").read_text())
    worst_all = 0.0
    for case in ("C2_canyon_ex01", "C2b_canyon_delta_norecip", "C3_canyon_sky_discrete", "U3_urban_matrix_bvh", "U3_urban_matrix_recip", "U3_urban_sky"):
        g = solves[case]
        p = dict(g["params"])
        if "discrete" in p:
            res = rb.view_factor_to_tregenza_sky(scene_for(case), rb.SkyParams(**p))
        else:
            res = rb.view_factor_matrix(scene_for(case), rb.MatrixParams(**p))
        worst = 0.0
        for name, row in g["result"].items():
            for key in set(row) | set(res[name]):
                worst = max(worst, abs(res[name].get(key, 0.0) - row.get(key, 0.0)))
        worst_all = max(worst_all, worst)
        assert worst <= 2e-5, (case, worst)
        print(f"[rank {rank}/{world}] {case}: max |dF| vs reference = {worst:.2e}", flush=True)
    # a scene with one oversized emitter (the ground) so that the ray-split path and its per-iteration all-reduce run
    meshes = synthetic.urban_block(4, 4, 8, 0)
    prm = rb.MatrixParams(samples=4, rays=32, seed=2, bvh="builtin", reciprocity=False, max_iters=12, min_iters=3, tol=5e-4)
    n_once = [int(e.n_cells * 32) for e in rb.PreparedSolver(meshes).get_emitters(samples=4, rays=32, flip_faces=False)]
    plan = M.plan_shards(list(range(len(meshes))), n_once, world)
    n_shared = sum(1 for j in plan[rank] if j[3])
    res = rb.view_factor_matrix(meshes, prm)
    blob = json.dumps(res, sort_keys=True, default=float)
    gathered = [None] * world
    torch.distributed.all_gather_object(gathered, blob)
    assert all(b == gathered[0] for b in gathered), "ranks disagree"
    # the same solve in memory-budgeted chunks (several solves per rank feed the device-side reduction)
    os.environ["RSK_SOLVE_MEMORY_MB"] = "0.02"
    assert json.dumps(rb.view_factor_matrix(meshes, prm), sort_keys=True, default=float) == blob, "chunked solve differs"
    del os.environ["RSK_SOLVE_MEMORY_MB"]
    # shared-ray workflow: each rank traces its emitters once with the dual kernel; must equal the two separate solves
    mp_ = rb.MatrixParams(samples=4, rays=32, seed=2, bvh="builtin", reciprocity=True, max_iters=12, min_iters=3, tol=5e-4)
    sp_ = rb.SkyParams(samples=4, rays=32, seed=2, bvh="builtin", max_iters=8, min_iters=3, tol=1e-3, discrete=True)
    assert M.outside_workflow_shareable(mp_, sp_)
    both = M.view_factor_matrix_and_sky(meshes, matrix_params=mp_, sky_params=sp_)
    assert both[0] == rb.view_factor_matrix(meshes, mp_), "dual solve (matrix side) differs from the separate solve"
    assert both[1] == rb.view_factor_to_tregenza_sky(meshes, sp_), "dual solve (sky side) differs from the separate solve"
    os.environ["RSK_SOLVE_MEMORY_MB"] = "0.02"
    assert M.view_factor_matrix_and_sky(meshes, matrix_params=mp_, sky_params=sp_) == both, "chunked dual solve differs"
    del os.environ["RSK_SOLVE_MEMORY_MB"]
    print(f"[rank {rank}/{world}] shared-ray workflow == separate solves", flush=True)
    # the same ray-split solve unsharded on this rank's GPU: bit-identical
    M._DIST_OVERRIDE = (0, 1)
    try:
        assert json.dumps(rb.view_factor_matrix(meshes, prm), sort_keys=True, default=float) == blob, "multi-GPU result differs from the single-GPU result"
    finally:
        M._DIST_OVERRIDE = None
    if rank == 0:
        print(f"[rank 0] ray-split solve ({n_shared} shared emitters per rank) == single-GPU result, bit for bit", flush=True)
        print(f"DIST_CHECK_OK world={world} shared={n_shared} worst={worst_all:.2e}", flush=True)
    D.shutdown_native()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
