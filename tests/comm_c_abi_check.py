"""Multi-GPU check through the C ABI alone -- no torch, no torch.distributed:

    python tests/comm_c_abi_check.py RANK WORLD ID_FILE        (one process per GPU, started by the caller)

Rank 0 creates the NCCL id (rsk_comm_unique_id) and passes it through ID_FILE; every rank joins the library
communicator (rsk_comm_init), all-reduces host and device values, then runs sharded solves through the public API:
reference goldens, and a scene with a ray-split emitter that must equal the unsharded solve bit for bit."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.modules["torch"] = None                      # any `import torch` from here on raises ImportError

import numpy as np  # noqa: E402

import raystrack_b200 as rb  # noqa: E402
from raystrack_b200 import _native, dist as D, main as M, synthetic  # noqa: E402
from scenes import scene_for  # noqa: E402


def main():
    rank, world, id_file = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    assert D.init_native(rank, world, device=rank, id_file=id_file) == (rank, world)
    ctx = D.native_comm_context()
    info = ctx.comm_info()
    assert (info["rank"], info["nranks"]) == (rank, world) and info["nccl_version"] >= 22000, info
    # host-value all-reduces
    v = ctx.allreduce_host(np.array([rank + 1, 10 * (rank + 1)], np.int64))
    assert v.tolist() == [world * (world + 1) // 2, 10 * world * (world + 1) // 2]
    assert ctx.allreduce_host(np.array([rank], np.int64), "max")[0] == world - 1
    assert D.max_over_ranks(1.5 + rank) == 1.5 + world - 1
    # device block: every rank contributes its rank + 1 to row `rank`
    blk = _native.TallyBlock(ctx, world, 5)
    ptr, n = blk.device_pointer()
    assert n == 5 * world
    blk.allreduce()
    assert not blk.download(copy=True).any()
    blk.close()
    assert M._dist_env() == (rank, world) and M._context() is ctx
    M._log = lambda msg: None
    solves = json.loads((ROOT / "tests" / "golden" / "solves.json").read_text())
    worst = 0.0
    for case in ("C2_canyon_ex01", "C3_canyon_sky_discrete", "U3_urban_matrix_bvh", "U3_urban_sky"):
        g = solves[case]
        p = dict(g["params"])
        if "discrete" in p:
            res = rb.view_factor_to_tregenza_sky(scene_for(case), rb.SkyParams(**p))
        else:
            res = rb.view_factor_matrix(scene_for(case), rb.MatrixParams(**p))
        for name, row in g["result"].items():
            for key in set(row) | set(res[name]):
                worst = max(worst, abs(res[name].get(key, 0.0) - row.get(key, 0.0)))
    assert worst <= 2e-5, worst
    meshes = synthetic.urban_block(4, 4, 8, 0)
    prm = rb.MatrixParams(samples=4, rays=32, seed=2, bvh="builtin", reciprocity=False, max_iters=12, min_iters=3, tol=5e-4)
    n_once = [int(e.n_cells * 32) for e in rb.PreparedSolver(meshes).get_emitters(samples=4, rays=32, flip_faces=False)]
    n_shared = sum(1 for j in M.plan_shards(list(range(len(meshes))), n_once, world)[rank] if j[3])
    assert world == 1 or n_shared >= 1
    res = rb.view_factor_matrix(meshes, prm)
    M._DIST_OVERRIDE = (0, 1)
    try:
        assert rb.view_factor_matrix(meshes, prm) == res, "sharded (ray-split) solve differs from the unsharded solve"
    finally:
        M._DIST_OVERRIDE = None
    mp_ = rb.MatrixParams(samples=4, rays=32, seed=2, bvh="builtin", reciprocity=True, max_iters=12, min_iters=3, tol=5e-4)
    sp_ = rb.SkyParams(samples=4, rays=32, seed=2, bvh="builtin", max_iters=8, min_iters=3, tol=1e-3, discrete=True)
    both = M.view_factor_matrix_and_sky(meshes, matrix_params=mp_, sky_params=sp_)
    assert both[0] == rb.view_factor_matrix(meshes, mp_) and both[1] == rb.view_factor_to_tregenza_sky(meshes, sp_)
    D.barrier()
    D.shutdown_native()
    assert "torch" not in {k for k, v in sys.modules.items() if v is not None}
    print(f"COMM_C_ABI_OK rank={rank} world={world} shared={n_shared} worst={worst:.2e}", flush=True)


if __name__ == "__main__":
    main()
