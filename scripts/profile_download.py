"""Where the milliseconds of the result download go (C5, one iteration): dense block vs pinned view vs device CSR."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from raystrack_b200 import _native, synthetic                      # noqa: E402
from raystrack_b200.main import _csr_from_dense, _rotation_table, _surface_masks    # noqa: E402
from raystrack_b200.prepared import PreparedSolver                 # noqa: E402

meshes = synthetic.urban_block(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
ps = PreparedSolver(meshes)
ctx = _native.Context.for_device(0)
sc = ps.get_device_scene(use_bvh=True, ctx=ctx).native
em = ps.get_device_emitters(samples=4, rays=64, flip_faces=False, ctx=ctx).native
ems = ps.get_emitter_summaries(samples=4, rays=64, flip_faces=False, ctx=ctx)
n = len(meshes)
active = _surface_masks(ems, *ps.get_mesh_bounds())
ids = np.arange(n, dtype=np.int32)
solve = _native.Solve(ctx, sc, em, ids, active, _rotation_table(1, n, 4), ids.copy(), max_iters=2, min_iters=2, interval=1,
                      tol_mode="stderr", tol=0.0, emit_sid=ids, min_sid=np.zeros(n, np.int32))
solve.step(2)


def T(label, f, reps=4):
    out = None
    for r in range(reps):
        ctx.synchronize()
        t = time.perf_counter()
        out = f()
        dt = time.perf_counter() - t
        print(f"{label:28s} run {r}: {1e3 * dt:8.2f} ms", flush=True)
    return out


dense, it, tot = T("read_block (pageable)", solve.read_block)
T("read_block_view (pinned)", solve.read_block_view)
T("read_counters", solve.read_counters)
csr = T("read_csr (device)", solve.read_csr)
T("_csr_from_dense (host)", lambda: _csr_from_dense(dense, tot))
print("nnz", int(csr[0][-1]), "of", dense.size)
solve.close()
