"""Wall time and GPU time of the scene upload + device-side preparation + BVH build (C5 by default), warm context."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from raystrack_b200 import _native, synthetic                      # noqa: E402
from raystrack_b200.prepared import PreparedSolver                 # noqa: E402

meshes = synthetic.urban_block(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
ctx = _native.Context.for_device()
for rep in range(5):
    ps = PreparedSolver(meshes)
    ps._flat()
    ctx.synchronize()
    t = time.perf_counter()
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx)
    ctx.synchronize()
    wall = 1e3 * (time.perf_counter() - t)
    info = sc.info()
    print(f"rep {rep}: upload+prepare+build wall {wall:.2f} ms, build kernels+syncs {info['build_us'] / 1e3:.2f} ms, "
          f"{info['n_nodes']} nodes, depth {info['depth']}")
    ps.clear_device_cache()
