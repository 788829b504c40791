"""Cold-start cost of the first preparation in a process (context, memory pool, Halton tables, BVH): C5."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from raystrack_b200 import _native, synthetic                      # noqa: E402
from raystrack_b200.prepared import PreparedSolver                 # noqa: E402

meshes = synthetic.urban_block(20)
t0 = time.time()
ctx = _native.Context.for_device(0)
t1 = time.time()
for k in range(3):
    t = time.time()
    ps = PreparedSolver(meshes)
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx)
    t_s = time.time()
    em = ps.get_device_emitters(samples=4, rays=64, flip_faces=False, ctx=ctx)
    ctx.synchronize()
    print(f"pass {k}: scene {1e3 * (t_s - t):.1f} ms, emitters {1e3 * (time.time() - t_s):.1f} ms, BVH device {sc.info()['build_us'] / 1e3:.2f} ms")
    ps.clear_device_cache()
print(f"context creation {1e3 * (t1 - t0):.1f} ms")
