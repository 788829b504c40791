"""Quick throughput probe on the synthetic urban scene (not the bench contract; see bench.py)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from raystrack_b200 import _native, synthetic                      # noqa: E402
from raystrack_b200.main import _rotation_table, _surface_masks    # noqa: E402
from raystrack_b200.prepared import PreparedSolver                 # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=20)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--samples", type=int, default=4)
ap.add_argument("--rays", type=int, default=64)
ap.add_argument("--sky", action="store_true")
ap.add_argument("--terrain", action="store_true", help="the terrain + small objects scene instead of the urban block")
ap.add_argument("--ground-grid", type=int, default=32, help="quads per side of the ground mesh (32 = C5)")
ap.add_argument("--face-grid", type=int, default=16, help="quads per side of every wall / roof (16 = C5)")
ap.add_argument("--recip", action="store_true", help="reciprocity schedule: emitter i ignores meshes j <= i")
ap.add_argument("--world", type=int, default=1, help="time the shard rank --rank of plan_shards(world) would get (no collectives)")
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--cost", action="store_true", help="weight the shard plan by the measured cost per ray (as the public call does)")
args = ap.parse_args()

t = time.time()
meshes = synthetic.terrain_with_objects() if args.terrain else synthetic.urban_block(args.side, args.face_grid, args.ground_grid)
ps = PreparedSolver(meshes)
print(f"meshes {len(meshes)} tris {ps.total_faces} gen {time.time()-t:.2f}s", flush=True)
t = time.time()
ems = ps.get_emitters(samples=args.samples, rays=args.rays, flip_faces=False)
hs = ps.get_scene(use_bvh=True)
print(f"host prep {time.time()-t:.2f}s", flush=True)
ctx = _native.Context.for_device(0)
t = time.time()
sc = ps.get_device_scene(use_bvh=True, ctx=ctx).native
print(f"scene upload+BVH {time.time()-t:.2f}s info {sc.info()}", flush=True)
t = time.time()
em = ps.get_device_emitters(samples=args.samples, rays=args.rays, flip_faces=False, ctx=ctx).native
print(f"emitters+tables {time.time()-t:.2f}s", flush=True)
n = len(meshes)
centers, extents = ps.get_mesh_bounds()
active = _surface_masks(ems, centers, extents)
from raystrack_b200.main import plan_shards                       # noqa: E402
n_once_all = [int(e.n_cells * args.rays) for e in ems]
cost = None
if args.cost:
    from raystrack_b200.main import _emitter_cost_per_ray
    class _W:                     # the device objects as PreparedSolver wraps them
        def __init__(self, native): self.native = native
    t = time.time()
    cost = _emitter_cost_per_ray(ctx, _W(sc), _W(em), list(range(n)), n_once_all, active, _rotation_table(1, n, 4), np.arange(n, dtype=np.int32),
                                 np.zeros(n, np.int32), 0, 1)
    print(f"cost per ray: min {cost.min():.2f} max {cost.max():.2f} ({1e3 * (time.time() - t):.1f} ms)", flush=True)
plan = plan_shards(list(range(n)), n_once_all, args.world, cost_per_ray=cost)[args.rank]
ids = np.asarray([j[0] for j in plan], np.int32)
ranges = np.asarray([[j[1], j[2]] for j in plan], np.int64)
table = _rotation_table(1, n, args.iters + 2)
solve = _native.Solve(ctx, sc, em, ids, active[ids], table, ids.copy(), max_iters=args.iters + 2, min_iters=args.iters + 2, interval=1,
                      tol_mode="stderr", tol=0.0, emit_sid=ids, min_sid=(ids + 1) if args.recip else np.zeros(len(ids), np.int32),
                      sky=args.sky, discrete=True, ray_range=ranges)
solve.step(1)
r0 = solve.rays_traced()
ctx.trace_counters(reset=True)
ctx.timer_start()
solve.step(args.iters)
ms = ctx.timer_stop()
rays = solve.rays_traced() - r0
crc = ""
if not args.sky:
    import zlib
    hf, hb, it, tot, _, _ = solve.read_matrix()
    crc = f"  tallies crc32 {zlib.crc32(hf.tobytes() + hb.tobytes()):08x}"
cnt = ctx.trace_counters()
per_ray = "".join(f" {k}/ray {v / cnt['rays']:.2f}" for k, v in cnt.items() if k != "rays") if cnt["rays"] else ""
print(f"{rays} rays in {ms:.1f} ms -> {rays/ms/1e6:.4f} Grays/s{crc}{per_ray}", flush=True)
if not args.sky:
    print("hit fraction", (hf.sum() + hb.sum()) / tot.sum(), "iters", it[:3], "launches", ctx.launch_count())
