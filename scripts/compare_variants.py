"""Which rays does a kernel variant resolve differently from the shipped kernel?  (C5 scene, one iteration.)

    python scripts/compare_variants.py dump <name>         # with RSK_LIB set: tallies -> $TMPDIR/tallies_<name>.npz
    python scripts/compare_variants.py rays <name> <emitters...>   # per-ray results of some emitters -> $TMPDIR/rays_<name>.npz
    python scripts/compare_variants.py diff A B            # drives the steps above in sub-processes and reports
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"
TMP = Path(os.environ.get("TMPDIR", "/tmp"))        # the tally / per-ray dumps are large and only needed for the comparison
LIBS = ROOT / "raystrack_b200" / "_lib" / "variants"


def setup():
    from raystrack_b200 import _native, synthetic
    from raystrack_b200.main import _rotation_table, _surface_masks
    from raystrack_b200.prepared import PreparedSolver
    meshes = synthetic.urban_block(20)
    ps = PreparedSolver(meshes)
    ems = ps.get_emitters(samples=4, rays=64, flip_faces=False)
    ctx = _native.Context.for_device(0)
    sc = ps.get_device_scene(use_bvh=True, ctx=ctx).native
    em = ps.get_device_emitters(samples=4, rays=64, flip_faces=False, ctx=ctx).native
    n = len(meshes)
    active = _surface_masks(ems, *ps.get_mesh_bounds())
    return _native, ctx, sc, em, n, active, _rotation_table(1, n, 1)


if __name__ == "__main__":
    cmd = sys.argv[1]
    OUT.mkdir(exist_ok=True)
    if cmd == "dump":
        _native, ctx, sc, em, n, active, table = setup()
        ids = np.arange(n, dtype=np.int32)
        solve = _native.Solve(ctx, sc, em, ids, active, table, ids.copy(), max_iters=1, min_iters=1, interval=1, tol_mode="stderr",
                              tol=0.0, emit_sid=ids, min_sid=np.zeros(n, np.int32))
        solve.step(1)
        hf, hb, it, tot, _, _ = solve.read_matrix()
        np.savez(TMP / f"tallies_{sys.argv[2]}.npz", hf=hf, hb=hb)
    elif cmd == "rays":
        _native, ctx, sc, em, n, active, table = setup()
        out = {}
        for e in map(int, sys.argv[3:]):
            o, d, hit, front = _native.trace_rays(ctx, sc, em, e, active[e], e, 0, table[e], want_rays=True)
            out[f"o{e}"], out[f"d{e}"], out[f"h{e}"], out[f"f{e}"] = o, d, hit, front
        np.savez(TMP / f"rays_{sys.argv[2]}.npz", **out)
    else:
        a, b = sys.argv[2], sys.argv[3]
        for name in (a, b):
            subprocess.run([sys.executable, __file__, "dump", name], env=dict(os.environ, RSK_LIB=str(LIBS / f"librsk_{name}.so")), check=True)
        ta, tb = np.load(TMP / f"tallies_{a}.npz"), np.load(TMP / f"tallies_{b}.npz")
        df, db = ta["hf"] - tb["hf"], ta["hb"] - tb["hb"]
        rows = np.unique(np.concatenate([np.nonzero(df)[0], np.nonzero(db)[0]]))
        moved = (np.abs(df).sum() + np.abs(db).sum()) // 2
        print(f"{a} vs {b}: {int(moved)} of {int(ta['hf'].sum() + ta['hb'].sum())} hits tallied differently, in {rows.size} emitters; "
              f"row totals equal: {bool(np.array_equal(ta['hf'].sum(1) + ta['hb'].sum(1), tb['hf'].sum(1) + tb['hb'].sum(1)))}", flush=True)
        small = [int(r) for r in rows if r != 2000][:24]          # the ground emitter (45 M rays) is left out of the per-ray dump
        if small:
            for name in (a, b):
                subprocess.run([sys.executable, __file__, "rays", name, *map(str, small)],
                               env=dict(os.environ, RSK_LIB=str(LIBS / f"librsk_{name}.so")), check=True)
            ra, rb = np.load(TMP / f"rays_{a}.npz"), np.load(TMP / f"rays_{b}.npz")
            keep = {}
            for e in small:
                bad = np.nonzero((ra[f"h{e}"] != rb[f"h{e}"]) | (ra[f"f{e}"] != rb[f"f{e}"]))[0]
                print(f"emitter {e}: {bad.size} rays differ", [(int(k), int(ra[f'h{e}'][k]), int(ra[f'f{e}'][k]), int(rb[f'h{e}'][k]), int(rb[f'f{e}'][k])) for k in bad[:4]], flush=True)
                keep[f"k{e}"], keep[f"o{e}"], keep[f"d{e}"] = bad, ra[f"o{e}"][bad], ra[f"d{e}"][bad]
                keep[f"a{e}"] = np.stack([ra[f"h{e}"][bad], ra[f"f{e}"][bad]], 1)
                keep[f"b{e}"] = np.stack([rb[f"h{e}"][bad], rb[f"f{e}"][bad]], 1)
            np.savez(OUT / f"diff_{a}_{b}.npz", **keep)
