"""Build experimental variants of librsk_b200 (compile-time knobs of the trace kernel) for A/B runs on the GPU:

    python scripts/kernel_variants.py build            # here (no GPU needed)
    python scripts/kernel_variants.py run [--iters 3]   # on the GPU box: perf_c5.py once per variant
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
VARIANTS = {      # compile-time knobs of csrc/rsk_trace.cu(h) and csrc/rsk_bvh.cu; the product is built with the defaults
    "shipped": (),
    "counters": ("RSK_COUNTERS=1",),
    "tri_single": ("RSK_TRI_PAIRS=0",),
    "sky_single": ("RSK_TRI_PAIRS_SKY=0",),
    "ctas3": ("RSK_MIN_CTAS_PER_SM=3",),
    "ctas3_tri3": ("RSK_MIN_CTAS_PER_SM=3", "RSK_TRI_PAIRS=3"),
    "ctas3_tri4": ("RSK_MIN_CTAS_PER_SM=3", "RSK_TRI_PAIRS=4"),
    "tri3": ("RSK_TRI_PAIRS=3",),
    "tri4": ("RSK_TRI_PAIRS=4",),
    "stack4": ("RSK_SMEM_STACK_N=4",),
    "refill20": ("RSK_REFILL_BELOW=20",),
    "refill28": ("RSK_REFILL_BELOW=28",),
    "postpone6": ("RSK_POSTPONE=6",),
    "postpone16": ("RSK_POSTPONE=16",),
    "postpone24": ("RSK_POSTPONE=24",),
    "idle2": ("RSK_POSTPONE_IDLE=2",),
    "idle8": ("RSK_POSTPONE_IDLE=8",),
    "stack6": ("RSK_SMEM_STACK_N=6",),
}
OUT = ROOT / "raystrack_b200" / "_lib" / "variants"

if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "build"
    only = [a for a in sys.argv[2:] if not a.startswith("--")] if cmd == "build" else []
    if cmd == "build":
        from raystrack_b200 import _build
        OUT.mkdir(parents=True, exist_ok=True)
        for name, defs in VARIANTS.items():
            if only and name not in only:
                continue
            print(name, _build.build(force=True, defines=defs or ("RSK_VARIANT_BUILD=1",), out=OUT / f"librsk_{name}.so"), flush=True)
    else:
        extra = sys.argv[2:] or ["--iters", "3"]
        only_run = os.environ.get("RSK_VARIANTS", "").split(",") if os.environ.get("RSK_VARIANTS") else None
        for name in VARIANTS:
            if only_run and name not in only_run:
                continue
            lib = OUT / f"librsk_{name}.so"
            if not lib.exists():
                continue
            env = dict(os.environ, RSK_LIB=str(lib))
            r = subprocess.run([sys.executable, str(ROOT / "scripts" / "perf_c5.py"), *extra], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if "Grays" in l]
            print(f"{name:28s} {line[-1] if line else r.stderr[-300:]}", flush=True)
