"""Build librsk_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_lib"
LIB = OUT_DIR / "librsk_b200.so"
SOURCES = ["rsk_api.cu", "rsk_qmc.cu", "rsk_bvh.cu", "rsk_trace.cu", "rsk_stats.cu", "rsk_prepare.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: cannot build librsk_b200.so")
    return exe


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "raystrack_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: Path | None = None) -> Path:
    """Compile and link.  ``defines`` (e.g. ("RSK_PRMT=0",)) + ``out`` build an experimental variant next to the
    product library (scripts/kernel_variants.py); the product itself is always built without defines."""
    target = Path(out) if out else LIB
    if not force and not defines and not _stale():
        return LIB
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OUT_DIR / ("obj" if not defines else "obj_" + "_".join(d.replace("=", "") for d in defines))
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None

    def compile_one(src: str) -> Path:
        obj = obj_dir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        if host_cxx:
            cmd[1:1] = ["-ccbin", host_cxx]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
            "-o", str(target), *map(str, objs)]
    if host_cxx:
        link[1:1] = ["-ccbin", host_cxx]
    r = subprocess.run(link, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
