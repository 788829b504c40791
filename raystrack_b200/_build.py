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
SOURCES = ["rsk_api.cu", "rsk_qmc.cu", "rsk_bvh.cu", "rsk_trace.cu", "rsk_stats.cu", "rsk_prepare.cu", "rsk_comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: cannot build librsk_b200.so")
    return exe


HASH_FILE = OUT_DIR / "librsk_b200.hash"


def source_hash() -> str:
    """sha256 (first 16 hex digits) over the CUDA sources and the public header, in name order."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "raystrack_b200.h"]
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def stale() -> bool:
    """True when the library is missing or was built from other sources than the ones in the tree."""
    if not LIB.exists() or not HASH_FILE.exists():
        return True
    return HASH_FILE.read_text().strip() != source_hash()


_stale = stale


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
    digest = source_hash()
    env = dict(os.environ)
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None

    def compile_one(src: str) -> Path:
        obj = obj_dir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], f'-DRSK_SOURCE_HASH="{digest}"', "-c", str(CSRC / src), "-o", str(obj)]
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
            "-o", str(target), *map(str, objs), "-ldl"]
    if host_cxx:
        link[1:1] = ["-ccbin", host_cxx]
    r = subprocess.run(link, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if not defines and out is None:
        HASH_FILE.write_text(digest + "\n")
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
