"""Multi-GPU parity on a box with >= 2 B200s: torchrun launches tests/dist_gpu_check.py (one rank per GPU, NCCL)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_two_gpu_sharded_solves_match_reference():
    from raystrack_b200 import _native
    n = _native.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (visible: {n}); the N>1 host path is covered on CPU by test_dist_gloo.py")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_CHECK_OK world=2" in r.stdout
