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


def test_two_gpu_solves_through_the_c_abi_alone(tmp_path):
    """rsk_comm_init / rsk_allreduce_i64 / rsk_tally_block_*: two processes, no torch, id passed through a file."""
    from raystrack_b200 import _native
    n = _native.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (visible: {n})")
    id_file = tmp_path / "nccl_id.bin"
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "comm_c_abi_check.py"), str(r), "2", str(id_file)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=300)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"COMM_C_ABI_OK rank={r} world=2" in out, out[-3000:]
