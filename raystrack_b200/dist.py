"""Multi-GPU plumbing: one process per GPU.  The path shards by emitter (matrix rows are independent, reference
main.py:1758-1939): the BVH and triangles are replicated, each rank solves its emitters without any data-path
collective, and the integer tally blocks are summed once at the end -- integer sums make the result independent of
the GPU count.  Emitters that alone exceed a rank's fair share are split by ray range and their per-iteration
tallies summed before every statistics update.

All GPU collectives run inside librsk_b200 (``rsk_comm_init`` / ``rsk_allreduce_i64`` / ``rsk_tally_block_*``: its own
NCCL communicator, on the context's stream, ordered with the kernels).  ``torch.distributed`` is optional: when a
process group exists (torchrun) it is only used to hand the 128-byte NCCL id to the other ranks; without torch the id
travels through a file (``init_native(id_file=...)`` / ``RSK_COMM_FILE``).  The host-tensor helpers at the bottom
serve the CPU tests (gloo, ``tests/test_dist_gloo.py``)."""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

from . import _native

_NATIVE: Optional[dict] = None         # {"ctx": Context, "rank": int, "world": int}
_NATIVE_FAILED: Optional[str] = None   # why the library communicator could not be created (tried once per process)


# --------------------------------------------------------------------------------------------- torch process group

def init_from_env(backend: str | None = None) -> tuple[int, int]:
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 and "RANK" not in os.environ:
        return 0, 1
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def _torch_group() -> tuple[int, int]:
    if "torch" not in sys.modules:          # a process group cannot exist unless the caller imported torch: do not pay for it
        return 0, 1
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def nccl_active() -> bool:
    """True when a torch.distributed NCCL group with more than one rank is running."""
    if _torch_group()[1] <= 1:
        return False
    import torch.distributed as dist
    return dist.get_backend() == "nccl"


# --------------------------------------------------------------------------------------------- library communicator

def _find_nccl_library() -> None:
    """Point the library at PyTorch's bundled NCCL when nothing else is configured (it prefers a copy that is already
    loaded into the process, then $RSK_NCCL_LIBRARY, then the system's libnccl.so.2)."""
    if os.environ.get("RSK_NCCL_LIBRARY"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations or []) if spec else []:
            cand = Path(base) / "lib" / "libnccl.so.2"
            if cand.exists():
                os.environ["RSK_NCCL_LIBRARY"] = str(cand)
                return
    except Exception:
        pass


def _exchange_id_file(path: Path, rank: int, timeout: float) -> bytes:
    """Rank 0 writes the id (atomically), the others wait for the file."""
    if rank == 0:
        uid = _native.Context.comm_unique_id()
        tmp = path.with_suffix(path.suffix + ".tmp")
        tmp.write_bytes(uid)
        os.replace(tmp, path)
        return uid
    t0 = time.time()
    while time.time() - t0 < timeout:
        if path.exists() and path.stat().st_size == _native.COMM_ID_BYTES:
            return path.read_bytes()
        time.sleep(0.01)
    raise TimeoutError(f"NCCL id file {path} did not appear within {timeout:.0f} s")


def init_native(rank: Optional[int] = None, world: Optional[int] = None, *, device: Optional[int] = None,
                unique_id: Optional[bytes] = None, id_file: Optional[str] = None, timeout: float = 120.0) -> tuple[int, int]:
    """Create this process's library communicator (collective: every rank calls it).  The NCCL id comes from
    ``unique_id``, else is created by rank 0 and passed through the torch process group if one exists, else through
    ``id_file`` / ``$RSK_COMM_FILE`` on a shared file system.  rank/world default to the torch group or RANK/WORLD_SIZE."""
    global _NATIVE
    if _NATIVE is not None:
        return _NATIVE["rank"], _NATIVE["world"]
    t_rank, t_world = _torch_group()
    if rank is None:
        rank = t_rank if t_world > 1 else int(os.environ.get("RANK", "0"))
    if world is None:
        world = t_world if t_world > 1 else int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    _find_nccl_library()
    ctx = _native.Context.for_device(device)
    if unique_id is None:
        id_file = id_file or os.environ.get("RSK_COMM_FILE")
        if t_world == world and t_world > 1:
            import torch.distributed as dist
            box = [_native.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        elif id_file:
            unique_id = _exchange_id_file(Path(id_file), rank, timeout)
        else:
            raise RuntimeError("init_native: no way to share the NCCL id (no torch process group, no id_file / RSK_COMM_FILE)")
    ctx.comm_init(unique_id, rank, world)
    _NATIVE = {"ctx": ctx, "rank": int(rank), "world": int(world)}
    return int(rank), int(world)


def shutdown_native() -> None:
    global _NATIVE
    if _NATIVE is not None:
        _NATIVE["ctx"].comm_destroy()
        _NATIVE = None


def native_comm_active() -> bool:
    return _NATIVE is not None


def native_comm_env() -> tuple[int, int]:
    return (_NATIVE["rank"], _NATIVE["world"]) if _NATIVE else (0, 1)


def native_comm_context(create: bool = False) -> Optional[_native.Context]:
    """The context that owns the library communicator.  ``create``: join it now if a torch process group with more
    than one rank exists and this process sees a GPU (tried once; None when NCCL is unavailable)."""
    global _NATIVE_FAILED
    if _NATIVE is not None:
        return _NATIVE["ctx"]
    if not create or _NATIVE_FAILED is not None:
        return None
    if _torch_group()[1] <= 1:
        return None
    try:
        if _native.device_count() <= 0:
            raise RuntimeError("no CUDA device")
        init_native()
    except Exception as e:      # noqa: BLE001 -- the caller falls back to whole-emitter sharding + host reduction
        _NATIVE_FAILED = f"{type(e).__name__}: {e}"
        return None
    return _NATIVE["ctx"] if _NATIVE else None


# --------------------------------------------------------------------------------------------- helpers on top of either

def barrier() -> None:
    if _NATIVE is not None:
        _NATIVE["ctx"].allreduce_host(np.zeros(1, np.int64))
        return
    if _torch_group()[1] > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(value: float, device: int = 0) -> float:
    """MAX over the ranks of a non-negative number (timings, "jobs still running")."""
    if _NATIVE is not None:
        # the IEEE-754 bit pattern of a non-negative double is monotonic in its value: an int64 MAX does it
        bits = np.array([max(float(value), 0.0)], np.float64).view(np.int64).copy()
        return float(_NATIVE["ctx"].allreduce_host(bits, "max").view(np.float64)[0])
    if _torch_group()[1] <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum_(arrays: Sequence[np.ndarray], device: int = 0) -> None:
    """In-place SUM all-reduce of (small) int64 NumPy arrays: iteration counters, ray totals.  Large tally blocks
    stay on the device (``_native.TallyBlock``)."""
    if _NATIVE is not None:
        flat = np.concatenate([np.ascontiguousarray(a, np.int64).reshape(-1) for a in arrays])
        for lo in range(0, flat.size, 4096):
            part = np.ascontiguousarray(flat[lo:lo + 4096])
            flat[lo:lo + 4096] = _NATIVE["ctx"].allreduce_host(part)
        out = flat
    else:
        if _torch_group()[1] <= 1:
            return
        import torch
        import torch.distributed as dist
        flat = np.concatenate([np.ascontiguousarray(a, np.int64).reshape(-1) for a in arrays])
        t = torch.from_numpy(flat)
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out = t.cpu().numpy()
    pos = 0
    for a in arrays:
        n = a.size
        a[...] = out[pos:pos + n].reshape(a.shape)
        pos += n


# --------------------------------------------------------------------------------------------- host-tensor exchange (CPU tests)

def attach_tally_tensor(solve, n_jobs: int, device: int = 0):
    """Host stand-in path (tests/test_dist_gloo.py): the per-iteration tallies of the first ``n_jobs`` jobs of a solve
    whose tallies live in a torch tensor; None when there is nothing to reduce."""
    tensor, n_per_job = solve.device_iter_tallies()
    if solve.n_local == 0 or n_jobs <= 0:
        return None
    return tensor[: n_jobs * n_per_job]


def all_reduce_tensor_(tensor) -> None:
    import torch.distributed as dist
    if tensor is None:
        return
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
