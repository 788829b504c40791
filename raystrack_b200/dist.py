"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the
CPU tests).  The path shards by emitter (matrix rows are independent, reference main.py:1758-1939): the BVH and
triangles are replicated, each rank solves its emitters without any data-path collective, and the integer tally
blocks are summed once at the end -- integer sums make the result independent of the GPU count."""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np


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


def allreduce_sum_(arrays: Sequence[np.ndarray], device: int = 0) -> None:
    """In-place SUM all-reduce of int64 NumPy arrays over the default group (one flat NCCL all-reduce)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
        return
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


def nccl_active() -> bool:
    """True when a torch.distributed NCCL group with more than one rank is running."""
    try:
        import torch.distributed as dist
        return bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl")
    except Exception:
        return False


class _DeviceBlock:
    """A raw device allocation of the library presented through ``__cuda_array_interface__`` (int64, C order)."""

    def __init__(self, pointer: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<i8", "data": (int(pointer), False),
                                         "version": 3, "strides": None}


class DeviceReducer:
    """Sums the per-rank tally blocks on the GPUs: every rank scatters the rows of its solves into a zeroed
    ``[n_rows, n_cols]`` int64 device tensor, one NCCL all-reduce adds them up over NVLink, and the result comes back
    to the host in a single copy into pinned memory.  (Going through NumPy costs three more 64 MB pageable copies per
    call at the bench scene.)  Needs the library context to share torch's current stream."""

    def __init__(self, n_rows: int, n_cols: int, device: int):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", int(device))
        self.full = torch.zeros((int(n_rows), int(n_cols)), dtype=torch.int64, device=self.device)

    def add_rows(self, rows: np.ndarray, pointer: int, n_local: int, keep: np.ndarray) -> None:
        """full[rows[keep]] = block[keep], block = the solve's int64 [n_local, n_cols] totals at ``pointer``."""
        torch = self.torch
        if n_local == 0 or not keep.any():
            return
        block = torch.as_tensor(_DeviceBlock(pointer, (n_local, self.full.shape[1])), device=self.device)
        dst = torch.as_tensor(np.asarray(rows, np.int64)[keep], device=self.device)
        if keep.all():
            self.full.index_copy_(0, dst, block)
        else:
            self.full.index_copy_(0, dst, block[torch.as_tensor(np.nonzero(keep)[0], device=self.device)])

    def finish(self) -> np.ndarray:
        import torch.distributed as dist
        torch = self.torch
        dist.all_reduce(self.full, op=dist.ReduceOp.SUM)
        host = torch.empty(self.full.shape, dtype=torch.int64, pin_memory=True)
        host.copy_(self.full, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()


def barrier() -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: int = 0) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def attach_tally_tensor(solve, n_jobs: int, device: int = 0):
    """Give ``solve`` a torch-owned per-iteration tally buffer and return the tensor view of its first ``n_jobs``
    rows (the ray-split jobs, which come first) for in-place NCCL all-reduces; None when there is nothing to reduce.
    The returned object keeps the whole buffer alive."""
    import torch
    _, n_per_job = solve.device_iter_tallies()
    if solve.n_local == 0 or n_jobs <= 0:
        return None
    full = torch.zeros(solve.n_local * n_per_job, dtype=torch.int64, device=f"cuda:{device}")
    solve.set_iter_tally_buffer(full.data_ptr(), full.numel())
    solve._tally_tensor = full
    return full[: n_jobs * n_per_job]


def all_reduce_device_(tensor, device: int = 0) -> None:
    """SUM all-reduce of a device tensor on torch's current stream (NCCL over NVLink)."""
    import torch.distributed as dist
    if tensor is None:
        return
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
