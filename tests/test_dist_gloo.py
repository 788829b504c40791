"""N>1 host path on CPU: two gloo ranks run the sharded solve driver against a deterministic fake device and must
reproduce the single-rank result (integer tallies make the result independent of the rank count)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


class FakeSolve:
    """Stands in for _native.Solve: tallies are a pure function of (emitter, ray index range, iteration)."""

    def __init__(self, ctx, scene, em, emit_ids, surf_active, cp_table, rot_base, *, max_iters, min_iters, interval, tol_mode, tol,
                 emit_sid=None, min_sid=None, sky=False, discrete=False, ray_range=None):
        import torch
        self.ids = np.asarray(emit_ids)
        self.n_local = len(self.ids)
        self.rr = np.asarray(ray_range).reshape(-1, 2)
        self.n_hist = (145 if discrete else 1) if sky else 2 * surf_active.shape[1]
        self.max_iters = max_iters
        self.it = np.zeros(self.n_local, np.int64)
        self.iter_t = torch.zeros(max(self.n_local, 1) * self.n_hist, dtype=torch.int64)
        self.total = np.zeros((self.n_local, self.n_hist), np.int64)
        self.n_once = scene.n_once

    def _trace(self):
        t = self.iter_t.numpy().reshape(-1, self.n_hist)
        for k, e in enumerate(self.ids):
            if self.it[k] >= self.max_iters:
                continue
            b, en = self.rr[k]
            idx = np.arange(-(-b // 997) * 997, en, 997)  # a sparse, slice-independent "ray sample"
            bins = (idx * 31 + e * 7 + self.it[k]) % self.n_hist
            t[k] += np.bincount(bins, minlength=self.n_hist)

    def _fold(self):
        t = self.iter_t.numpy().reshape(-1, self.n_hist)
        for k in range(self.n_local):
            if self.it[k] >= self.max_iters:
                continue
            self.total[k] += t[k]
            self.it[k] += 1
        t[:] = 0

    def step(self, n):
        for _ in range(n):
            self._trace()
            self._fold()
        return self.poll()

    enqueue_trace = _trace
    enqueue_fold = _fold

    def poll(self):
        return int(np.sum(self.it < self.max_iters))

    def device_iter_tallies(self):
        return self.iter_t, self.n_hist

    def read_block(self):
        return self.total.copy(), self.it.astype(np.int32), self.it * self.n_once[self.ids]

    def close(self):
        pass


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from raystrack_b200 import _native, dist as D, main as M
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    _native.Solve = FakeSolve
    D.attach_tally_tensor = lambda solve, n_jobs, device=0: (solve.iter_t[: solve.n_hist * n_jobs] if n_jobs and solve.n_local else None)

    class Obj:
        pass

    n_once = np.array([40000, 12288, 500000, 8192, 65536, 20000, 3000000, 16384], np.int64)
    scene, em, ctx = Obj(), Obj(), Obj()
    scene.native, em.native, ctx.device = scene, em, 0
    scene.n_once = n_once
    n = len(n_once)
    active = np.ones((n, n), np.uint8)
    todo = [0, 1, 2, 4, 5, 6, 7]
    tallies, iters, totals = M._solve_sharded(ctx, scene, em, todo, list(n_once), active, np.zeros((20, 7), np.float32),
                                              max_iters=5, min_iters=2, interval=1, tol_mode="stderr", tol=0.0,
                                              emit_sid=np.arange(n), min_sid=np.zeros(n, np.int64))
    np.savez(Path(out_dir) / f"r{world}_{rank}.npz", tallies=tallies, iters=iters, totals=totals,
             shared=sum(1 for j in M.plan_shards(todo, list(n_once), world)[rank] if j[3]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    _worker(0, 1, 0, tmp_path)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    one = np.load(tmp_path / "r1_0.npz")
    a, b = np.load(tmp_path / "r2_0.npz"), np.load(tmp_path / "r2_1.npz")
    assert int(a["shared"]) >= 1                                   # the 3M-ray emitter is ray-split over both ranks
    for key in ("tallies", "iters", "totals"):
        assert np.array_equal(a[key], b[key])                      # every rank returns the full result
        assert np.array_equal(a[key], one[key])                    # ... equal to the single-GPU result
    assert one["tallies"].sum() > 0 and np.array_equal(one["iters"], [5, 5, 5, 0, 5, 5, 5, 5])


def test_allreduce_helpers_single_process():
    from raystrack_b200 import dist as D
    a = np.arange(6, dtype=np.int64).reshape(2, 3)
    D.allreduce_sum_([a])                                          # no group: identity
    assert np.array_equal(a, np.arange(6).reshape(2, 3))
    assert D.max_over_ranks(3.5) == 3.5


def test_nccl_id_file_rendezvous(tmp_path, monkeypatch):
    """dist._exchange_id_file: rank 0 publishes the 128-byte id atomically, the other ranks wait for the complete file."""
    import threading
    from raystrack_b200 import _native, dist as D
    uid = bytes(range(128))
    monkeypatch.setattr(_native.Context, "comm_unique_id", staticmethod(lambda: uid))
    path = tmp_path / "id.bin"
    got = {}
    t = threading.Thread(target=lambda: got.setdefault("r1", D._exchange_id_file(path, 1, 10.0)))
    t.start()
    assert D._exchange_id_file(path, 0, 10.0) == uid
    t.join(15)
    assert got["r1"] == uid and not path.with_suffix(".bin.tmp").exists()
    import pytest
    with pytest.raises(TimeoutError):
        D._exchange_id_file(tmp_path / "never.bin", 1, 0.05)
    assert D.native_comm_active() is False and D.native_comm_env() == (0, 1)
