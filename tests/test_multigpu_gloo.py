"""N > 1 host path on CPU: two processes, gloo backend, the sharding + all-reduce driver of
hyperion_b200/multigpu.py with the CPU oracle standing in for the per-GPU engine.

What the reference does with MPI (src/mpi/mpi_routines.f90:266-361: rank r seeded seed+r,
reduce of specific_energy_sum, sync of energy_current, broadcast of the scaled grid) must come
out of ONE all-reduce of [deposit grid | scalars]: every rank ends the iteration with the same
specific_energy, and it equals the emulated two-rank reference run.
"""
import os
import socket
import sys

import numpy as np
import pytest

from hyperion_b200.multigpu import ShardedLucy, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_tile_the_id_range():
    for n in (0, 1, 7, 10000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            blocks = [shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert blocks[-1][0] + blocks[-1][1] == n
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard(10, 2, 2)


class _OracleEngine:
    """Adapter: the oracle behind the engine interface ShardedLucy drives."""

    def __init__(self, model, rank):
        import torch
        from oracle import oracle
        self.o = oracle.Oracle(model, rank=rank)
        self.torch = torch
        self.buf = None

    def lucy_begin(self):
        self.o.lucy_begin()

    def lucy_photons(self, first_id, n, iteration):
        self.o.lucy_photons(n)      # the reference stream is per rank, not per id

    def reduction_buffer(self):
        sums = self.o.get_energy_sum().ravel()
        self.buf = self.torch.from_numpy(np.concatenate([sums, [self.o.energy_current]]))
        return self.buf

    def lucy_finish(self):
        if self.buf is not None:
            b = self.buf.numpy()
            self.o.set_energy_sum(b[:-1].reshape((self.o.n_dust,) + tuple(self.o.shape)))
            self.o.energy_current = float(b[-1])
        return self.o.lucy_finish()


def _worker(rank, world, port, n_photons, n_iter, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import bitlevel_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_car.npz"))
    model = bitlevel_model(z, False, False)
    eng = _OracleEngine(model, rank)
    drv = ShardedLucy(eng, rank, world, all_reduce=lambda t: dist.all_reduce(t))
    for it in range(n_iter):
        drv.iteration(n_photons, it + 1)
    np.save(os.path.join(out_dir, "se_%d.npy" % rank), eng.o.get_specific_energy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_match_emulated_reference(golden_car, tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle
    from helpers import bitlevel_model
    oracle.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, n_iter = 20001, 2
    mp.spawn(_worker, args=(2, port, n, n_iter, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "se_0.npy")
    b = np.load(tmp_path / "se_1.npy")
    assert np.array_equal(a, b), "ranks disagree after the all-reduce"
    ref, _ = oracle.run_lucy_ranks(bitlevel_model(golden_car, False, False), n, n_ranks=2, n_iter=n_iter)
    assert np.array_equal(a, ref[-1])
