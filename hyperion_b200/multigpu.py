"""Photon-packet sharding across GPUs: one process per GPU, one all-reduce per Lucy iteration.

Replaces the reference's MPI layer for the photon loop (``src/mpi/mpi_routines.f90``):

* ``mp_n_photons`` (``:62-264``) hands out chunks of packets dynamically from a master rank;
  here rank ``r`` of ``W`` takes the contiguous id block ``shard(n, r, W)`` -- packets are keyed
  by id (counter RNG), so the union of the shards is exactly the single-GPU run;
* ``mp_collect_physical_arrays`` + ``mp_sync`` + ``mp_broadcast_specific_energy``
  (``:272-361``: reduce to rank 0, scale there, broadcast back) become ONE all-reduce of the
  deposit grid with the iteration's scalars appended (``hyp_lucy_device_buffers``); every rank
  then runs the same scale/clamp epilogue locally.

The collective is injected (``all_reduce(buffer)``) so the same driver runs over NCCL on device
buffers (``bench.py``, ``runner.py``) and over gloo on host buffers in the CPU tests.
"""
from __future__ import annotations


def shard(n_photons: int, rank: int, world: int):
    """Contiguous block of packet ids for ``rank``: returns (first_id, count).

    Blocks differ in size by at most one packet and tile [0, n_photons) exactly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    base, extra = divmod(int(n_photons), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


class ShardedLucy:
    """Drives one rank's share of a Lucy iteration (``do_lucy``, ``src/main/iter_lucy.f90:66-237``).

    ``engine`` exposes ``lucy_begin() / lucy_photons(first_id, n, iteration) /
    reduction_buffer() / lucy_finish()``; ``all_reduce`` sums the buffer returned by
    ``reduction_buffer()`` in place over all ranks (``None`` for a single rank).
    """

    def __init__(self, engine, rank=0, world=1, all_reduce=None):
        if world > 1 and all_reduce is None:
            raise ValueError("a collective is required when world > 1")
        self.engine = engine
        self.rank = rank
        self.world = world
        self.all_reduce = all_reduce

    def iteration(self, n_photons_total: int, iteration: int, id_offset: int = 0):
        """One Lucy iteration over packets [id_offset, id_offset + n_photons_total); returns the
        engine's statistics (global counters when world > 1: they are reduced with the grid)."""
        first, count = shard(n_photons_total, self.rank, self.world)
        eng = self.engine
        eng.lucy_begin()
        eng.lucy_photons(id_offset + first, count, iteration)
        if self.world > 1:
            self.all_reduce(eng.reduction_buffer())
        return eng.lucy_finish()
