"""Process layout of a multi-GPU run: who am I, and how do N ranks get started.

The reference starts its MPI build as ``{mpi_command} -n N hyperion_<grid>_mpi [-f] in out``
(``scripts/hyperion:65-92``; ``Model.run(mpi=True, n_processes=N)`` passes ``-m N``,
``hyperion/model/model.py:1053-1074``) and every rank learns its place from ``mpi_comm_rank``
(``src/mpi/mpi_core.f90:35-40``).  Here a rank is one process driving one GPU, and it learns its
place from the environment its launcher left behind:

* ``torchrun``                      RANK / WORLD_SIZE / LOCAL_RANK
* Open MPI ``mpirun``               OMPI_COMM_WORLD_RANK / _SIZE / _LOCAL_RANK
* MPICH / Intel MPI / Hydra, PMI    PMI_RANK / PMI_SIZE (+ MPI_LOCALRANKID)
* Slurm ``srun``                    SLURM_PROCID / SLURM_NTASKS / SLURM_LOCALID
* ``bin/hyperion_mpirun -n N ...``  (``spawn`` below: for ``~/.hyperionrc [mpi] command``) sets the torchrun names
* none of these, ``HYPERION_B200_NGPU=N`` set: the serial name re-launches itself as N ranks (``spawn``)

so ``mpirun -n N hyperion_car_mpi in out`` is N coordinated GPU ranks writing ONE output file, as
the reference's is, instead of N uncoordinated copies.
"""
from __future__ import annotations

import os
import subprocess
import sys
import zlib

_SCHEMES = (
    ("RANK", "WORLD_SIZE", "LOCAL_RANK"),
    ("OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_RANK"),
    ("PMI_RANK", "PMI_SIZE", "MPI_LOCALRANKID"),
    ("PMIX_RANK", "PMIX_SIZE", "PMIX_LOCAL_RANK"),
    ("SLURM_PROCID", "SLURM_NTASKS", "SLURM_LOCALID"),
)


def layout(env=None):
    """(rank, world, local_rank, scheme) of this process; scheme is the variable the rank came from
    or None for a plain serial start."""
    env = os.environ if env is None else env
    for r, w, l in _SCHEMES:
        if r in env and w in env:
            try:
                rank, world = int(env[r]), int(env[w])
            except ValueError:
                continue
            if world < 1 or not (0 <= rank < world):
                raise ValueError("inconsistent process layout: %s=%s %s=%s" % (r, env[r], w, env[w]))
            try:
                local = int(env.get(l, rank))
            except ValueError:
                local = rank
            return rank, world, local, r
    return 0, 1, 0, None


def rendezvous(output_file, env=None):
    """MASTER_ADDR / MASTER_PORT for torch.distributed when the launcher did not set them (mpirun,
    srun): one node, and a port every rank derives from the output file name."""
    env = os.environ if env is None else env
    addr = env.get("MASTER_ADDR", "127.0.0.1")
    port = env.get("MASTER_PORT") or env.get("HYPERION_B200_MASTER_PORT")
    if not port:
        port = 20000 + zlib.crc32(os.path.abspath(output_file).encode()) % 20000
    return addr, int(port)


def wanted_gpus(env=None):
    """HYPERION_B200_NGPU: ranks a serial start should fan out to (1 = stay serial)."""
    env = os.environ if env is None else env
    try:
        return max(1, int(env.get("HYPERION_B200_NGPU", "1")))
    except ValueError:
        return 1


def spawn(n, argv, env=None):
    """Start ``argv`` n times on this node with the torchrun variables set, wait for all, return the
    first non-zero exit status (the other ranks are terminated when one fails, as ``mpirun`` does)."""
    base = dict(os.environ if env is None else env)
    base.pop("HYPERION_B200_NGPU", None)
    base.setdefault("MASTER_ADDR", "127.0.0.1")
    if "MASTER_PORT" not in base:
        base["MASTER_PORT"] = str(20000 + (os.getpid() * 7919) % 20000)
    procs = []
    for r in range(n):
        e = dict(base, RANK=str(r), WORLD_SIZE=str(n), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen(list(argv), env=e))
    status = 0
    alive = list(procs)
    while alive:
        for p in list(alive):
            try:
                rc = p.wait(timeout=0.2)
            except subprocess.TimeoutExpired:
                continue
            alive.remove(p)
            if rc != 0 and status == 0:
                status = rc
                for q in alive:
                    q.terminate()
    return status


def main(argv=None):
    """``hyperion_mpirun -n N program [args...]``: the part of the ``mpirun`` command line
    ``scripts/hyperion:89-92`` uses.  Other ``mpirun`` options are not understood."""
    argv = list(sys.argv[1:] if argv is None else argv)
    n = None
    while argv and argv[0].startswith("-"):
        if argv[0] in ("-n", "-np", "--np") and len(argv) >= 2:
            n = int(argv[1])
            argv = argv[2:]
        else:
            sys.stderr.write("hyperion_mpirun: unknown option %s (usage: hyperion_mpirun -n N program [args])\n" % argv[0])
            return 2
    if n is None or n < 1 or not argv:
        sys.stderr.write("usage: hyperion_mpirun -n N program [args]\n")
        return 2
    return spawn(n, argv)


if __name__ == "__main__":
    sys.exit(main())
