"""ctypes wrapper of the CPU oracle (oracle/hyperion_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under hyperion_b200/
imports this module.

``OracleRun`` drives the restated ``main`` -> ``do_lucy`` sequence
(src/main/main.f90:157-234) for one or several emulated MPI ranks (rank r is
seeded ``seed + r`` and the deposit grids are summed, as
src/mpi/mpi_routines.f90:266-314 does).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from hyperion_b200.capi import CApi, IterStats, _ptr
from hyperion_b200.flatmodel import FlatModel, apply_model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhyperion_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "hyperion_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libhyperion_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB


def load():
    global _lib
    if _lib is None:
        build()   # no-op unless the source is newer than the library
        _lib = C.CDLL(_LIB)
        _lib.orc_ctx_create.restype = C.c_int
        _lib.orc_finalize_setup.restype = C.c_int
        _lib.orc_lucy_photons.restype = C.c_int
        _lib.orc_run_lucy_iteration.restype = C.c_int
        _lib.orc_get_energy_current.restype = C.c_double
        _lib.orc_set_energy_sum.restype = C.c_int
        _lib.orc_test_random.restype = C.c_double
        _lib.orc_test_interp1d_loglog.restype = C.c_double
        _lib.orc_test_planck.restype = C.c_double
        _lib.orc_rng_draws.restype = C.c_uint64
        for n in ("orc_final_begin", "orc_final_photons", "orc_final_finish", "orc_raytracing_photons",
                  "orc_set_monochromatic", "orc_final_mono_photons"):
            getattr(_lib, n).restype = C.c_int
    return _lib


class Oracle(CApi):
    """One emulated reference process."""

    def __init__(self, model: FlatModel, rank=0):
        super().__init__(load(), "orc_")
        self.ctx = C.c_void_p()
        self.check(self.lib.orc_ctx_create(C.byref(self.ctx)))
        apply_model(self, self.ctx, model)
        self.check(self.lib.orc_finalize_setup(self.ctx, C.c_int32(rank)))
        self.n_dust = len(model.dust)
        self.shape = model.shape

    def __del__(self):
        if getattr(self, "ctx", None):
            self.lib.orc_ctx_destroy(self.ctx)
            self.ctx = None

    def _grid(self):
        return np.empty((self.n_dust,) + tuple(self.shape), dtype=np.float64)

    def run_lucy_iteration(self, n_photons):
        st = IterStats()
        self.check(self.lib.orc_run_lucy_iteration(self.ctx, C.c_int64(n_photons), C.byref(st)))
        return st

    def lucy_begin(self):
        self.check(self.lib.orc_lucy_begin(self.ctx))

    def lucy_photons(self, n):
        self.check(self.lib.orc_lucy_photons(self.ctx, C.c_int64(n)))

    def lucy_finish(self):
        st = IterStats()
        self.check(self.lib.orc_lucy_finish(self.ctx, C.byref(st)))
        return st

    # -- final / raytracing iterations (do_final, do_raytracing) -------------------------------
    def final_begin(self):
        self.check(self.lib.orc_final_begin(self.ctx))

    def final_photons(self, n, peeloff_scattering_only=False):
        self.check(self.lib.orc_final_photons(self.ctx, C.c_int64(n), C.c_int32(int(peeloff_scattering_only))))

    def final_mono_photons(self, inu, n_sources, n_total_sources, n_dust, n_total_dust, peeloff_scattering_only=False):
        self.final_mono_photons_raw(self.ctx, inu, 0, n_sources, n_total_sources, 0, n_dust, n_total_dust,
                                    peeloff_scattering_only)

    def final_finish(self):
        st = IterStats()
        self.check(self.lib.orc_final_finish(self.ctx, C.byref(st)))
        return st

    def raytracing_photons(self, n_sources, n_dust):
        st = IterStats()
        self.check(self.lib.orc_raytracing_photons(self.ctx, C.c_int64(n_sources), C.c_int64(n_dust), C.byref(st)))
        return st

    def sed(self, group, uncertainties=False):
        return self.get_sed(self.ctx, group, uncertainties)

    def image(self, group, uncertainties=False):
        return self.get_image(self.ctx, group, uncertainties)

    def get_specific_energy(self):
        out = self._grid()
        self.check(self.lib.orc_get_specific_energy(self.ctx, _ptr(out)))
        return out

    def get_energy_sum(self):
        out = self._grid()
        self.check(self.lib.orc_get_energy_sum(self.ctx, _ptr(out)))
        return out

    def get_energy_sum_spectrum(self):
        out = np.empty((self.n_nu_bins, self.n_dust) + tuple(self.shape), dtype=np.float64)
        self.check(self.lib.orc_get_energy_sum_spectrum(self.ctx, _ptr(out)))
        return out

    def set_energy_sum_spectrum(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.check(self.lib.orc_set_energy_sum_spectrum(self.ctx, _ptr(a)))

    def get_n_photons(self):
        out = np.empty(tuple(self.shape), dtype=np.int64)
        self.check(self.lib.orc_get_n_photons(self.ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_n_photons(self, a):
        a = np.ascontiguousarray(a, dtype=np.int64)
        self.check(self.lib.orc_set_n_photons(self.ctx, a.ctypes.data_as(C.c_void_p)))

    def put_specific_energy(self, a):
        """Overwrite the specific energy (test hook; the minimum specific energies stay)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.check(self.lib.orc_put_specific_energy(self.ctx, _ptr(a)))

    def set_pda_exact_limit(self, n):
        self.check(self.lib.orc_set_pda_exact_limit(self.ctx, C.c_int32(n)))

    def solve_pda(self):
        """solve_pda on the current specific energy and n_photons; returns the number of PDA cells."""
        n = self.lib.orc_solve_pda(self.ctx)
        if n < 0:
            self.check(1)
        return n

    def get_density(self):
        out = self._grid()
        self.check(self.lib.orc_get_density(self.ctx, _ptr(out)))
        return out

    def set_energy_sum(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.check(self.lib.orc_set_energy_sum(self.ctx, _ptr(a)))

    @property
    def energy_current(self):
        return self.lib.orc_get_energy_current(self.ctx)

    @energy_current.setter
    def energy_current(self, v):
        self.lib.orc_set_energy_current(self.ctx, C.c_double(v))


def run_lucy_ranks(model: FlatModel, n_photons, n_ranks=1, n_iter=1, first_rank=0):
    """Emulate ``mpirun -n n_ranks`` of the reference: equal photon split,
    rank r seeded seed+r, deposit grids and energy_current summed, every rank
    continues from the same scaled grid.  Returns (specific_energy per
    iteration, list of per-iteration stats dicts)."""
    ranks = [Oracle(model, rank=first_rank + r) for r in range(n_ranks)]
    split = [n_photons // n_ranks + (1 if r < n_photons % n_ranks else 0) for r in range(n_ranks)]
    out, stats = [], []
    with ThreadPoolExecutor(max_workers=n_ranks) as pool:
        for _ in range(n_iter):
            for o in ranks:
                o.lucy_begin()
            list(pool.map(lambda a: a[0].lucy_photons(a[1]), zip(ranks, split)))
            total = sum(o.get_energy_sum() for o in ranks) if n_ranks > 1 else None
            if n_ranks > 1 and model.spectrum_bin_edges is not None:
                total_nu = sum(o.get_energy_sum_spectrum() for o in ranks)
                for o in ranks:
                    o.set_energy_sum_spectrum(total_nu)
            e_cur = sum(o.energy_current for o in ranks)
            sts = []
            for o in ranks:
                if total is not None:
                    o.set_energy_sum(total)
                o.energy_current = e_cur
                sts.append(o.lucy_finish().as_dict())
            agg = dict(sts[0])
            for k in ("n_photons", "killed_geo", "killed_int", "n_crossings", "n_absorptions",
                      "n_scatterings", "n_escaped"):
                agg[k] = sum(s[k] for s in sts)
            agg["energy_emitted"] = e_cur
            stats.append(agg)
            out.append(ranks[0].get_specific_energy())
    return out, stats
