"""ctypes binding of ``include/hyperion_b200.h``.

This is the host side of the drop-in boundary: where the reference launches a
Fortran binary (``hyperion/model/model.py:1053-1080``), this module loads
``libhyperion_b200.so`` and drives the CUDA engine through its C ABI.  There is
no CPU fallback: if the shared library (or a CUDA device) is missing the load
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .flatmodel import FlatConf, FlatDust, FlatModel, FlatSource, apply_model

_dp = C.POINTER(C.c_double)


class DustTables(C.Structure):
    _fields_ = [
        ("version", C.c_int32), ("is_lte", C.c_int32), ("sublimation_mode", C.c_int32),
        ("sublimation_specific_energy", C.c_double),
        ("n_nu", C.c_int32), ("nu", _dp), ("albedo", _dp), ("chi", _dp),
        ("n_mu", C.c_int32), ("mu", _dp),
        ("P1", _dp), ("P2", _dp), ("P3", _dp), ("P4", _dp),
        ("n_e", C.c_int32), ("specific_energy", _dp),
        ("chi_planck", _dp), ("kappa_planck", _dp),
        ("chi_inv_planck", _dp), ("kappa_inv_planck", _dp),
        ("chi_rosseland", _dp), ("kappa_rosseland", _dp),
        ("n_emiss_nu", C.c_int32), ("emiss_nu", _dp),
        ("n_jnu", C.c_int32), ("emiss_jnu", _dp), ("jnu_var", _dp),
    ]


class Spot(C.Structure):
    _fields_ = [
        ("luminosity", C.c_double), ("longitude", C.c_double), ("latitude", C.c_double), ("radius", C.c_double),
        ("spectrum_type", C.c_int32), ("temperature", C.c_double),
        ("n_spec", C.c_int32), ("spec_nu", _dp), ("spec_fnu", _dp),
    ]


class Source(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("peeloff", C.c_int32), ("luminosity", C.c_double),
        ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
        ("radius", C.c_double), ("limb_darkening", C.c_int32),
        ("spectrum_type", C.c_int32), ("temperature", C.c_double),
        ("n_spec", C.c_int32), ("spec_nu", _dp), ("spec_fnu", _dp),
        ("box", C.c_double * 6), ("theta", C.c_double), ("phi", C.c_double),
        ("n_points", C.c_int64), ("points_xyz", _dp), ("points_lum", _dp),
        ("n_map", C.c_int64), ("map", _dp),
        ("n_spots", C.c_int32), ("spots", C.POINTER(Spot)),
    ]


class RunConf(C.Structure):
    _fields_ = [
        ("seed", C.c_int64), ("n_inter_max", C.c_int64), ("n_reabs_max", C.c_int64),
        ("kill_on_absorb", C.c_int32), ("kill_on_scatter", C.c_int32),
        ("sample_sources_evenly", C.c_int32), ("enforce_energy_range", C.c_int32),
        ("use_mrw", C.c_int32), ("mrw_gamma", C.c_double), ("n_mrw_max", C.c_int64),
        ("propagation_check_frequency", C.c_double),
        ("forced_first_interaction", C.c_int32), ("forced_first_interaction_algorithm", C.c_int32),
        ("baes16_xi", C.c_double),
        ("specific_energy_additional", C.c_int32),
        ("use_pda", C.c_int32), ("count_photons", C.c_int32),
    ]


class ImageConf(C.Structure):
    _fields_ = [
        ("n_view", C.c_int32), ("theta", _dp), ("phi", _dp),
        ("inside_observer", C.c_int32), ("ignore_optical_depth", C.c_int32),
        ("peeloff_x", C.c_double), ("peeloff_y", C.c_double), ("peeloff_z", C.c_double),
        ("d_min", C.c_double), ("d_max", C.c_double),
        ("compute_image", C.c_int32), ("n_x", C.c_int32), ("n_y", C.c_int32),
        ("x_min", C.c_double), ("x_max", C.c_double), ("y_min", C.c_double), ("y_max", C.c_double),
        ("compute_sed", C.c_int32), ("n_ap", C.c_int32), ("ap_min", C.c_double), ("ap_max", C.c_double),
        ("n_wav", C.c_int32), ("wav_min", C.c_double), ("wav_max", C.c_double),
        ("track_origin", C.c_int32), ("track_n_scat", C.c_int32),
        ("uncertainties", C.c_int32), ("compute_stokes", C.c_int32), ("io_bytes", C.c_int32),
        ("binned", C.c_int32), ("n_theta", C.c_int32), ("n_phi", C.c_int32),
        ("use_filters", C.c_int32), ("filt_n", C.POINTER(C.c_int32)), ("filt_nu", _dp), ("filt_tr", _dp),
        ("filt_nu0", _dp),
        ("inu_min", C.c_int32), ("inu_max", C.c_int32),
    ]


class IterStats(C.Structure):
    _fields_ = [
        ("energy_emitted", C.c_double), ("n_photons", C.c_int64),
        ("killed_geo", C.c_int64), ("killed_int", C.c_int64),
        ("n_crossings", C.c_int64), ("n_absorptions", C.c_int64),
        ("n_scatterings", C.c_int64), ("n_escaped", C.c_int64),
        ("kernel_ms", C.c_double), ("epilogue_ms", C.c_double),
        ("flight_ms", C.c_double), ("n_rounds", C.c_int64), ("n_launches", C.c_int64), ("n_peel_crossings", C.c_int64), ("n_peeloffs", C.c_int64),
        ("n_peel_cached", C.c_int64), ("n_wave_rounds", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _ptr(a):
    return a.ctypes.data_as(_dp)


class HyperionError(RuntimeError):
    pass


class CApi:
    """Thin object wrapper over a shared library exporting ``<prefix>*``
    functions with the signatures of ``include/hyperion_b200.h``."""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        self._keep = []
        f = self._fn
        f("last_error").restype = C.c_char_p
        for name in ("set_grid_cartesian", "set_grid_spherical", "set_grid_cylindrical", "set_grid_octree", "set_grid_amr", "set_grid_voronoi", "add_dust", "add_source", "set_run_conf", "set_density",
                     "set_specific_energy", "lucy_begin", "lucy_finish", "get_specific_energy",
                     "get_density", "get_energy_sum", "add_peeled_group", "final_begin", "final_photons",
                     "final_finish", "raytracing_photons", "image_shape", "get_sed", "get_image"):
            f(name).restype = C.c_int

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def check(self, rc):
        if rc != 0:
            msg = self._fn("last_error")()
            raise HyperionError((msg or b"unknown error").decode("utf-8", "replace"))

    # -- setters -----------------------------------------------------------
    def set_grid_cartesian(self, ctx, n1, n2, n3, w1, w2, w3):
        self.check(self._fn("set_grid_cartesian")(ctx, C.c_int32(n1), C.c_int32(n2), C.c_int32(n3),
                                                  _ptr(w1), _ptr(w2), _ptr(w3)))

    def set_grid_spherical(self, ctx, n1, n2, n3, w1, w2, w3):
        self.check(self._fn("set_grid_spherical")(ctx, C.c_int32(n1), C.c_int32(n2), C.c_int32(n3),
                                                  _ptr(w1), _ptr(w2), _ptr(w3)))

    def set_grid_cylindrical(self, ctx, n1, n2, n3, w1, w2, w3):
        self.check(self._fn("set_grid_cylindrical")(ctx, C.c_int32(n1), C.c_int32(n2), C.c_int32(n3),
                                                    _ptr(w1), _ptr(w2), _ptr(w3)))

    def set_grid_octree(self, ctx, refined, center, half):
        refined = np.ascontiguousarray(refined, dtype=np.int32)
        self.check(self._fn("set_grid_octree")(ctx, C.c_int32(len(refined)),
                                               refined.ctypes.data_as(C.POINTER(C.c_int32)),
                                               *[C.c_double(float(v)) for v in tuple(center) + tuple(half)]))

    def set_grid_voronoi(self, ctx, v):
        i32 = C.POINTER(C.c_int32)
        self.check(self._fn("set_grid_voronoi")(ctx, C.c_int32(len(v["volume"])), _ptr(v["coordinates"]), _ptr(v["bb_min"]),
                                                _ptr(v["bb_max"]), _ptr(v["volume"]), v["sparse_idx"].ctypes.data_as(i32),
                                                v["sparse_neighs"].ctypes.data_as(i32), _ptr(v["box"])))

    def set_specific_energy_spectrum_bins(self, ctx, edges):
        edges = np.ascontiguousarray(edges, dtype=np.float64)
        f = self._fn("set_specific_energy_spectrum_bins")
        f.restype = C.c_int
        self.check(f(ctx, C.c_int32(len(edges)), _ptr(edges)))
        self.n_nu_bins = len(edges) - 1

    def get_specific_energy_spectrum(self):
        """[n_bins, n_dust, cells...] as the reference writes /specific_energy_spectrum."""
        out = np.empty((self.n_nu_bins, self.n_dust) + tuple(self.shape), dtype=np.float64)
        f = self._fn("get_specific_energy_spectrum")
        f.restype = C.c_int
        self.check(f(self.ctx, _ptr(out)))
        return out

    def set_grid_amr(self, ctx, levels):
        n_grids = np.array([len(lev) for lev in levels], dtype=np.int32)
        dims = np.array([g[:3] for lev in levels for g in lev], dtype=np.int32).ravel()
        bounds = np.array([g[3:9] for lev in levels for g in lev], dtype=np.float64).ravel()
        i32 = C.POINTER(C.c_int32)
        self.check(self._fn("set_grid_amr")(ctx, C.c_int32(len(levels)), n_grids.ctypes.data_as(i32),
                                            dims.ctypes.data_as(i32), _ptr(bounds)))

    def add_dust(self, ctx, d: FlatDust):
        t = DustTables()
        t.version, t.is_lte, t.sublimation_mode = d.version, int(d.is_lte), d.sublimation_mode
        t.sublimation_specific_energy = d.sublimation_specific_energy
        t.n_nu, t.n_mu, t.n_e = len(d.nu), len(d.mu), len(d.specific_energy)
        t.n_emiss_nu, t.n_jnu = len(d.emiss_nu), len(d.jnu_var)
        for k in ("nu", "albedo", "chi", "mu", "P1", "P2", "P3", "P4", "specific_energy", "chi_planck",
                  "kappa_planck", "chi_inv_planck", "kappa_inv_planck", "chi_rosseland", "kappa_rosseland",
                  "emiss_nu", "emiss_jnu", "jnu_var"):
            setattr(t, k, _ptr(getattr(d, k)))
        self.check(self._fn("add_dust")(ctx, C.byref(t)))

    def add_source(self, ctx, s: FlatSource):
        t = Source()
        t.type, t.peeloff, t.luminosity = s.type, int(s.peeloff), s.luminosity
        t.x, t.y, t.z = [float(v) for v in s.position]
        t.radius, t.limb_darkening = s.radius, int(s.limb_darkening)
        for k in range(6):
            t.box[k] = float(s.bounds[k])
        t.theta, t.phi = float(s.direction[0]), float(s.direction[1])
        keep_pts = None
        if s.points is not None:
            xyz = np.ascontiguousarray(s.points, dtype=np.float64).reshape(-1, 3)
            lum = np.ascontiguousarray(s.points_luminosity, dtype=np.float64)
            if len(lum) != len(xyz):
                raise HyperionError("point collection: positions and luminosities differ in length")
            keep_pts = (xyz, lum)
            t.n_points, t.points_xyz, t.points_lum = len(lum), _ptr(xyz), _ptr(lum)
        keep_spots = []
        if s.spots:
            arr = (Spot * len(s.spots))()
            for q, d in zip(arr, s.spots):
                q.luminosity, q.longitude, q.latitude, q.radius = (float(d[k]) for k in ("luminosity", "longitude", "latitude", "radius"))
                if d.get("temperature") is not None:
                    q.spectrum_type, q.temperature = 2, float(d["temperature"])
                else:
                    nu = np.ascontiguousarray(d["spectrum_nu"], dtype=np.float64)
                    fnu = np.ascontiguousarray(d["spectrum_fnu"], dtype=np.float64)
                    keep_spots.append((nu, fnu))
                    q.spectrum_type, q.n_spec, q.spec_nu, q.spec_fnu = 1, len(nu), _ptr(nu), _ptr(fnu)
            keep_spots.append(arr)
            t.n_spots, t.spots = len(s.spots), arr
        keep_map = None
        if s.map is not None:
            keep_map = np.ascontiguousarray(s.map, dtype=np.float64).reshape(-1)
            t.n_map, t.map = len(keep_map), _ptr(keep_map)
        keep = None
        if s.lte:
            t.spectrum_type = 3
        elif s.temperature is not None:
            t.spectrum_type, t.temperature = 2, float(s.temperature)
        else:
            nu = np.ascontiguousarray(s.spectrum_nu, dtype=np.float64)
            fnu = np.ascontiguousarray(s.spectrum_fnu, dtype=np.float64)
            keep = (nu, fnu)
            t.spectrum_type, t.n_spec, t.spec_nu, t.spec_fnu = 1, len(nu), _ptr(nu), _ptr(fnu)
        self.check(self._fn("add_source")(ctx, C.byref(t)))
        del keep, keep_pts, keep_map, keep_spots

    def set_run_conf(self, ctx, c: FlatConf):
        t = RunConf()
        t.seed, t.n_inter_max, t.n_reabs_max = c.seed, c.n_inter_max, c.n_reabs_max
        t.kill_on_absorb, t.kill_on_scatter = int(c.kill_on_absorb), int(c.kill_on_scatter)
        t.sample_sources_evenly, t.enforce_energy_range = int(c.sample_sources_evenly), int(c.enforce_energy_range)
        t.use_mrw, t.mrw_gamma, t.n_mrw_max = int(c.use_mrw), c.mrw_gamma, c.n_mrw_max
        t.propagation_check_frequency = c.propagation_check_frequency
        t.forced_first_interaction = int(c.forced_first_interaction)
        t.forced_first_interaction_algorithm = {"wr99": 1, "baes16": 2}[c.forced_first_interaction_algorithm]
        t.baes16_xi = c.baes16_xi
        t.specific_energy_additional = int(c.specific_energy_additional)
        t.use_pda, t.count_photons = int(c.use_pda), int(c.count_photons)
        self.check(self._fn("set_run_conf")(ctx, C.byref(t)))

    def add_peeled_group(self, ctx, g):
        """g: :class:`hyperion_b200.flatmodel.FlatPeeledGroup`"""
        t = ImageConf()
        if g.binned:
            t.binned, t.n_theta, t.n_phi = 1, int(g.n_theta), int(g.n_phi)
            t.n_view = int(g.n_theta) * int(g.n_phi)
        else:
            theta = np.ascontiguousarray(g.theta, dtype=np.float64)
            phi = np.ascontiguousarray(g.phi, dtype=np.float64)
            t.n_view, t.theta, t.phi = len(theta), _ptr(theta), _ptr(phi)
        t.inside_observer, t.ignore_optical_depth = int(g.inside_observer), int(g.ignore_optical_depth)
        t.peeloff_x, t.peeloff_y, t.peeloff_z = [float(v) for v in g.peeloff_origin]
        t.d_min, t.d_max = g.d_min, g.d_max
        t.compute_image = int(g.image is not None)
        if g.image is not None:
            t.n_x, t.n_y, t.x_min, t.x_max, t.y_min, t.y_max = g.image
        t.compute_sed = int(g.sed is not None)
        if g.sed is not None:
            t.n_ap, t.ap_min, t.ap_max = g.sed
        t.n_wav, t.wav_min, t.wav_max = g.wavelengths
        keep_f = None
        if g.filters:
            fn = np.array([len(f[0]) for f in g.filters], dtype=np.int32)
            fnu = np.ascontiguousarray(np.concatenate([np.asarray(f[0], dtype=np.float64) for f in g.filters]))
            ftr = np.ascontiguousarray(np.concatenate([np.asarray(f[1], dtype=np.float64) for f in g.filters]))
            fnu0 = np.array([float(f[2]) for f in g.filters], dtype=np.float64)
            keep_f = (fn, fnu, ftr, fnu0)
            t.use_filters, t.n_wav = 1, len(g.filters)
            t.filt_n = fn.ctypes.data_as(C.POINTER(C.c_int32))
            t.filt_nu, t.filt_tr, t.filt_nu0 = _ptr(fnu), _ptr(ftr), _ptr(fnu0)
        t.track_origin = {"no": 0, "basic": 1, "yes": 1, "detailed": 2, "scatterings": 3}[g.track_origin]
        t.track_n_scat = g.track_n_scat
        t.uncertainties, t.compute_stokes, t.io_bytes = int(g.uncertainties), int(g.stokes), g.io_bytes
        t.inu_min, t.inu_max = int(getattr(g, "inu_min", 0)), int(getattr(g, "inu_max", 0))
        if t.inu_min > 0:
            t.n_wav = t.inu_max - t.inu_min + 1
        self.check(self._fn("add_peeled_group")(ctx, C.byref(t)))
        del keep_f

    def image_shape(self, ctx, group, which):
        dims = (C.c_int64 * 6)()
        nd = C.c_int32()
        self.check(self._fn("image_shape")(ctx, C.c_int32(group), C.c_int32(which), dims, C.byref(nd)))
        return tuple(dims[i] for i in range(nd.value))

    def _get_cube(self, ctx, group, which, uncertainties=False):
        shape = self.image_shape(ctx, group, which)
        out = np.zeros(shape, dtype=np.float64)
        unc = np.zeros(shape, dtype=np.float64) if uncertainties else None
        fn = self._fn("get_sed" if which == 0 else "get_image")
        self.check(fn(ctx, C.c_int32(group), _ptr(out), None if unc is None else _ptr(unc)))
        return (out, unc) if uncertainties else out

    def get_sed(self, ctx, group, uncertainties=False):
        return self._get_cube(ctx, group, 0, uncertainties)

    def get_image(self, ctx, group, uncertainties=False):
        return self._get_cube(ctx, group, 1, uncertainties)

    def set_monochromatic(self, ctx, frequencies, energy_threshold):
        nu = np.ascontiguousarray(frequencies, dtype=np.float64)
        self.check(self._fn("set_monochromatic")(ctx, C.c_int32(len(nu)), _ptr(nu), C.c_double(energy_threshold)))

    def final_mono_photons_raw(self, ctx, inu, first_s, n_s, n_s_tot, first_d, n_d, n_d_tot, scattering_only):
        self.check(self._fn("final_mono_photons")(ctx, C.c_int32(inu), C.c_int64(first_s), C.c_int64(n_s), C.c_int64(n_s_tot),
                                                  C.c_int64(first_d), C.c_int64(n_d), C.c_int64(n_d_tot),
                                                  C.c_int32(int(scattering_only))))

    def set_density(self, ctx, n_dust, density):
        density = np.ascontiguousarray(density, dtype=np.float64)
        self.check(self._fn("set_density")(ctx, C.c_int32(n_dust), _ptr(density)))

    def set_specific_energy(self, ctx, se, min_e):
        se_p = None if se is None else _ptr(np.ascontiguousarray(se, dtype=np.float64))
        me_p = None if min_e is None else _ptr(np.ascontiguousarray(min_e, dtype=np.float64))
        self.check(self._fn("set_specific_energy")(ctx, se_p, me_p))


# ---------------------------------------------------------------------------
# the product library
# ---------------------------------------------------------------------------
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libhyperion_b200.so")
_lib = None


def load_library(path=None):
    """Load ``libhyperion_b200.so``; raises if it has not been built
    (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None or path is not None:
        p = path or os.environ.get("HYPERION_B200_LIB") or _LIB_PATH
        if not os.path.exists(p):
            raise HyperionError("%s not found: build it with __graft_entry__.build(); "
                                "there is no CPU fallback" % p)
        _lib = C.CDLL(p)
    return _lib


class Engine(CApi):
    """One GPU's photon-propagation context (``hyp_ctx``)."""

    def __init__(self, device_id=0, lib=None):
        super().__init__(lib or load_library(), "hyp_")
        L = self.lib
        L.hyp_ctx_create.restype = C.c_int
        L.hyp_ctx_destroy.restype = None
        L.hyp_stream.restype = C.c_void_p
        for n in ("hyp_finalize_setup", "hyp_lucy_photons", "hyp_lucy_device_buffers",
                  "hyp_run_lucy_iteration", "hyp_image_device_buffers"):
            getattr(L, n).restype = C.c_int
        self.ctx = C.c_void_p()
        self.check(L.hyp_ctx_create(C.c_int(device_id), C.byref(self.ctx)))
        self.device_id = device_id
        self.n_dust = 0
        self.n_cells = 0
        self.shape = None

    def close(self):
        if self.ctx:
            self.lib.hyp_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_model(self, model: FlatModel):
        apply_model(self, self.ctx, model)
        self.n_dust = len(model.dust)
        self.shape = model.shape
        self.n_cells = model.n_cells
        self.check(self.lib.hyp_finalize_setup(self.ctx))

    def update_density(self, density):
        self.set_density(self.ctx, self.n_dust, density)

    def update_density_device(self, device_ptr):
        """Replace the densities from a device buffer in the .rtin layout [n_dust][n3][n2][n1] (fp64) that
        lives on this engine's GPU, e.g. one filled by an NCCL broadcast."""
        self.check(self.lib.hyp_set_density(self.ctx, C.c_int32(self.n_dust), C.c_void_p(int(device_ptr))))

    def lucy_begin(self):
        self.check(self.lib.hyp_lucy_begin(self.ctx))

    def lucy_photons(self, first_id, n, iteration):
        self.check(self.lib.hyp_lucy_photons(self.ctx, C.c_int64(first_id), C.c_int64(n), C.c_int64(iteration)))

    def lucy_device_buffers(self):
        p = C.c_void_p()
        n = C.c_int64()
        self.check(self.lib.hyp_lucy_device_buffers(self.ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def reduction_buffer(self):
        """The iteration's deposit grid + scalars as a torch CUDA tensor aliasing the engine's
        device buffer (for the NCCL all-reduce of :mod:`hyperion_b200.multigpu`).  The engine's
        stream has been synchronised when this returns."""
        import torch
        ptr, n = self.lucy_device_buffers()

        class _Buf:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Buf(), device=torch.device("cuda", self.device_id))

    def lucy_finish(self):
        st = IterStats()
        self.check(self.lib.hyp_lucy_finish(self.ctx, C.byref(st)))
        return st

    def run_lucy_iteration(self, n_photons, iteration=1):
        st = IterStats()
        self.check(self.lib.hyp_run_lucy_iteration(self.ctx, C.c_int64(n_photons), C.c_int64(iteration),
                                                   C.byref(st)))
        return st

    # -- final (imaging) and raytracing iterations: do_final, do_raytracing ---------------------
    def final_begin(self):
        self.check(self.lib.hyp_final_begin(self.ctx))

    def final_photons(self, first_id, n, peeloff_scattering_only=False):
        self.check(self.lib.hyp_final_photons(self.ctx, C.c_int64(first_id), C.c_int64(n),
                                              C.c_int32(int(peeloff_scattering_only))))

    def final_mono_photons(self, inu, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust, n_total_dust,
                           peeloff_scattering_only=False):
        """do_final_mono for frequency ``inu`` (1-based): this rank's share of the source and thermal packets."""
        self.final_mono_photons_raw(self.ctx, inu, first_source_id, n_sources, n_total_sources, first_dust_id, n_dust,
                                    n_total_dust, peeloff_scattering_only)

    def final_finish(self):
        st = IterStats()
        self.check(self.lib.hyp_final_finish(self.ctx, C.byref(st)))
        return st

    def raytracing_photons(self, n_sources, n_dust, first_source_id=0, n_total_sources=None,
                           first_dust_id=0, n_total_dust=None):
        """Packets [first_*_id, first_*_id + n_*) of a raytracing iteration of n_total_* packets
        (defaults: this call is the whole iteration)."""
        st = IterStats()
        self.check(self.lib.hyp_raytracing_photons(
            self.ctx, C.c_int64(first_source_id), C.c_int64(n_sources),
            C.c_int64(n_sources if n_total_sources is None else n_total_sources),
            C.c_int64(first_dust_id), C.c_int64(n_dust),
            C.c_int64(n_dust if n_total_dust is None else n_total_dust), C.byref(st)))
        return st

    def image_device_buffers(self):
        p = C.c_void_p()
        n = C.c_int64()
        self.check(self.lib.hyp_image_device_buffers(self.ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def image_buffer(self):
        """All image / SED accumulators + scalars as a torch CUDA tensor aliasing device memory
        (for the end-of-run reduction, mp_collect_images src/mpi/mpi_routines.f90:363-471)."""
        import torch
        ptr, n = self.image_device_buffers()

        class _Buf:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Buf(), device=torch.device("cuda", self.device_id))

    def sed(self, group, uncertainties=False):
        return self.get_sed(self.ctx, group, uncertainties)

    def image(self, group, uncertainties=False):
        return self.get_image(self.ctx, group, uncertainties)

    def _get(self, fn, out=None):
        if out is None:
            out = np.empty((self.n_dust,) + tuple(self.shape), dtype=np.float64)
        self.check(fn(self.ctx, _ptr(out)))
        return out

    def get_specific_energy(self, out=None):
        return self._get(self.lib.hyp_get_specific_energy, out)

    def get_density(self, out=None):
        return self._get(self.lib.hyp_get_density, out)

    def set_specific_energy_array(self, se):
        """Replace the specific energy [n_dust, n3, n2, n1] of a loaded model (hyp_set_specific_energy)."""
        self.set_specific_energy(self.ctx, se, None)

    def solve_pda(self, n_photons=None):
        """solve_pda (src/grid/grid_pda_3d.f90:105-169) on the current specific energy with the given packet counts
        [n3, n2, n1] (default: those of the last Lucy iteration); returns the number of PDA cells."""
        n = C.c_int64(0)
        p = None
        if n_photons is not None:
            a = np.ascontiguousarray(n_photons, dtype=np.int64)
            p = a.ctypes.data_as(C.c_void_p)
        self.check(self.lib.hyp_solve_pda(self.ctx, p, C.byref(n)))
        return n.value

    def get_n_photons(self):
        """Packets that visited each cell in the last Lucy iteration (``n_photons``, grid_physics_3d.f90:38)."""
        out = np.empty(tuple(self.shape), dtype=np.int64)
        self.check(self.lib.hyp_get_n_photons(self.ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_energy_sum(self, out=None):
        return self._get(self.lib.hyp_get_energy_sum, out)

    @property
    def stream(self):
        return self.lib.hyp_stream(self.ctx)
