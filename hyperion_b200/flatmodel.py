"""Neutral in-memory model: what the .rtin file holds for the photon path.

The reference keeps this state in Fortran module globals filled by
``setup_initial`` (``src/main/setup_rt.f90:27-304``).  Here it is a handful of
numpy arrays that both the CUDA engine (through the C ABI) and the test oracle
consume, so that kernels can be exercised without any HDF5 file.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


@dataclass
class FlatDust:
    """One dust type: the columns of a Hyperion dust file
    (``src/dust/dust_type_4elem.f90:94-277``)."""
    nu: np.ndarray
    albedo: np.ndarray
    chi: np.ndarray
    mu: np.ndarray
    P1: np.ndarray  # [n_nu, n_mu]
    P2: np.ndarray
    P3: np.ndarray
    P4: np.ndarray
    specific_energy: np.ndarray
    chi_planck: np.ndarray
    kappa_planck: np.ndarray
    chi_inv_planck: np.ndarray
    kappa_inv_planck: np.ndarray
    chi_rosseland: np.ndarray
    kappa_rosseland: np.ndarray
    emiss_nu: np.ndarray
    emiss_jnu: np.ndarray  # [n_emiss_nu, n_jnu]
    jnu_var: np.ndarray
    version: int = 2
    is_lte: bool = True
    sublimation_mode: int = 0
    sublimation_specific_energy: float = 0.0

    def __post_init__(self):
        for k in ("nu", "albedo", "chi", "mu", "P1", "P2", "P3", "P4", "specific_energy", "chi_planck",
                  "kappa_planck", "chi_inv_planck", "kappa_inv_planck", "chi_rosseland", "kappa_rosseland",
                  "emiss_nu", "emiss_jnu", "jnu_var"):
            setattr(self, k, _f8(getattr(self, k)))
        n_nu, n_mu = len(self.nu), len(self.mu)
        for k in ("P1", "P2", "P3", "P4"):
            if getattr(self, k).shape != (n_nu, n_mu):
                raise ValueError("%s should have shape (n_nu, n_mu)" % k)
        if self.emiss_jnu.shape != (len(self.emiss_nu), len(self.jnu_var)):
            raise ValueError("emiss_jnu should have shape (n_emiss_nu, n_jnu)")

    @classmethod
    def from_hdf5_group(cls, g):
        """Read a dust group (``Dust/dust_%03i`` or the root of a dust file)."""
        sub = {b"no": 0, b"fast": 1, b"slow": 2, b"cap": 3}
        attrs = g.attrs
        version = int(np.asarray(attrs["version"]).ravel()[0])
        op = g["optical_properties"][...]
        mo = g["mean_opacities"][...]
        em = g["emissivities"][...]
        ev = g["emissivity_variable"][...]
        mu = g["scattering_angles"][...]["mu"]
        if version == 1:
            # dust_type_4elem.f90:232-238: version-1 files hold the inverse Planck
            # means in the *_rosseland columns
            chi_inv, kap_inv = mo["chi_rosseland"], mo["kappa_rosseland"]
        else:
            chi_inv, kap_inv = mo["chi_inv_planck"], mo["kappa_inv_planck"]
        mode = bytes(attrs["sublimation_mode"]).strip()
        return cls(nu=op["nu"], albedo=op["albedo"], chi=op["chi"], mu=mu,
                   P1=op["P1"], P2=op["P2"], P3=op["P3"], P4=op["P4"],
                   specific_energy=mo["specific_energy"], chi_planck=mo["chi_planck"],
                   kappa_planck=mo["kappa_planck"], chi_inv_planck=chi_inv, kappa_inv_planck=kap_inv,
                   chi_rosseland=mo["chi_rosseland"], kappa_rosseland=mo["kappa_rosseland"],
                   emiss_nu=em["nu"], emiss_jnu=em["jnu"], jnu_var=ev["specific_energy"],
                   version=version, is_lte=bytes(attrs["lte"]).strip().lower() in (b"yes", b"y", b"true"),
                   sublimation_mode=sub[mode],
                   sublimation_specific_energy=float(attrs.get("sublimation_specific_energy", 0.0)))

    def to_npz_dict(self, prefix):
        d = {}
        for k, v in self.__dict__.items():
            d[prefix + k] = np.asarray(v)
        return d

    @classmethod
    def from_npz_dict(cls, z, prefix):
        kw = {}
        for k in cls.__dataclass_fields__:
            v = z[prefix + k]
            kw[k] = v if v.ndim else v.item()
        return cls(**kw)


@dataclass
class FlatSource:
    """One source (``src/sources/source_type.f90:102-282``)."""
    type: int = 1              # the reference's numbering: 1 point, 2 sphere, 4 map, 5 extern_sph, 6 extern_box,
                               # 7 plane_parallel, 8 point_collection
    luminosity: float = 0.0
    position: tuple = (0.0, 0.0, 0.0)
    temperature: Optional[float] = None   # blackbody
    spectrum_nu: Optional[np.ndarray] = None
    spectrum_fnu: Optional[np.ndarray] = None
    radius: float = 0.0
    limb_darkening: bool = False
    peeloff: bool = True
    bounds: tuple = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)   # extern_box: xmin, xmax, ymin, ymax, zmin, zmax
    direction: tuple = (0.0, 0.0)                    # plane_parallel: (theta, phi) in degrees
    points: Optional[np.ndarray] = None              # point_collection: [n, 3] positions
    points_luminosity: Optional[np.ndarray] = None   # point_collection: [n]
    map: Optional[np.ndarray] = None                 # map (type 4): luminosity per cell, shaped like one density array
    lte: bool = False                                # map only: spectrum = emissivity of the dust in the emitting cell
    # sphere only: spots, each a dict(luminosity, longitude, latitude, radius [degrees], temperature | spectrum_nu + spectrum_fnu)
    spots: Optional[list] = None


@dataclass
class FlatConf:
    """Run configuration (``hyperion/conf/conf_files.py:48-73``)."""
    seed: int = -124902
    n_inter_max: int = 1000000
    n_reabs_max: int = 1000000
    kill_on_absorb: bool = False
    kill_on_scatter: bool = False
    sample_sources_evenly: bool = False
    enforce_energy_range: bool = True
    use_mrw: bool = False
    mrw_gamma: float = 1.0
    n_mrw_max: int = 1000
    propagation_check_frequency: float = 1.e-3
    specific_energy_additional: bool = False   # specific_energy_type = 'additional'
    use_pda: bool = False                      # 'pda' (setup_rt.f90:75; src/grid/grid_pda_3d.f90)
    count_photons: bool = False                # keep n_photons without the PDA (output_n_photons /= 'none')
    n_initial_iter: int = 5
    n_initial_photons: int = 0
    forced_first_interaction: bool = True
    forced_first_interaction_algorithm: str = "wr99"
    baes16_xi: float = 0.5


@dataclass
class FlatPeeledGroup:
    """One ``Output/Peeled/group_%05i`` (``hyperion/conf/conf_files.py`` PeeledImageConf;
    ``src/images/images_peeled.f90:272-382``, ``src/images/image_type.f90:153-335``)."""
    theta: np.ndarray = None            # degrees
    phi: np.ndarray = None
    wavelengths: tuple = (1, 1.0, 1000.0)          # (n_wav, wav_min, wav_max) microns
    image: Optional[tuple] = None       # (n_x, n_y, x_min, x_max, y_min, y_max)
    sed: Optional[tuple] = None         # (n_ap, ap_min, ap_max)
    track_origin: str = "no"
    track_n_scat: int = 0
    uncertainties: bool = False
    stokes: bool = True
    io_bytes: int = 8
    inside_observer: bool = False
    ignore_optical_depth: bool = False
    peeloff_origin: tuple = (0.0, 0.0, 0.0)
    d_min: float = -np.inf
    d_max: float = np.inf
    # a binned group (``Output/Binned/group_00001``, BinnedImageConf, ``src/images/images_binned.f90``):
    # theta / phi unused, escaping packets are binned into n_theta x n_phi direction bins
    binned: bool = False
    n_theta: int = 0
    n_phi: int = 0
    # filter convolution (``use_filters``, ``hyperion/conf/conf_files.py:862-885``): a list of
    # (nu, normalised transmission, nu0) triples replaces the wavelength grid
    filters: Optional[list] = None
    # monochromatic mode (``image_type.f90:243-258``): the channels are frequencies inu_min .. inu_max (1-based)
    # of ``FlatModel.frequencies``; 0 = not monochromatic
    inu_min: int = 0
    inu_max: int = 0


@dataclass
class FlatModel:
    w1: np.ndarray
    w2: np.ndarray
    w3: np.ndarray
    density: np.ndarray                 # [n_dust, n3, n2, n1]
    dust: List[FlatDust]
    sources: List[FlatSource]
    conf: FlatConf = field(default_factory=FlatConf)
    specific_energy: Optional[np.ndarray] = None
    minimum_specific_energy: Optional[np.ndarray] = None
    peeled: List["FlatPeeledGroup"] = field(default_factory=list)
    binned: Optional["FlatPeeledGroup"] = None   # image group index len(peeled) on the engine / oracle
    # monochromatic mode (``set_monochromatic``; src/main/setup_rt.f90:49-56,220-222): frequencies in Hz, or None
    frequencies: Optional[np.ndarray] = None
    monochromatic_energy_threshold: float = 1.e-10
    grid_type: str = "car"              # "car" (x, y, z walls), "sph" (r, theta, phi), "cyl" (w, z, phi), "oct"
    # octree (grid_type "oct", hyperion/grid/octree_grid.py): depth-first refinement flags, centre and
    # HALF-widths of the root cell; density is then [n_dust, n_nodes] and w1/w2/w3 are unused
    refined: Optional[np.ndarray] = None
    oct_center: tuple = (0.0, 0.0, 0.0)
    oct_half: tuple = (1.0, 1.0, 1.0)
    # AMR (grid_type "amr", hyperion/grid/amr_grid.py): amr_levels[level] = list of grids, each
    # (n1, n2, n3, xmin, xmax, ymin, ymax, zmin, zmax); density is [n_dust, n_cells] with the cells of
    # all grids concatenated level-major, grid-major, x fastest (src/core/type_cell_id_amr.f90:115-133)
    amr_levels: Optional[list] = None
    # Voronoi mesh (grid_type "vor", hyperion/grid/voronoi_grid.py:417-478): a dict with the columns of the 'cells'
    # table -- "coordinates", "bb_min", "bb_max" [n, 3], "volume" [n] -- the neighbour lists "sparse_neighs" /
    # "sparse_idx" in the file's numbering (>= 0: cell, -1 .. -6: the walls xmin, xmax, ymin, ymax, zmin, zmax of
    # the box) and "box" = (xmin, xmax, ymin, ymax, zmin, zmax); density is [n_dust, n_cells]
    voronoi: Optional[dict] = None
    # frequency-resolved specific energy (Model.set_specific_energy_spectrum_bins, /specific_energy_spectrum_bin_edges
    # of the .rtin): n + 1 strictly increasing frequencies; None = not computed
    spectrum_bin_edges: Optional[np.ndarray] = None

    def __post_init__(self):
        self.density = _f8(self.density)
        if self.grid_type == "vor":
            v = self.voronoi
            for k in ("coordinates", "bb_min", "bb_max"):
                v[k] = np.ascontiguousarray(v[k], dtype=np.float64).reshape(-1, 3)
            v["volume"] = _f8(v["volume"])
            v["sparse_neighs"] = np.ascontiguousarray(v["sparse_neighs"], dtype=np.int32)
            v["sparse_idx"] = np.ascontiguousarray(v["sparse_idx"], dtype=np.int32)
            v["box"] = _f8(v["box"])
            if self.density.ndim == 1:
                self.density = self.density[None]
            if self.density.shape != (len(self.dust), len(v["volume"])):
                raise ValueError("density should have shape (n_dust, n_cells)")
            return
        if self.grid_type == "amr":
            n = sum(g[0] * g[1] * g[2] for lev in self.amr_levels for g in lev)
            if self.density.ndim == 1:
                self.density = self.density[None]
            if self.density.shape != (len(self.dust), n):
                raise ValueError("density should have shape (n_dust, n_cells) = %s" % ((len(self.dust), n),))
            return
        if self.grid_type == "oct":
            self.refined = np.ascontiguousarray(self.refined, dtype=np.int32)
            if self.density.ndim == 1:
                self.density = self.density[None]
            if self.density.shape != (len(self.dust), len(self.refined)):
                raise ValueError("density should have shape (n_dust, n_nodes)")
            return
        self.w1, self.w2, self.w3 = _f8(self.w1), _f8(self.w2), _f8(self.w3)
        n1, n2, n3 = len(self.w1) - 1, len(self.w2) - 1, len(self.w3) - 1
        if self.density.ndim == 3:
            self.density = self.density[None]
        if self.density.shape != (len(self.dust), n3, n2, n1):
            raise ValueError("density should have shape (n_dust, n3, n2, n1) = %s, got %s" %
                             ((len(self.dust), n3, n2, n1), self.density.shape))

    @property
    def shape(self):
        if self.grid_type == "vor":
            return (len(self.voronoi["volume"]),)
        if self.grid_type == "oct":
            return (len(self.refined),)
        if self.grid_type == "amr":
            return (sum(g[0] * g[1] * g[2] for lev in self.amr_levels for g in lev),)
        return (len(self.w3) - 1, len(self.w2) - 1, len(self.w1) - 1)

    @property
    def n_cells(self):
        return int(np.prod(self.shape))

    def amr_slices(self):
        """[(level, grid, slice into the flat cell axis, (n3, n2, n1))] in cell-id order."""
        out, start = [], 0
        for il, lev in enumerate(self.amr_levels):
            for ig, g in enumerate(lev):
                n = g[0] * g[1] * g[2]
                out.append((il, ig, slice(start, start + n), (g[2], g[1], g[0])))
                start += n
        return out

    def volumes(self):
        if self.grid_type == "vor":
            return np.maximum(self.voronoi["volume"], 0.0)
        if self.grid_type == "amr":
            vol = np.zeros(self.n_cells)
            for il, ig, sl, _ in self.amr_slices():
                g = self.amr_levels[il][ig]
                vol[sl] = ((g[4] - g[3]) / g[0]) * ((g[6] - g[5]) / g[1]) * ((g[8] - g[7]) / g[2])
            return vol
        if self.grid_type == "oct":
            # node volumes in depth-first order (grid_geometry_octree.f90:160-183,250-253)
            vol = np.zeros(len(self.refined))
            hx, hy, hz = self.oct_half
            stack = [[0, 8 * hx * hy * hz]]
            vol[0] = stack[0][1]
            idx = 0
            pending = [(8, vol[0] / 8.)] if self.refined[0] else []
            while pending:
                left, v = pending[-1]
                if left == 0:
                    pending.pop()
                    continue
                pending[-1] = (left - 1, v)
                idx += 1
                vol[idx] = v
                if self.refined[idx]:
                    pending.append((8, v / 8.))
            return vol
        if self.grid_type == "sph":
            # grid_geometry_spherical_3d.f90:147-160
            dr3, dcost, dphi = np.diff(self.w1 ** 3), -np.diff(np.cos(self.w2)), np.diff(self.w3)
            return dr3[None, None, :] * dcost[None, :, None] * dphi[:, None, None] / 3.
        if self.grid_type == "cyl":
            # grid_geometry_cylindrical_3d.f90:141-147
            dw2, dz, dphi = np.diff(self.w1 ** 2), np.diff(self.w2), np.diff(self.w3)
            return dw2[None, None, :] * dz[None, :, None] * dphi[:, None, None] / 2.
        dx, dy, dz = np.diff(self.w1), np.diff(self.w2), np.diff(self.w3)
        return (dx[None, None, :] * dy[None, :, None]) * dz[:, None, None]


def apply_model(api, ctx, model: FlatModel):
    """Push a FlatModel through a C API object exposing the hyp_*/orc_* setters.

    ``api`` is a binding object with methods named like the header's functions
    minus the prefix (see :mod:`hyperion_b200.capi`)."""
    if model.grid_type == "amr":
        api.set_grid_amr(ctx, model.amr_levels)
        n1 = n2 = n3 = 0
    elif model.grid_type == "oct":
        api.set_grid_octree(ctx, model.refined, model.oct_center, model.oct_half)
        n1 = n2 = n3 = 0
    elif model.grid_type == "vor":
        api.set_grid_voronoi(ctx, model.voronoi)
        n1 = n2 = n3 = 0
    else:
        n3, n2, n1 = model.shape
    if model.grid_type in ("oct", "amr", "vor"):
        pass
    elif model.grid_type == "sph":
        api.set_grid_spherical(ctx, n1, n2, n3, model.w1, model.w2, model.w3)
    elif model.grid_type == "cyl":
        api.set_grid_cylindrical(ctx, n1, n2, n3, model.w1, model.w2, model.w3)
    else:
        api.set_grid_cartesian(ctx, n1, n2, n3, model.w1, model.w2, model.w3)
    for d in model.dust:
        api.add_dust(ctx, d)
    for s in model.sources:
        api.add_source(ctx, s)
    api.set_run_conf(ctx, model.conf)
    if model.spectrum_bin_edges is not None:
        api.set_specific_energy_spectrum_bins(ctx, model.spectrum_bin_edges)
    api.set_density(ctx, len(model.dust), model.density)
    api.set_specific_energy(ctx, model.specific_energy, model.minimum_specific_energy)
    if model.frequencies is not None:
        api.set_monochromatic(ctx, model.frequencies, model.monochromatic_energy_threshold)
    for g in model.peeled:
        api.add_peeled_group(ctx, g)
    if model.binned is not None:
        api.add_peeled_group(ctx, model.binned)
