"""Parse a Hyperion ``.rtin`` model file into the neutral in-memory model.

Host-side replacement of ``setup_initial`` (``src/main/setup_rt.f90:27-304``) and the readers
it calls (``setup_dust`` ``src/dust/dust.f90:30``, ``setup_grid_geometry``
``src/grid/grid_geometry_cartesian_3d.f90:77-135``, ``setup_grid_physics``
``src/grid/grid_physics_3d.f90:111-322``, ``setup_sources`` ``src/sources/source.f90:48-80``).
The file layout is the one ``hyperion.model.Model.write`` produces
(``hyperion/model/model.py:513-740``; ``docs/advanced/model_file.rst``).

Errors use the reference's wording where its tests look for it
(``hyperion/model/tests/test_fortran.py``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .flatmodel import FlatConf, FlatDust, FlatModel, FlatPeeledGroup, FlatSource
from .io import h5min


class ModelError(RuntimeError):
    """A condition the reference reports through ``error()`` (``fortranlib/src/lib_messages.f90:126-179``)."""


def _s(v):
    """HDF5 string attribute -> python str."""
    if isinstance(v, np.ndarray):
        v = v.reshape(-1)[0] if v.size else b""
    if isinstance(v, bytes):
        return v.split(b"\x00")[0].decode("utf-8").strip()
    return str(v).strip()


def _yes(v):
    """Booleans are the strings yes/no (``hyperion/util/functions.py:28-33``; accepted spellings
    ``fortranlib/src/lib_hdf5_110.f90:2880-2884``)."""
    t = _s(v).lower()
    if t in ("yes", "y", "true"):
        return True
    if t in ("no", "n", "false"):
        return False
    raise ModelError("Unknown logical value: %s" % t)


def _num(v):
    a = np.asarray(v).reshape(-1)
    return a[0]


@dataclass
class RunSettings:
    """Root / Output attributes beyond what the photon kernels need (``setup_rt.f90:38-157``)."""
    n_initial_iter: int = 0
    n_initial_photons: int = 0
    n_last_photons: int = 0
    n_last_photons_sources: int = 0     # monochromatic mode: per frequency (setup_rt.f90:178,241)
    n_last_photons_dust: int = 0
    n_ray_photons_sources: int = 0
    n_ray_photons_dust: int = 0
    n_stats: int = 0
    raytracing: bool = False
    monochromatic: bool = False
    pda: bool = False
    forced_first_interaction: bool = True
    forced_first_interaction_algorithm: str = "wr99"
    baes16_xi: float = 0.5
    check_convergence: bool = False
    convergence_absolute: float = 0.0
    convergence_relative: float = 0.0
    convergence_percentile: float = 100.0
    specific_energy_type: str = "initial"
    physics_io_bytes: int = 8
    copy_input: bool = True
    output_specific_energy: str = "last"
    output_density: str = "none"
    output_density_diff: str = "none"
    output_n_photons: str = "none"
    output_specific_energy_spectrum: str = "none"
    geometry_id: str = ""
    grid_type: str = "car"
    extra: dict = field(default_factory=dict)


def _attr(attrs, name, default=None, required=False):
    if name in attrs:
        return attrs[name]
    if required:
        raise ModelError("attribute %s is missing from the input file" % name)
    return default


def read_rtin(filename):
    """Returns (FlatModel, RunSettings, h5min.File).  Only what the implemented hot path needs is
    interpreted; unsupported options raise ModelError instead of being silently ignored."""
    f = h5min.File(filename)
    A = f.attrs
    if "python_version" not in A:
        raise ModelError("cannot read files made with the Python module before version 0.8.7")

    rs = RunSettings()
    rs.monochromatic = _yes(_attr(A, "monochromatic", b"no"))
    rs.raytracing = _yes(_attr(A, "raytracing", b"no"))
    rs.n_stats = int(_num(_attr(A, "n_stats", 0)))
    rs.pda = _yes(_attr(A, "pda", b"no"))
    conf = FlatConf()
    conf.seed = int(_num(_attr(A, "seed", -124902)))
    conf.n_inter_max = int(_num(_attr(A, "n_inter_max", required=True)))
    conf.n_reabs_max = int(_num(_attr(A, "n_reabs_max", required=True)))
    conf.use_mrw = _yes(_attr(A, "mrw", b"no"))
    if conf.use_mrw:
        conf.mrw_gamma = float(_num(_attr(A, "mrw_gamma", required=True)))
        conf.n_mrw_max = int(_num(_attr(A, "n_inter_mrw_max", required=True)))
    conf.kill_on_absorb = _yes(_attr(A, "kill_on_absorb", b"no"))
    conf.kill_on_scatter = _yes(_attr(A, "kill_on_scatter", b"no"))
    if "forced_first_scattering" in A:
        rs.forced_first_interaction = _yes(A["forced_first_scattering"])
    else:
        rs.forced_first_interaction = _yes(_attr(A, "forced_first_interaction", b"yes"))
        rs.forced_first_interaction_algorithm = _s(_attr(A, "forced_first_interaction_algorithm", b"wr99")).lower()
        if rs.forced_first_interaction_algorithm == "baes16":
            rs.baes16_xi = float(_num(_attr(A, "forced_first_interaction_baes16_xi", required=True)))
        elif rs.forced_first_interaction_algorithm != "wr99":
            raise ModelError("Unknown forced first interaction algorithm: " + rs.forced_first_interaction_algorithm)
    conf.propagation_check_frequency = float(_num(_attr(A, "propagation_check_frequency", 1.e-3)))
    conf.sample_sources_evenly = _yes(_attr(A, "sample_sources_evenly", b"no"))
    conf.enforce_energy_range = _yes(_attr(A, "enforce_energy_range", b"yes"))

    # photon counts may be stored as floats (hyperion/conf/conf_files.py:142-227)
    rs.n_initial_iter = int(_num(_attr(A, "n_initial_iter", 0)))
    if rs.n_initial_iter > 0:
        rs.n_initial_photons = int(float(_num(_attr(A, "n_initial_photons", required=True))))
        if rs.n_initial_photons == 0:
            raise ModelError("Number of initial iterations is non-zero, but number of specific_energy photons is zero")
    rs.n_last_photons = int(float(_num(_attr(A, "n_last_photons", 0))))
    if rs.monochromatic:
        # setup_rt.f90:49-56,178,241: the imaging iteration runs n_last_photons_sources / _dust packets per frequency
        rs.n_last_photons = 0
        rs.n_last_photons_sources = int(float(_num(_attr(A, "n_last_photons_sources", 0))))
        rs.n_last_photons_dust = int(float(_num(_attr(A, "n_last_photons_dust", 0))))
    if rs.raytracing:
        rs.n_ray_photons_sources = int(float(_num(_attr(A, "n_ray_photons_sources", 0))))
        rs.n_ray_photons_dust = int(float(_num(_attr(A, "n_ray_photons_dust", 0))))
    conf.n_initial_iter = rs.n_initial_iter
    conf.n_initial_photons = rs.n_initial_photons
    rs.specific_energy_type = _s(_attr(A, "specific_energy_type", b"initial"))
    if rs.specific_energy_type not in ("initial", "additional"):
        raise ModelError("specific_energy_type should be 'additional' or 'initial'")
    rs.physics_io_bytes = int(_num(_attr(A, "physics_io_bytes", 8)))
    if rs.physics_io_bytes not in (4, 8):
        raise ModelError("unexpected value of physics_io_bytes (should be 4 or 8)")
    rs.copy_input = _yes(_attr(A, "copy_input", b"yes"))
    if rs.n_initial_iter > 0:
        rs.check_convergence = _yes(_attr(A, "check_convergence", b"no"))
        if rs.check_convergence:
            rs.convergence_absolute = float(_num(A["convergence_absolute"]))
            rs.convergence_relative = float(_num(A["convergence_relative"]))
            rs.convergence_percentile = float(_num(A["convergence_percentile"]))

    # ---- dust
    dust = []
    if "Dust" in f:
        g_dust = f["Dust"]
        for name in sorted(g_dust.keys()):
            dust.append(FlatDust.from_hdf5_group(g_dust[name]))

    # ---- grid geometry
    geo = f["Grid/Geometry"]
    rs.grid_type = _s(geo.attrs["grid_type"])
    rs.geometry_id = _s(geo.attrs["geometry"])
    octree = None
    amr_levels = None
    if rs.grid_type == "amr":
        # grid_geometry_amr.f90:111-187: nlevels; level_%05d: ngrids; grid_%05d: n1..n3, xmin..zmax
        amr_levels = []
        for il in range(int(_num(geo.attrs["nlevels"]))):
            gl = geo["level_%05d" % (il + 1)]
            grids = []
            for ig in range(int(_num(gl.attrs["ngrids"]))):
                a = gl["grid_%05d" % (ig + 1)].attrs
                grids.append(tuple(int(_num(a[k])) for k in ("n1", "n2", "n3")) +
                             tuple(float(_num(a[k])) for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")))
            amr_levels.append(grids)
        octree = dict(amr_levels=amr_levels)
        grid_type, w1, w2, w3 = "amr", None, None, None
    elif rs.grid_type == "oct":
        # grid_geometry_octree.f90:189-262: table 'cells' (column 'refined'), root cell centre and half-widths
        refined = np.asarray(geo["cells"][...]["refined"], dtype=np.int32)
        octree = dict(refined=refined,
                      oct_center=tuple(float(_num(geo.attrs[k])) for k in ("x", "y", "z")),
                      oct_half=tuple(float(_num(geo.attrs[k])) for k in ("dx", "dy", "dz")))
        if (len(refined) - 1) % 8 != 0:
            raise ModelError("refined should have shape 8 * n + 1")
        grid_type, w1, w2, w3 = "oct", None, None, None
    elif rs.grid_type == "vor":
        # grid_geometry_voronoi.f90:92-187: table 'cells' (columns 'coordinates', 'bb_min', 'bb_max', 'volume'),
        # 'sparse_neighs' / 'sparse_idx', the box as attributes
        cells = geo["cells"][...]
        mesh = dict(coordinates=np.asarray(cells["coordinates"], dtype=np.float64),
                    bb_min=np.asarray(cells["bb_min"], dtype=np.float64),
                    bb_max=np.asarray(cells["bb_max"], dtype=np.float64),
                    volume=np.asarray(cells["volume"], dtype=np.float64),
                    sparse_neighs=np.asarray(geo["sparse_neighs"][...], dtype=np.int32),
                    sparse_idx=np.asarray(geo["sparse_idx"][...], dtype=np.int32),
                    box=np.array([float(_num(geo.attrs[k])) for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")]))
        if len(mesh["sparse_idx"]) != len(mesh["volume"]) + 1:
            raise ModelError("sparse_idx should have one entry per cell plus one")
        octree = dict(voronoi=mesh)
        grid_type, w1, w2, w3 = "vor", None, None, None
    else:
        if rs.grid_type == "car":
            cols, names, grid_type = ("x", "y", "z"), ("dx", "dy", "dz"), "car"
        elif rs.grid_type == "sph_pol":
            # grid_geometry_spherical_3d.f90:111-128
            cols, names, grid_type = ("r", "t", "p"), ("dr", "dt", "dphi"), "sph"
        elif rs.grid_type == "cyl_pol":
            # grid_geometry_cylindrical_3d.f90:109-121
            cols, names, grid_type = ("w", "z", "p"), ("dw", "dz", "dphi"), "cyl"
        else:
            raise ModelError("unknown grid type '%s'" % rs.grid_type)
        w1 = np.asarray(geo["walls_1"][...][cols[0]], dtype=np.float64)
        w2 = np.asarray(geo["walls_2"][...][cols[1]], dtype=np.float64)
        w3 = np.asarray(geo["walls_3"][...][cols[2]], dtype=np.float64)
        if grid_type == "sph":
            if np.any(w1 < 0.):
                raise ModelError("r walls should be positive")
            if np.any(w2 < 0.) or np.any(w2 > np.pi):
                raise ModelError("theta walls should be between 0 and pi")
            if np.any(w3 < 0.) or np.any(w3 > 2 * np.pi):
                raise ModelError("phi walls should be between 0 and 2*pi")
        if grid_type == "cyl":
            if np.any(w1 < 0.):
                raise ModelError("w walls should be positive")
            if np.any(w3 < 0.) or np.any(w3 > 2 * np.pi):
                raise ModelError("phi walls should be between 0 and 2*pi")
        for w, nm in zip((w1, w2, w3), names):
            if np.any(np.diff(w) <= 0):
                raise ModelError("all %s values should be greater than zero" % nm)

    # ---- grid physics
    q = f["Grid/Quantities"]
    if amr_levels is not None:
        # one dataset per grid, [n_dust, n3, n2, n1]; flattened to [n_dust, n_cells] in cell-id order
        def gather(name):
            parts = []
            for il, lev in enumerate(amr_levels):
                for ig, g in enumerate(lev):
                    path = "level_%05d/grid_%05d/%s" % (il + 1, ig + 1, name)
                    if path not in q:
                        return None
                    a = np.asarray(q[path][...], dtype=np.float64)
                    if a.shape[1:] != (g[2], g[1], g[0]):
                        raise ModelError("%s array has wrong shape" % name)
                    parts.append(a.reshape(a.shape[0], -1))
            return np.concatenate(parts, axis=1)
        density = gather("density")
        if density is None:
            density = np.zeros((0, sum(g[0] * g[1] * g[2] for lev in amr_levels for g in lev)))
        if density.shape[0] != len(dust):
            raise ModelError("density array has wrong number of dust types")
        if np.any(density < 0):
            raise ModelError("density should be positive")
        se_amr = gather("specific_energy")
    grid_shape = (len(octree["refined"]),) if octree and "refined" in octree else \
        (len(octree["voronoi"]["volume"]),) if octree and "voronoi" in octree else \
        (len(w3) - 1, len(w2) - 1, len(w1) - 1) if amr_levels is None else None
    if amr_levels is not None:
        pass
    elif "density" in q:
        dset = q["density"]
        if "geometry" in dset.attrs and _s(dset.attrs["geometry"]) != rs.geometry_id:
            raise ModelError("geometry id of density does not match that of the grid")
        density = np.asarray(dset[...], dtype=np.float64)
        if density.shape[1:] != grid_shape:
            raise ModelError("density array has wrong shape")
        if density.shape[0] != len(dust):
            raise ModelError("density array has wrong number of dust types")
        if np.any(density < 0):
            raise ModelError("density should be positive")
    else:
        density = np.zeros((0,) + grid_shape)
    se = None
    if amr_levels is not None:
        se = se_amr
    elif "specific_energy" in q:
        se = np.asarray(q["specific_energy"][...], dtype=np.float64)
        if np.any(se < 0):
            raise ModelError("specific_energy should be positive")
    min_e = None
    if "minimum_specific_energy" in q.attrs:
        min_e = np.asarray(q.attrs["minimum_specific_energy"], dtype=np.float64).reshape(-1)

    # ---- sources
    sources = []
    if "Sources" in f:
        g_src = f["Sources"]
        for name in sorted(g_src.keys()):
            g = g_src[name]
            a = g.attrs
            stype = _s(a["type"])
            # source_read (src/sources/source_type.f90:102-282); type numbers are the reference's
            types = {"point": 1, "sphere": 2, "map": 4, "extern_sph": 5, "extern_box": 6, "plane_parallel": 7,
                     "point_collection": 8}
            if stype not in types:
                raise ModelError("unknown type in source list: " + stype)
            spots = []
            if stype == "sphere":
                # source_read (source_type.f90:150-188): every sub-group of a sphere is a spot
                for k in sorted(g.keys()):
                    if not isinstance(g[k], type(g)):
                        continue
                    sa = g[k].attrs
                    d = dict(luminosity=float(_num(sa["luminosity"])), longitude=float(_num(sa["longitude"])),
                             latitude=float(_num(sa["latitude"])), radius=float(_num(sa["radius"])))
                    sspec = _s(sa["spectrum"])
                    if sspec == "temperature":
                        d["temperature"] = float(_num(sa["temperature"]))
                    elif sspec == "spectrum":
                        t = g[k]["spectrum"][...]
                        d["spectrum_nu"] = np.asarray(t["nu"], dtype=np.float64)
                        d["spectrum_fnu"] = np.asarray(t["fnu"], dtype=np.float64)
                    else:
                        raise ModelError("Spot cannot have LTE spectrum")
                    spots.append(d)
            spec = _s(a["spectrum"])
            kw = dict(type=types[stype], peeloff=_yes(a["peeloff"]))
            if stype == "point_collection":
                pos = np.asarray(g["position"][...], dtype=np.float64)
                lum = np.asarray(g["luminosity"][...], dtype=np.float64).reshape(-1)
                if pos.ndim != 2 or pos.shape[1] != 3 or pos.shape[0] != lum.shape[0]:
                    raise ModelError("point collection: position should be (n, 3) and luminosity (n,)")
                kw["points"], kw["points_luminosity"] = pos, lum
                kw["luminosity"] = float(lum.sum())
            else:
                kw["luminosity"] = float(_num(a["luminosity"]))
            if stype == "map":
                # grid_load_pdf_map -> read_grid_3d (src/grid/grid_geometry_common_3d.f90:47-63)
                if amr_levels is not None:
                    # one 'Luminosity map' dataset per level_%05i/grid_%05i (hyperion/grid/amr_grid.py:422-476),
                    # flattened in cell-id order
                    parts = []
                    for il, lev in enumerate(amr_levels):
                        for ig, gr in enumerate(lev):
                            a_ = np.asarray(g["level_%05d/grid_%05d/Luminosity map" % (il + 1, ig + 1)][...], dtype=np.float64)
                            if a_.shape != (gr[2], gr[1], gr[0]):
                                raise ModelError("Luminosity map has wrong shape")
                            parts.append(a_.reshape(-1))
                    lm = np.concatenate(parts)
                else:
                    lm = np.asarray(g["Luminosity map"][...], dtype=np.float64)
                    if lm.shape != grid_shape:
                        raise ModelError("Luminosity map has wrong shape")
                kw["map"] = lm
            if stype in ("point", "sphere", "extern_sph", "plane_parallel"):
                kw["position"] = (float(_num(a["x"])), float(_num(a["y"])), float(_num(a["z"])))
            if stype in ("sphere", "extern_sph", "plane_parallel"):
                kw["radius"] = float(_num(a["r"]))
            if stype == "sphere":
                kw["limb_darkening"] = _yes(a["limb"])
                if spots:
                    kw["spots"] = spots
            if stype == "extern_box":
                kw["bounds"] = tuple(float(_num(a[k])) for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"))
            if stype == "plane_parallel":
                kw["direction"] = (float(_num(a["theta"])), float(_num(a["phi"])))
            if spec == "temperature":
                kw["temperature"] = float(_num(a["temperature"]))
            elif spec == "spectrum":
                t = g["spectrum"][...]
                nu, fnu = np.asarray(t["nu"], dtype=np.float64), np.asarray(t["fnu"], dtype=np.float64)
                if np.any(np.diff(nu) < 0):
                    raise ModelError("spectrum frequency should be monotonically increasing")
                kw["spectrum_nu"], kw["spectrum_fnu"] = nu, fnu
            elif spec == "lte" and stype == "map":
                kw["lte"] = True
            elif spec == "lte":
                raise ModelError({"point": "Point source", "sphere": "Spherical source",
                                  "extern_sph": "External spherical source", "extern_box": "External box source",
                                  "plane_parallel": "Plane parallel",
                                  "point_collection": "Point source collection"}[stype] + " cannot have LTE spectrum")
            else:
                raise ModelError("unknown spectrum specifier: " + spec)
            sources.append(FlatSource(**kw))
    if not sources and rs.n_initial_iter > 0:
        raise ModelError("no sources set up - need sources for initial iteration(s)")

    # ---- output switches
    out = f["Output"].attrs if "Output" in f else {}
    for key in ("output_density", "output_density_diff", "output_specific_energy", "output_n_photons"):
        if key in out:
            val = _s(out[key])
            if val not in ("all", "last", "none"):
                raise ModelError("%s should be one of all/last/none" % key)
            setattr(rs, key, val)

    # setup_initial (src/main/setup_rt.f90:78-104): the frequency-resolved specific energy is computed whenever it is
    # written, and then needs its bin edges (table column 'nu' at the root of the file)
    spectrum_bin_edges = None
    if "output_specific_energy_spectrum" in out:
        val = _s(out["output_specific_energy_spectrum"])
        if val not in ("all", "last", "none"):
            raise ModelError("output_specific_energy_spectrum should be one of all/last/none")
        rs.output_specific_energy_spectrum = val
    if rs.output_specific_energy_spectrum != "none":
        if "specific_energy_spectrum_bin_edges" not in f:
            raise ModelError("specific_energy_spectrum_bin_edges should be present in the input when "
                             "output_specific_energy_spectrum is enabled")
        spectrum_bin_edges = np.asarray(f["specific_energy_spectrum_bin_edges"][...]["nu"], dtype=np.float64)
        if np.any(np.diff(spectrum_bin_edges) <= 0):     # setup_grid_physics, grid_physics_3d.f90:126-128
            raise ModelError("specific_energy_spectrum_bin_edges should be strictly increasing")

    conf.forced_first_interaction = rs.forced_first_interaction
    conf.forced_first_interaction_algorithm = rs.forced_first_interaction_algorithm
    conf.baes16_xi = rs.baes16_xi
    if sources and not rs.monochromatic and rs.n_last_photons == 0 and "n_last_photons" not in A:
        raise ModelError("attribute n_last_photons is missing from the input file")
    if rs.monochromatic:
        if sources and "n_last_photons_sources" not in A:
            raise ModelError("attribute n_last_photons_sources is missing from the input file")
        if dust and "n_last_photons_dust" not in A:
            raise ModelError("attribute n_last_photons_dust is missing from the input file")
        if not sources:
            rs.n_last_photons_sources = 0
    if not sources and rs.n_last_photons > 0:
        raise ModelError("no sources set up - need sources for last iteration")

    no_dust = len(dust) == 0
    if no_dust:
        # setup_initial (src/main/setup_rt.f90:164-168): a model without dust skips the initial iterations and
        # the thermal raytracing, and still images its sources.  The engine marches a vacuum: one placeholder
        # dust type (grey, covering every frequency a source can emit) with zero density in every cell.
        print(" WARNING: no dust present, so skipping initial iterations [main]", flush=True)
        rs.n_initial_iter = 0
        rs.n_initial_photons = 0
        rs.n_ray_photons_dust = 0
        rs.n_last_photons_dust = 0
        rs.check_convergence = False
        conf.n_initial_iter = 0
        conf.n_initial_photons = 0
        from .synthetic import make_dust
        dust = [make_dust([1.e-2, 1.e30], [0., 0.], [1., 1.], n_temp=4, temp_min=0.1, temp_max=1.e5)]
        density = np.zeros((1,) + tuple(density.shape[1:]))
        se, min_e = None, None

    model = FlatModel(w1, w2, w3, density, dust, sources, conf, specific_energy=se, minimum_specific_energy=min_e,
                      grid_type=grid_type, **(octree or {}))
    model.no_dust = no_dust
    model.spectrum_bin_edges = spectrum_bin_edges
    if rs.monochromatic:
        # setup_rt.f90:220-222, hyperion/model/model.py:133-137
        if "frequencies" not in f:
            raise ModelError("frequencies should be given if use_exact_nu is .true.")
        model.frequencies = np.asarray(f["frequencies"][...]["nu"], dtype=np.float64)
        model.monochromatic_energy_threshold = float(_num(_attr(A, "monochromatic_energy_threshold", 1.e-10)))
    model.peeled = read_peeled_groups(f, rs.monochromatic)
    if "Output" in f and "Binned" in f["Output"] and len(f["Output"]["Binned"].keys()) > 0:
        # setup_final_iteration (src/main/setup_rt.f90:314-331)
        names = sorted(f["Output"]["Binned"].keys())
        if len(names) > 1:
            raise ModelError("can't have more than one binned image group")
        if rs.forced_first_interaction:
            raise ModelError("can't use binned images with forced first interaction")
        g = f["Output"]["Binned"][names[0]]
        kw = _image_conf(g, rs.monochromatic)
        kw.update(binned=True, n_theta=int(_num(_attr(g.attrs, "n_theta", required=True))),
                  n_phi=int(_num(_attr(g.attrs, "n_phi", required=True))))
        model.binned = FlatPeeledGroup(**kw)
    return model, rs, f


def _image_conf(g, monochromatic=False):
    """The part of an image group every kind shares (``image_setup``, ``src/images/image_type.f90:153-335``)."""
    a = g.attrs
    kw = {}
    if monochromatic and "use_filters" in a and _yes(a["use_filters"]):
        raise ModelError("cannot use filters in monochromatic mode")
    if monochromatic:
        # image_setup (image_type.f90:243-258): the channels are the frequencies inu_min .. inu_max of /frequencies
        n_wav = int(_num(_attr(a, "n_wav", required=True)))
        if n_wav < 1:
            raise ModelError("n_nu should be >= 1")
        kw["wavelengths"] = (n_wav, 1.0, 1.0)
        kw["inu_min"] = int(_num(_attr(a, "inu_min", required=True)))
        kw["inu_max"] = int(_num(_attr(a, "inu_max", required=True)))
    elif "use_filters" in a and _yes(a["use_filters"]):
        # image_setup (image_type.f90:174-183, 274-284): n_filt tables filter_%05i(nu, tn) with attribute nu0
        n_filt = int(_num(_attr(a, "n_filt", required=True)))
        if n_filt < 1:
            raise ModelError("n_nu should be >= 1")
        filters = []
        for i in range(n_filt):
            d = g["filter_%05i" % (i + 1)]
            t = d[...]
            filters.append((np.asarray(t["nu"], dtype=np.float64), np.asarray(t["tn"], dtype=np.float64),
                            float(_num(d.attrs["nu0"]))))
        kw["filters"] = filters
        kw["wavelengths"] = (n_filt, 1.0, 1.0)
    else:
        n_wav = int(_num(_attr(a, "n_wav", required=True)))
        if n_wav < 1:
            raise ModelError("n_nu should be >= 1")
        kw["wavelengths"] = (n_wav, float(_num(_attr(a, "wav_min", required=True))), float(_num(_attr(a, "wav_max", required=True))))
    kw["stokes"] = _yes(a["compute_stokes"]) if "compute_stokes" in a else True
    if _yes(_attr(a, "compute_image", required=True)):
        kw["image"] = (int(_num(a["n_x"])), int(_num(a["n_y"])), float(_num(a["x_min"])), float(_num(a["x_max"])),
                       float(_num(a["y_min"])), float(_num(a["y_max"])))
    if _yes(_attr(a, "compute_sed", required=True)):
        kw["sed"] = (int(_num(a["n_ap"])), float(_num(a["ap_min"])), float(_num(a["ap_max"])))
    kw["track_origin"] = _s(_attr(a, "track_origin", required=True))
    if kw["track_origin"] not in ("no", "basic", "yes", "detailed", "scatterings"):
        raise ModelError("unknown track_origin flag: " + kw["track_origin"])
    kw["track_n_scat"] = int(_num(a["track_n_scat"])) if "track_n_scat" in a else 0
    kw["uncertainties"] = _yes(_attr(a, "uncertainties", required=True))
    kw["io_bytes"] = int(_num(_attr(a, "io_bytes", required=True)))
    if kw["io_bytes"] not in (4, 8):
        raise ModelError("unexpected value of io_bytes (should be 4 or 8)")
    return kw


def read_peeled_groups(f, monochromatic=False):
    """``setup_final_iteration`` / ``peeled_images_setup`` / ``image_setup``
    (``src/main/setup_rt.f90:306-347``, ``src/images/images_peeled.f90:272-382``,
    ``src/images/image_type.f90:153-335``): one FlatPeeledGroup per ``Output/Peeled/group_%05i``."""
    groups = []
    if "Output" not in f or "Peeled" not in f["Output"]:
        return groups
    gp = f["Output"]["Peeled"]
    for name in sorted(gp.keys()):
        g = gp[name]
        a = g.attrs
        n_view = int(_num(_attr(a, "n_view", required=True)))
        if not n_view > 0:
            raise ModelError("n_view should be a positive integer")
        inside = _yes(_attr(a, "inside_observer", required=True))
        ang = g["angles"][...]
        kw = dict(theta=np.asarray(ang["theta"], dtype=np.float64), phi=np.asarray(ang["phi"], dtype=np.float64),
                  inside_observer=inside, ignore_optical_depth=_yes(_attr(a, "ignore_optical_depth", b"no")),
                  d_min=float(_num(_attr(a, "d_min", required=True))), d_max=float(_num(_attr(a, "d_max", required=True))),
                  peeloff_origin=tuple(float(_num(_attr(a, k, required=True))) for k in ("peeloff_x", "peeloff_y", "peeloff_z")))
        if len(kw["theta"]) != n_view:
            raise ModelError("n_view does not match the length of the angles table")
        kw.update(_image_conf(g, monochromatic))
        groups.append(FlatPeeledGroup(**kw))
    return groups
