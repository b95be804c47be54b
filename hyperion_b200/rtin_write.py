"""Write a FlatModel as a ``.rtin`` file in the layout ``hyperion.model.Model.write`` produces
(``hyperion/model/model.py:513-740``, ``hyperion/grid/cartesian_grid.py:295-358``,
``hyperion/sources/source.py:264-284``, ``hyperion/dust/dust_type.py:377-443``,
``hyperion/conf/conf_files.py:795-822``).

The reference front end needs h5py + astropy to do this; where those are missing (this engine's
test and benchmark boxes) this writer produces the same file so that the drop-in binary can be
exercised end to end on synthetic models.
"""
from __future__ import annotations

import hashlib

import numpy as np

from .flatmodel import FlatModel
from .io import h5write


def _yn(b):
    return b"yes" if b else b"no"


def _table(cols):
    dt = []
    n = None
    for name, a in cols:
        a = np.asarray(a)
        n = len(a)
        dt.append((name, a.dtype.str) if a.ndim == 1 else (name, a.dtype.str, a.shape[1:]))
    t = np.zeros(n, dtype=dt)
    for name, a in cols:
        t[name] = a
    return t


def write_dust(g, d):
    """``SphericalDust.write`` (``hyperion/dust/dust_type.py:377-443``): a version-2 dust group."""
    g.attrs["version"] = np.int32(2 if d.version != 1 else 1)
    g.attrs["type"] = np.int32(1)
    g.attrs["python_version"] = "0.9.12"
    g.attrs["emissvar"] = "E"
    g.attrs["lte"] = _yn(d.is_lte)
    g.attrs["sublimation_mode"] = {0: "no", 1: "fast", 2: "slow", 3: "cap"}[d.sublimation_mode]
    if d.sublimation_mode:
        g.attrs["sublimation_specific_energy"] = float(d.sublimation_specific_energy)
    g.create_dataset("optical_properties", _table([("nu", d.nu), ("albedo", d.albedo), ("chi", d.chi),
                                                   ("P1", d.P1), ("P2", d.P2), ("P3", d.P3), ("P4", d.P4)]))
    g.create_dataset("scattering_angles", _table([("mu", d.mu)]))
    temperature = getattr(d, "temperature", np.zeros_like(d.specific_energy))
    if d.version == 1:
        cols = [("specific_energy", d.specific_energy), ("chi_planck", d.chi_planck), ("kappa_planck", d.kappa_planck),
                ("chi_rosseland", d.chi_inv_planck), ("kappa_rosseland", d.kappa_inv_planck)]
    else:
        cols = [("temperature", temperature), ("specific_energy", d.specific_energy),
                ("chi_planck", d.chi_planck), ("kappa_planck", d.kappa_planck),
                ("chi_inv_planck", d.chi_inv_planck), ("kappa_inv_planck", d.kappa_inv_planck),
                ("chi_rosseland", d.chi_rosseland), ("kappa_rosseland", d.kappa_rosseland)]
    g.create_dataset("mean_opacities", _table(cols))
    g.create_dataset("emissivities", _table([("nu", d.emiss_nu), ("jnu", d.emiss_jnu)]))
    g.create_dataset("emissivity_variable", _table([("specific_energy", d.jnu_var)]))


def write_peeled_group(g, p):
    """``PeeledImageConf.write`` (``hyperion/conf/conf_files.py:795-1130,1293-1345``)."""
    a = g.attrs
    if p.binned:
        # BinnedImageConf (hyperion/conf/conf_files.py:1242-1275)
        a["n_theta"], a["n_phi"] = np.int64(p.n_theta), np.int64(p.n_phi)
    else:
        a["n_view"] = np.int64(len(p.theta))
        g.create_dataset("angles", _table([("theta", np.asarray(p.theta, dtype=np.float64)),
                                           ("phi", np.asarray(p.phi, dtype=np.float64))]))
        a["inside_observer"] = _yn(p.inside_observer)
        a["ignore_optical_depth"] = _yn(p.ignore_optical_depth)
        a["peeloff_x"], a["peeloff_y"], a["peeloff_z"] = [float(v) for v in p.peeloff_origin]
        a["d_min"], a["d_max"] = float(p.d_min), float(p.d_max)
    a["compute_image"] = _yn(p.image is not None)
    if p.image is not None:
        a["n_x"], a["n_y"] = np.int64(p.image[0]), np.int64(p.image[1])
        a["x_min"], a["x_max"], a["y_min"], a["y_max"] = [float(v) for v in p.image[2:]]
    a["compute_sed"] = _yn(p.sed is not None)
    if p.sed is not None:
        a["n_ap"] = np.int64(p.sed[0])
        a["ap_min"], a["ap_max"] = float(p.sed[1]), float(p.sed[2])
    if p.inu_min > 0:
        # _write_wavelength_index_range (hyperion/conf/conf_files.py:1076-1079)
        a["n_wav"] = np.int64(p.inu_max - p.inu_min + 1)
        a["inu_min"], a["inu_max"] = np.int64(p.inu_min), np.int64(p.inu_max)
    elif p.filters:
        # _write_filters (hyperion/conf/conf_files.py:874-881) + Filter.to_hdf5_group (hyperion/filter/filter.py:90-126)
        a["use_filters"] = b"yes"
        a["n_filt"] = np.int64(len(p.filters))
        for i, (nu, tn, nu0) in enumerate(p.filters):
            d = g.create_dataset("filter_%05i" % (i + 1), _table([("nu", np.asarray(nu, dtype=np.float64)),
                                                                   ("tr", np.asarray(tn, dtype=np.float64)),
                                                                   ("tn", np.asarray(tn, dtype=np.float64))]))
            d.attrs["nu0"] = float(nu0)
    else:
        a["use_filters"] = b"no"
        a["n_wav"] = np.int64(p.wavelengths[0])
        a["wav_min"], a["wav_max"] = float(p.wavelengths[1]), float(p.wavelengths[2])
    a["track_origin"] = p.track_origin
    a["track_n_scat"] = np.int64(p.track_n_scat)
    a["uncertainties"] = _yn(p.uncertainties)
    a["compute_stokes"] = _yn(p.stokes)
    a["io_bytes"] = np.int64(p.io_bytes)


def write_rtin(filename, model: FlatModel, n_initial_iter=5, n_initial_photons=10000, n_last_photons=0,
               output_specific_energy="last", copy_input=True, check_convergence=None, physics_io_bytes=8,
               raytracing=False, n_ray_photons=(0, 0), extra_root_attrs=None, n_last_photons_mono=(0, 0),
               output_n_photons="none", output_specific_energy_spectrum="none"):
    f = h5write.File()
    c = model.conf
    A = f.attrs
    A["python_version"] = "0.9.12"
    mono = model.frequencies is not None
    A["monochromatic"] = _yn(mono)
    if mono:
        # Model._write_monochromatic (hyperion/model/model.py:133-137), RunConf n_photons (conf_files.py:260-268)
        f.create_dataset("frequencies", _table([("nu", np.asarray(model.frequencies, dtype=np.float64))]))
        A["monochromatic_energy_threshold"] = float(model.monochromatic_energy_threshold)
        A["n_last_photons_sources"] = float(n_last_photons_mono[0])
        A["n_last_photons_dust"] = float(n_last_photons_mono[1])
    A["raytracing"] = _yn(raytracing)
    A["n_stats"] = np.int64(0)
    A["n_inter_max"] = np.int64(c.n_inter_max)
    A["n_reabs_max"] = np.int64(c.n_reabs_max)
    A["pda"] = _yn(c.use_pda)
    A["mrw"] = _yn(c.use_mrw)
    if c.use_mrw:
        A["mrw_gamma"] = float(c.mrw_gamma)
        A["n_inter_mrw_max"] = np.int64(c.n_mrw_max)
    A["kill_on_absorb"] = _yn(c.kill_on_absorb)
    A["kill_on_scatter"] = _yn(c.kill_on_scatter)
    A["forced_first_interaction"] = _yn(c.forced_first_interaction)
    A["forced_first_interaction_algorithm"] = c.forced_first_interaction_algorithm
    A["forced_first_interaction_baes16_xi"] = float(c.baes16_xi)
    A["propagation_check_frequency"] = float(c.propagation_check_frequency)
    A["sample_sources_evenly"] = _yn(c.sample_sources_evenly)
    A["enforce_energy_range"] = _yn(c.enforce_energy_range)
    A["seed"] = np.int32(c.seed)
    A["n_initial_iter"] = np.int64(n_initial_iter)
    A["n_initial_photons"] = float(n_initial_photons)     # the front end stores what the user passed
    if not mono:
        A["n_last_photons"] = float(n_last_photons)
    if raytracing:
        A["n_ray_photons_sources"] = float(n_ray_photons[0])
        A["n_ray_photons_dust"] = float(n_ray_photons[1])
    A["specific_energy_type"] = "additional" if c.specific_energy_additional else "initial"
    A["physics_io_bytes"] = np.int32(physics_io_bytes)
    A["copy_input"] = _yn(copy_input)
    if check_convergence is None:
        A["check_convergence"] = b"no"
    else:
        A["check_convergence"] = b"yes"
        A["convergence_absolute"], A["convergence_relative"], A["convergence_percentile"] = \
            [float(x) for x in check_convergence]
    for k, v in (extra_root_attrs or {}).items():
        A[k] = v

    geo = f.create_group("Grid/Geometry")
    h = hashlib.md5()
    if model.grid_type == "amr":
        h.update(np.array([v for lev in model.amr_levels for g in lev for v in g], dtype=np.float64).tobytes())
    elif model.grid_type == "oct":
        # hyperion/grid/octree_grid.py:426-436
        h.update(np.ascontiguousarray(model.refined).tobytes())
        h.update(np.array(tuple(model.oct_center) + tuple(model.oct_half)).tobytes())
    elif model.grid_type == "vor":
        # hyperion/grid/voronoi_grid.py (get_geometry_id): the sites
        h.update(np.ascontiguousarray(model.voronoi["coordinates"]).tobytes())
    else:
        for w in (model.w1, model.w2, model.w3):
            h.update(np.ascontiguousarray(w).tobytes())
    gid = h.hexdigest()
    geo.attrs["geometry"] = gid
    if model.grid_type == "oct":
        geo.attrs["grid_type"] = "oct"
        for k, v in zip(("x", "y", "z", "dx", "dy", "dz"), tuple(model.oct_center) + tuple(model.oct_half)):
            geo.attrs[k] = float(v)
        geo.create_dataset("cells", _table([("refined", np.asarray(model.refined, dtype=np.int32))]))
    if model.grid_type == "vor":
        # hyperion/grid/voronoi_grid.py:417-478
        v = model.voronoi
        geo.attrs["grid_type"] = "vor"
        for k, x in zip(("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"), v["box"]):
            geo.attrs[k] = float(x)
        vol = np.array(v["volume"], dtype=np.float64)
        vol[~(vol > 0.) | ~np.isfinite(vol)] = -1.
        geo.create_dataset("cells", _table([("coordinates", v["coordinates"]), ("volume", vol),
                                            ("bb_min", v["bb_min"]), ("bb_max", v["bb_max"])]))
        geo.create_dataset("sparse_neighs", np.asarray(v["sparse_neighs"], dtype=np.int32))
        geo.create_dataset("sparse_idx", np.asarray(v["sparse_idx"], dtype=np.int32))
    # hyperion/grid/cartesian_grid.py:336-343, hyperion/grid/spherical_polar_grid.py (write)
    if model.grid_type == "amr":
        # hyperion/grid/amr_grid.py:372-412
        geo.attrs["grid_type"] = "amr"
        geo.attrs["nlevels"] = np.int64(len(model.amr_levels))
        for il, lev in enumerate(model.amr_levels):
            gl = geo.create_group("level_%05i" % (il + 1))
            gl.attrs["ngrids"] = np.int64(len(lev))
            for ig, g in enumerate(lev):
                gg = gl.create_group("grid_%05i" % (ig + 1))
                for k, v in zip(("n1", "n2", "n3"), g[:3]):
                    gg.attrs[k] = np.int64(v)
                for k, v in zip(("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"), g[3:]):
                    gg.attrs[k] = float(v)
    elif model.grid_type not in ("oct", "vor"):
        cols = {"sph": ("r", "t", "p"), "cyl": ("w", "z", "p"), "car": ("x", "y", "z")}[model.grid_type]
        geo.attrs["grid_type"] = {"sph": "sph_pol", "cyl": "cyl_pol", "car": "car"}[model.grid_type]
        geo.create_dataset("walls_1", _table([(cols[0], model.w1)]))
        geo.create_dataset("walls_2", _table([(cols[1], model.w2)]))
        geo.create_dataset("walls_3", _table([(cols[2], model.w3)]))
    q = f.create_group("Grid/Quantities")
    if model.grid_type == "amr":
        for il, ig, sl, shp in model.amr_slices():
            qg = q.create_group("level_%05i/grid_%05i" % (il + 1, ig + 1))
            qg.create_dataset("density", model.density[:, sl].reshape((-1,) + shp))
            if model.specific_energy is not None:
                qg.create_dataset("specific_energy", np.asarray(model.specific_energy)[:, sl].reshape((-1,) + shp))
    else:
        d = q.create_dataset("density", model.density)
        d.attrs["geometry"] = gid
        if model.specific_energy is not None:
            d = q.create_dataset("specific_energy", model.specific_energy)
            d.attrs["geometry"] = gid
    if model.minimum_specific_energy is not None:
        q.attrs["minimum_specific_energy"] = np.asarray(model.minimum_specific_energy, dtype=np.float64)

    gd = f.create_group("Dust")
    for i, dust in enumerate(model.dust):
        name = "dust_%03i" % (i + 1)
        first = next(j for j, other in enumerate(model.dust) if other is dust)
        if first < i:
            gd[name] = h5write.SoftLink("/Dust/dust_%03i" % (first + 1))   # model.py:654-660
        else:
            write_dust(gd.create_group(name), dust)

    gs = f.create_group("Sources")
    for i, s in enumerate(model.sources):
        g = gs.create_group("source_%05i" % (i + 1))
        # hyperion/sources/source.py: write() of each source class
        stype = {1: "point", 2: "sphere", 4: "map", 5: "extern_sph", 6: "extern_box", 7: "plane_parallel",
                 8: "point_collection"}[s.type]
        g.attrs["type"] = stype
        g.attrs["peeloff"] = _yn(s.peeloff)
        if stype == "point_collection":
            g.create_dataset("position", np.ascontiguousarray(s.points, dtype=np.float64).reshape(-1, 3))
            g.create_dataset("luminosity", np.ascontiguousarray(s.points_luminosity, dtype=np.float64))
        else:
            g.attrs["luminosity"] = float(s.luminosity)
        if stype == "map" and model.grid_type == "amr":
            flat = np.ascontiguousarray(s.map, dtype=np.float64).reshape(-1)
            for il, ig, sl, shp in model.amr_slices():
                g.create_dataset("level_%05i/grid_%05i/Luminosity map" % (il + 1, ig + 1), flat[sl].reshape(shp))
        elif stype == "map":
            g.create_dataset("Luminosity map", np.ascontiguousarray(s.map, dtype=np.float64))
        if stype in ("point", "sphere", "extern_sph", "plane_parallel"):
            g.attrs["x"], g.attrs["y"], g.attrs["z"] = [float(v) for v in s.position]
        if stype in ("sphere", "extern_sph", "plane_parallel"):
            g.attrs["r"] = float(s.radius)
        if stype == "sphere":
            g.attrs["limb"] = _yn(s.limb_darkening)
            for k, d in enumerate(s.spots or []):
                # SpotSource.write (hyperion/sources/source.py:355-365)
                gs2 = g.create_group("Spot %i" % k)
                gs2.attrs["type"] = "spot"
                for key in ("luminosity", "longitude", "latitude", "radius"):
                    gs2.attrs[key] = float(d[key])
                gs2.attrs["peeloff"] = _yn(True)
                if d.get("temperature") is not None:
                    gs2.attrs["spectrum"] = "temperature"
                    gs2.attrs["temperature"] = float(d["temperature"])
                else:
                    gs2.attrs["spectrum"] = "spectrum"
                    gs2.create_dataset("spectrum", _table([("nu", d["spectrum_nu"]), ("fnu", d["spectrum_fnu"])]))
        if stype == "extern_box":
            for k, v in zip(("xmin", "xmax", "ymin", "ymax", "zmin", "zmax"), s.bounds):
                g.attrs[k] = float(v)
        if stype == "plane_parallel":
            g.attrs["theta"], g.attrs["phi"] = float(s.direction[0]), float(s.direction[1])
        if s.lte:
            g.attrs["spectrum"] = "lte"
        elif s.temperature is not None:
            g.attrs["spectrum"] = "temperature"
            g.attrs["temperature"] = float(s.temperature)
        else:
            g.attrs["spectrum"] = "spectrum"
            g.create_dataset("spectrum", _table([("nu", s.spectrum_nu), ("fnu", s.spectrum_fnu)]))

    go = f.create_group("Output")
    go.attrs["output_density"] = "none"
    go.attrs["output_density_diff"] = "none"
    go.attrs["output_specific_energy"] = output_specific_energy
    go.attrs["output_n_photons"] = output_n_photons
    if output_specific_energy_spectrum != "none" or model.spectrum_bin_edges is not None:
        # hyperion/model/model.py (set_specific_energy_spectrum_bins / write): option + table of bin edges
        go.attrs["output_specific_energy_spectrum"] = output_specific_energy_spectrum
    if model.spectrum_bin_edges is not None:
        f.create_dataset("specific_energy_spectrum_bin_edges",
                         _table([("nu", np.asarray(model.spectrum_bin_edges, dtype=np.float64))]))
    gb = go.create_group("Binned")
    if model.binned is not None:
        write_peeled_group(gb.create_group("group_00001"), model.binned)
    gp = go.create_group("Peeled")
    for i, p in enumerate(model.peeled):
        write_peeled_group(gp.create_group("group_%05i" % (i + 1)), p)
    f.write(filename)
    return gid
