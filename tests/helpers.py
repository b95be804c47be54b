"""Shared model builders for the tests."""
import os

import numpy as np

from hyperion_b200.flatmodel import FlatConf, FlatDust, FlatModel, FlatPeeledGroup, FlatSource

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pc = 3.08568025e18     # hyperion/util/constants.py
lsun = 3.846e33


def kmh_dust(z):
    return FlatDust.from_npz_dict(z, "dust_")


def bitlevel_model(z, evenly, multi):
    """hyperion/model/tests/test_bit_level.py::TestBasic::test_specific_energy, grid_type='car'."""
    dust = kmh_dust(z)
    dens = [z["density_1"]] + ([z["density_2"], z["density_3"]] if multi else [])
    srcs = [FlatSource(type=1, luminosity=float(l), temperature=float(t), position=tuple(p))
            for l, t, p in zip(z["source_luminosity"], z["source_temperature"], z["source_position"])]
    return FlatModel(z["w1"], z["w2"], z["w3"], np.array(dens), [dust] * len(dens), srcs,
                     FlatConf(sample_sources_evenly=evenly))


def bitlevel_model_sph(zc, z, evenly, multi, grid_type="sph"):
    """Same test on the spherical polar grid (test_bit_level.py:58-62): r, theta, phi walls (or, with
    grid_type="cyl", the cylindrical polar grid of :52-56: w, z, phi walls); the dust and the five
    point sources are those of the Cartesian fixture ``zc``."""
    dust = kmh_dust(zc)
    dens = [z["density_1"]] + ([z["density_2"], z["density_3"]] if multi else [])
    srcs = [FlatSource(type=1, luminosity=float(l), temperature=float(t), position=tuple(p))
            for l, t, p in zip(zc["source_luminosity"], zc["source_temperature"], zc["source_position"])]
    return FlatModel(z["w1"], z["w2"], z["w3"], np.array(dens), [dust] * len(dens), srcs,
                     FlatConf(sample_sources_evenly=evenly), grid_type=grid_type)


def bitlevel_model_oct(zc, z, evenly, multi):
    """The same test on the octree of test_bit_level.py:93-96 (25 nodes, two refined children)."""
    dust = kmh_dust(zc)
    dens = [z["density_1"]] + ([z["density_2"], z["density_3"]] if multi else [])
    srcs = [FlatSource(type=1, luminosity=float(l), temperature=float(t), position=tuple(p))
            for l, t, p in zip(zc["source_luminosity"], zc["source_temperature"], zc["source_position"])]
    return FlatModel(None, None, None, np.array(dens), [dust] * len(dens), srcs,
                     FlatConf(sample_sources_evenly=evenly), grid_type="oct", refined=z["refined"],
                     oct_center=tuple(z["center"]), oct_half=tuple(z["half"]))


def bitlevel_model_amr(zc, z, evenly, multi):
    """The same test on the two-level AMR grid of test_bit_level.py:64-91."""
    dust = kmh_dust(zc)
    dens = [z["density_1"]] + ([z["density_2"], z["density_3"]] if multi else [])
    srcs = [FlatSource(type=1, luminosity=float(l), temperature=float(t), position=tuple(p))
            for l, t, p in zip(zc["source_luminosity"], zc["source_temperature"], zc["source_position"])]
    levels, k = [], 0
    for n in z["n_grids"]:
        levels.append([(int(g[0]), int(g[1]), int(g[2])) + tuple(float(v) for v in g[3:]) for g in z["levels"][k:k + n]])
        k += n
    return FlatModel(None, None, None, np.array(dens), [dust] * len(dens), srcs,
                     FlatConf(sample_sources_evenly=evenly), grid_type="amr", amr_levels=levels)


def bitlevel_model_vor(zc, evenly, multi, n_sites=160, seed=11):
    """The bit-level test's dust, sources and box on a seeded random Voronoi mesh (the reference has no Voronoi
    fixture: its test_voronoi_basics only checks that a 1000-site run completes).  Densities follow the
    Cartesian fixture's values at the sites' positions."""
    from hyperion_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    box = np.array([zc["w1"][0], zc["w1"][-1], zc["w2"][0], zc["w2"][-1], zc["w3"][0], zc["w3"][-1]])
    sites = np.stack([rng.uniform(box[2 * a], box[2 * a + 1], n_sites) for a in range(3)], axis=1)
    mesh = syn.voronoi_mesh(sites, box)
    idx = [np.clip(np.searchsorted(zc[w], sites[:, a]) - 1, 0, len(zc[w]) - 2) for a, w in enumerate(("w1", "w2", "w3"))]
    dust = kmh_dust(zc)
    names = ["density_1"] + (["density_2", "density_3"] if multi else [])
    dens = [zc[k][idx[2], idx[1], idx[0]] for k in names]
    srcs = [FlatSource(type=1, luminosity=float(l), temperature=float(t), position=tuple(p))
            for l, t, p in zip(zc["source_luminosity"], zc["source_temperature"], zc["source_position"])]
    return FlatModel(None, None, None, np.array(dens), [dust] * len(dens), srcs,
                     FlatConf(sample_sources_evenly=evenly), grid_type="vor", voronoi=mesh)


def peeloff_model_vor(zc, evenly):
    m = bitlevel_model_vor(zc, evenly, False)
    m.peeled = peeloff_groups()
    return m


def peeloff_model_amr(zc, z, evenly):
    m = bitlevel_model_amr(zc, z, evenly, False)
    m.peeled = peeloff_groups()
    return m


def peeloff_model_oct(zc, z, evenly):
    m = bitlevel_model_oct(zc, z, evenly, False)
    m.peeled = peeloff_groups()
    return m


def ulp_diff(a, b):
    """Distance in units in the last place, as hyperion/model/tests/test_helpers.py:59-144 measures it."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))


def peeloff_groups():
    """The three peeled image groups of test_bit_level.py::TestBasic::test_peeloff (:198-224)."""
    g1 = FlatPeeledGroup(theta=[33.4, 110.], phi=[65.4, 103.2], wavelengths=(5, 0.05, 200.),
                         image=(4, 5, -0.8 * pc, 0.8 * pc, -pc, pc), sed=(5, 0.1 * pc, pc), track_origin="no")
    g2 = FlatPeeledGroup(theta=[22.1], phi=[203.2], wavelengths=(4, 0.05, 200.),
                         image=(6, 6, -pc, pc, -pc, pc), sed=(2, 0.5 * pc, pc), track_origin="basic")
    g3 = FlatPeeledGroup(theta=[22.1], phi=[203.2], wavelengths=(4, 0.05, 200.),
                         image=(6, 6, -pc, pc, -pc, pc), sed=(2, 0.5 * pc, pc), track_origin="detailed")
    return [g1, g2, g3]


def peeloff_model(z, evenly):
    m = bitlevel_model(z, evenly, False)
    m.peeled = peeloff_groups()
    return m


def peeloff_model_sph(zc, z, evenly, grid_type="sph"):
    m = bitlevel_model_sph(zc, z, evenly, False, grid_type)
    m.peeled = peeloff_groups()
    return m
