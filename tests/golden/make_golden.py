"""Regenerate the committed golden fixtures from the reference checkout.

Run in the build container (needs /root/reference; never at test time):

    python tests/golden/make_golden.py

Writes tests/golden/bitlevel_car.npz holding
  * the inputs of hyperion/model/tests/test_bit_level.py::TestBasic::test_specific_energy
    for grid_type='car' (walls, the three random density grids, five random point
    sources -- regenerated with the same numpy legacy seeds 141412 / 12345 the
    reference test uses, test_bit_level.py:37-42,141-155),
  * the dust tables of hyperion/model/tests/data/kmh_lite.hdf5,
  * the reference's own outputs: specific_energy of all 5 Lucy iterations from the
    four golden files test_specific_energy.grid_type=car.*.rtout.
"""
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from hyperion_b200.io import h5min  # noqa: E402
from hyperion_b200.flatmodel import FlatDust  # noqa: E402

REF = "/root/reference"
DATA = os.path.join(REF, "hyperion/model/tests/data")
const = runpy.run_path(os.path.join(REF, "hyperion/util/constants.py"))
pc, lsun = const["pc"], const["lsun"]


def setup_all_grid_types(u, d):
    """test_bit_level.py:37-113 (only the random-number consumption order matters)."""
    np.random.seed(141412)
    out = {}
    # AMR level 1 and 2 quantities come first
    for i, shape in enumerate([(4, 6, 8)] * 3 + [(20, 6, 4)] * 3):
        out[("amr", i)] = np.random.random(shape) * d
    shapes = {"car": (3, 5, 7), "cyl": (5, 3, 7), "sph": (3, 7, 5), "oct": (25,)}
    for k in ("density", "density_2", "density_3"):
        for g in ("car", "cyl", "sph"):
            out[(k, g)] = np.random.random(shapes[g]) * d
        out[(k, "oct")] = np.random.random(25) * d
    return out


def sources():
    """test_bit_level.py:141-155"""
    np.random.seed(12345)
    src = []
    for i in range(5):
        lum = np.random.random() * lsun
        temp = np.random.uniform(2000., 10000.)
        pos = np.random.uniform(-pc, pc, 3)
        src.append((lum, temp, pos))
    return src


def main():
    dens = setup_all_grid_types(pc, 1.e-20)
    out = {
        "w1": np.linspace(-pc, pc, 8), "w2": np.linspace(-pc, pc, 6), "w3": np.linspace(-pc, pc, 4),
        "density_1": dens[("density", "car")], "density_2": dens[("density_2", "car")],
        "density_3": dens[("density_3", "car")],
    }
    src = sources()
    out["source_luminosity"] = np.array([s[0] for s in src])
    out["source_temperature"] = np.array([s[1] for s in src])
    out["source_position"] = np.array([s[2] for s in src])
    dust = FlatDust.from_hdf5_group(h5min.File(os.path.join(DATA, "kmh_lite.hdf5")))
    out.update(dust.to_npz_dict("dust_"))
    for evenly in (False, True):
        for multi in (False, True):
            fn = ("test_specific_energy.grid_type=car.sample_sources_evenly=%s."
                  "multiple_densities=%s.rtout" % (evenly, multi))
            f = h5min.File(os.path.join(DATA, fn))
            se = np.array([f["iteration_%05d/specific_energy" % i][...] for i in range(1, 6)])
            out["expected_evenly=%s_multi=%s" % (evenly, multi)] = se
            assert int(f.attrs["iterations"]) == 5
    # test_peeloff (test_bit_level.py:175-236): seds + images of the three peeled groups
    for ray in (False, True):
        for evenly in (False, True):
            fn = ("test_peeloff.grid_type=car.raytracing=%s.sample_sources_evenly=%s.rtout" % (ray, evenly))
            f = h5min.File(os.path.join(DATA, fn))
            for ig in (1, 2, 3):
                for kind in ("seds", "images"):
                    d = f["Peeled/group_%05d/%s" % (ig, kind)]
                    out["peeloff_ray=%s_evenly=%s_g%d_%s" % (ray, evenly, ig, kind)] = d[...]
            for k in ("killed_photons_geo_final", "killed_photons_int_final"):
                assert int(np.asarray(f.attrs[k]).ravel()[0]) == 0
    np.savez_compressed(os.path.join(HERE, "bitlevel_car.npz"), **out)
    print("wrote", os.path.join(HERE, "bitlevel_car.npz"))
    # spherical polar grid of the same tests (test_bit_level.py:58-62): the dust and sources are shared
    # with bitlevel_car.npz, only the geometry, densities and expected outputs are stored
    sph = {"w1": np.linspace(0., 3. * pc, 6), "w2": np.linspace(0., np.pi, 8), "w3": np.linspace(0., 2. * np.pi, 4),
           "density_1": dens[("density", "sph")], "density_2": dens[("density_2", "sph")],
           "density_3": dens[("density_3", "sph")]}
    golden_outputs(sph, "sph")
    np.savez_compressed(os.path.join(HERE, "bitlevel_sph.npz"), **sph)
    print("wrote", os.path.join(HERE, "bitlevel_sph.npz"))
    # cylindrical polar grid (test_bit_level.py:52-56)
    cyl = {"w1": np.linspace(0., 2. * pc, 8), "w2": np.linspace(-pc, pc, 4), "w3": np.linspace(0., 2. * np.pi, 6),
           "density_1": dens[("density", "cyl")], "density_2": dens[("density_2", "cyl")],
           "density_3": dens[("density_3", "cyl")]}
    golden_outputs(cyl, "cyl")
    np.savez_compressed(os.path.join(HERE, "bitlevel_cyl.npz"), **cyl)
    print("wrote", os.path.join(HERE, "bitlevel_cyl.npz"))
    # octree (test_bit_level.py:93-96): 25 nodes, root cell centred on the origin with half-width pc
    refined = [1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0]
    oct = {"refined": np.array(refined, dtype=np.int32), "center": np.zeros(3), "half": np.array([pc, pc, pc]),
           "density_1": dens[("density", "oct")], "density_2": dens[("density_2", "oct")],
           "density_3": dens[("density_3", "oct")]}
    golden_outputs(oct, "oct")
    np.savez_compressed(os.path.join(HERE, "bitlevel_oct.npz"), **oct)
    print("wrote", os.path.join(HERE, "bitlevel_oct.npz"))
    # AMR (test_bit_level.py:64-91): level 1 = one 8x6x4 grid over [-pc, pc]^3, level 2 = one 4x6x20 grid over
    # [-pc, 0]^3; cells concatenated level-major, x fastest (type_cell_id_amr.f90:115-133)
    amr = {"levels": np.array([[8, 6, 4, -pc, pc, -pc, pc, -pc, pc], [4, 6, 20, -pc, 0., -pc, 0., -pc, 0.]]),
           "n_grids": np.array([1, 1])}
    for k in range(3):
        amr["density_%d" % (k + 1)] = np.hstack([dens[("amr", k)].ravel(), dens[("amr", 3 + k)].ravel()])
    golden_outputs(amr, "amr")
    np.savez_compressed(os.path.join(HERE, "bitlevel_amr.npz"), **amr)
    print("wrote", os.path.join(HERE, "bitlevel_amr.npz"))


def golden_outputs(out, grid_type):
    for evenly in (False, True):
        for multi in (False, True):
            fn = ("test_specific_energy.grid_type=%s.sample_sources_evenly=%s."
                  "multiple_densities=%s.rtout" % (grid_type, evenly, multi))
            f = h5min.File(os.path.join(DATA, fn))
            if grid_type == "amr":
                # one dataset per level / grid: flatten in cell-id order
                se = []
                for i in range(1, 6):
                    parts = [f["iteration_%05d/level_%05d/grid_00001/specific_energy" % (i, lev)][...] for lev in (1, 2)]
                    se.append(np.concatenate([p.reshape(p.shape[0], -1) for p in parts], axis=1))
                out["expected_evenly=%s_multi=%s" % (evenly, multi)] = np.array(se)
                continue
            out["expected_evenly=%s_multi=%s" % (evenly, multi)] = \
                np.array([f["iteration_%05d/specific_energy" % i][...] for i in range(1, 6)])
    for ray in (False, True):
        for evenly in (False, True):
            fn = ("test_peeloff.grid_type=%s.raytracing=%s.sample_sources_evenly=%s.rtout" % (grid_type, ray, evenly))
            f = h5min.File(os.path.join(DATA, fn))
            for ig in (1, 2, 3):
                for kind in ("seds", "images"):
                    out["peeloff_ray=%s_evenly=%s_g%d_%s" % (ray, evenly, ig, kind)] = \
                        f["Peeled/group_%05d/%s" % (ig, kind)][...]


if __name__ == "__main__":
    main()
