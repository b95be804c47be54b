"""GPU parity of the final (imaging) and raytracing iterations: CUDA engine vs the CPU oracle.

The oracle reproduces the reference's test_peeloff golden files (tests/test_oracle_golden.py); the
engine uses per-packet counter RNG streams, so parity is statistical: B independent batches on each
side, per-bin z-scores of the batch means of every SED / image cube.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from helpers import peeloff_model, peeloff_model_amr, peeloff_model_oct, peeloff_model_sph, bitlevel_model, peeloff_groups, pc

pytestmark = pytest.mark.gpu


def _converged_energy(model, n=200000):
    """Specific energy after the reference's 5 Lucy iterations (oracle), used as the common
    starting state of both imaging runs."""
    from oracle import oracle
    o = oracle.Oracle(model)
    for _ in range(5):
        o.run_lucy_iteration(n)
    return o.get_specific_energy()


def _cubes(x, groups):
    out = {}
    for ig, g in enumerate(groups):
        if g.sed is not None:
            out["g%d_sed" % ig] = x.sed(ig)
        if g.image is not None:
            out["g%d_img" % ig] = x.image(ig)
    return out


def _gpu_batch(model, b, n_final, scattering_only, n_ray):
    from hyperion_b200.capi import Engine
    eng = Engine(0)
    eng.load_model(model)
    eng.final_begin()
    eng.final_photons(b * n_final, n_final, scattering_only)
    st = eng.final_finish()
    st2 = None
    if n_ray:
        st2 = eng.raytracing_photons(n_ray[0], n_ray[1], first_source_id=b * n_ray[0], first_dust_id=b * n_ray[1])
    cubes = _cubes(eng, model.peeled)
    eng.close()
    return cubes, st.as_dict(), None if st2 is None else st2.as_dict()


def _oracle_batch(model, b, n_final, scattering_only, n_ray):
    from oracle import oracle
    o = oracle.Oracle(model, rank=b)
    o.final_begin()
    o.final_photons(n_final, scattering_only)
    st = o.final_finish()
    st2 = None
    if n_ray:
        st2 = o.raytracing_photons(n_ray[0], n_ray[1])
    return _cubes(o, model.peeled), st.as_dict(), None if st2 is None else st2.as_dict()


def _run_both(model, B, n_final, scattering_only, n_ray):
    gpu = [_gpu_batch(model, b, n_final, scattering_only, n_ray) for b in range(B)]
    with ThreadPoolExecutor(max_workers=8) as pool:
        orc = list(pool.map(lambda b: _oracle_batch(model, b, n_final, scattering_only, n_ray), range(B)))
    return gpu, orc


def _compare(gpu, orc, zmax=5.5):
    keys = gpu[0][0].keys()
    report = {}
    for k in keys:
        a = np.array([g[0][k] for g in gpu])
        b = np.array([o[0][k] for o in orc])
        ma, mb = a.mean(0), b.mean(0)
        sa, sb = a.std(0, ddof=1) / np.sqrt(len(a)), b.std(0, ddof=1) / np.sqrt(len(b))
        den = np.sqrt(sa ** 2 + sb ** 2)
        # bins both sides populate in most batches (the z statistic assumes near-normal means)
        filled = ((a != 0).mean(0) > 0.9) & ((b != 0).mean(0) > 0.9) & (den > 0)
        assert filled.sum() > 0, k
        z = (ma[filled] - mb[filled]) / den[filled]
        report[k] = (float(np.abs(z).max()), float((z ** 2).mean()), int(filled.sum()))
        assert np.abs(z).max() < zmax, (k, report[k])
        assert (z ** 2).mean() < 1.8, (k, report[k])
    return report


@pytest.mark.parametrize("raytracing", [False, True])
def test_peeloff_matches_oracle(golden_car, raytracing):
    """The reference's test_peeloff model (test_bit_level.py:175-236): three peeled groups (no / basic /
    detailed origin tracking, Stokes on), forced first interaction; with raytracing the imaging
    iteration peels scattered light only and the raytracing iteration adds sources + thermal emission."""
    model = peeloff_model(golden_car, False)
    model.specific_energy = _converged_energy(model)
    B = 12
    gpu, orc = _run_both(model, B, 60000, raytracing, (20000, 30000) if raytracing else None)
    report = _compare(gpu, orc)
    print(report)
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    assert all(g[1]["killed_geo"] == 0 and g[1]["killed_int"] == 0 and g[1]["n_photons"] == 60000 for g in gpu)
    if raytracing:
        for key in ("n_peeloffs", "n_peel_crossings"):
            a = np.mean([g[2][key] for g in gpu])
            b = np.mean([o[2][key] for o in orc])
            assert abs(a / b - 1) < 0.02, (key, a, b)


@pytest.mark.parametrize("raytracing,geometry", [(False, "sph"), (True, "sph"), (True, "cyl")])
def test_peeloff_matches_oracle_spherical_grid(golden_car, golden_sph, golden_cyl, raytracing, geometry):
    """test_peeloff on the reference's spherical polar grid: peel-off rays cross spheres, cones and
    phi planes; the thermal raytracing packets start at random positions of (r, theta, phi) cells.
    Same on the cylindrical polar grid."""
    model = peeloff_model_sph(golden_car, golden_sph if geometry == "sph" else golden_cyl, False, geometry)
    model.specific_energy = _converged_energy(model)
    B = 12
    gpu, orc = _run_both(model, B, 60000, raytracing, (20000, 30000) if raytracing else None)
    print(_compare(gpu, orc))
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    assert all(g[1]["killed_geo"] == 0 and g[1]["n_photons"] == 60000 for g in gpu)


def test_peeloff_matches_oracle_octree(golden_car, golden_oct):
    """test_peeloff on the reference's octree, with raytracing: thermal packets are drawn from the
    leaves only (random_masked_cell) and weighted with the number of leaves."""
    model = peeloff_model_oct(golden_car, golden_oct, False)
    model.specific_energy = _converged_energy(model)
    gpu, orc = _run_both(model, 12, 60000, True, (20000, 30000))
    print(_compare(gpu, orc))
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    for key in ("n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[2][key] for g in gpu])
        b = np.mean([o[2][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)


def test_peeloff_matches_oracle_amr(golden_car, golden_amr):
    """test_peeloff on the reference's AMR grid, with raytracing (thermal packets from uncovered cells)."""
    model = peeloff_model_amr(golden_car, golden_amr, False)
    model.specific_energy = _converged_energy(model)
    gpu, orc = _run_both(model, 12, 60000, True, (20000, 30000))
    print(_compare(gpu, orc))
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    for key in ("n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[2][key] for g in gpu])
        b = np.mean([o[2][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)


def test_peeloff_multiple_dust_and_even_sampling(golden_car):
    """Three dust types, sources sampled evenly, BAES16 forced interaction, uncertainties on."""
    model = bitlevel_model(golden_car, True, True)
    groups = peeloff_groups()
    for g in groups:
        g.uncertainties = True
    groups[1].track_origin = "scatterings"
    groups[1].track_n_scat = 2
    model.peeled = groups
    model.conf.forced_first_interaction_algorithm = "baes16"
    model.specific_energy = _converged_energy(model)
    gpu, orc = _run_both(model, 12, 40000, False, (10000, 20000))
    print(_compare(gpu, orc))


def test_image_flux_is_additive_and_unattenuated_in_vacuum(golden_car):
    """Known answers (hyperion/model/tests/test_image.py:635-677, test_sed.py): with no dust every
    packet reaches the observer, so the SED integrated over frequency bins inside the largest
    aperture equals the luminosity of the sources; an image holds the same flux as the SED when it
    covers the aperture."""
    from hyperion_b200.capi import Engine
    from hyperion_b200.flatmodel import FlatPeeledGroup
    model = bitlevel_model(golden_car, False, False)
    model.density[...] = 0.0
    model.peeled = [FlatPeeledGroup(theta=[45.], phi=[30.], wavelengths=(20, 0.01, 5000.),
                                    image=(8, 8, -2 * pc, 2 * pc, -2 * pc, 2 * pc), sed=(1, 3 * pc, 3 * pc),
                                    stokes=False)]
    eng = Engine(0)
    eng.load_model(model)
    eng.final_begin()
    eng.final_photons(0, 200000, False)
    st = eng.final_finish()
    sed, img = eng.sed(0), eng.image(0)
    eng.close()
    assert st.n_escaped == 200000 and st.n_peeloffs == 200000
    nu_min, nu_max = 2.99792458e10 / (5000. * 1e-4), 2.99792458e10 / (0.01 * 1e-4)
    dnunorm = (nu_max / nu_min) ** (0.5 / 20) - (nu_max / nu_min) ** (-0.5 / 20)
    ltot = sum(s.luminosity for s in model.sources)
    assert abs(sed.sum() * dnunorm / ltot - 1) < 2e-3     # a little flux lies outside 0.01-5000 micron
    assert abs(img.sum() / sed.sum() - 1) < 1e-12


@pytest.mark.parametrize("limb", [False, True])
def test_stellar_surface_peeloff_flux_in_vacuum(golden_car, limb):
    """Known answer (hyperion/model/tests/test_sed.py): a spherical source in an empty grid, seen from any
    direction, shows its full luminosity -- the surface peel-off weights 4 mu (or the limb-darkened
    2 (1.5 mu^2 + mu)) average to one over the visible hemisphere, and peel-offs from the far side are
    blocked by the star itself (source_type.f90:692-707, grid_propagate_3d.f90:410-415)."""
    from hyperion_b200.capi import Engine
    from hyperion_b200.flatmodel import FlatPeeledGroup, FlatSource
    from helpers import lsun
    model = bitlevel_model(golden_car, False, False)
    model.density[...] = 0.0
    model.sources = [FlatSource(type=2, luminosity=lsun, temperature=5000., position=(0.1 * pc, 0.2 * pc, -0.1 * pc),
                                radius=0.05 * pc, limb_darkening=limb)]
    model.peeled = [FlatPeeledGroup(theta=[45., 120.], phi=[30., 200.], wavelengths=(20, 0.01, 5000.),
                                    image=(8, 8, -2 * pc, 2 * pc, -2 * pc, 2 * pc), sed=(1, 3 * pc, 3 * pc), stokes=False)]
    eng = Engine(0)
    eng.load_model(model)
    eng.final_begin()
    N = 400000
    eng.final_photons(0, N, False)
    st = eng.final_finish()
    sed, img = eng.sed(0), eng.image(0)
    eng.close()
    assert st.n_escaped == N
    assert abs(st.n_peeloffs / N - 1) < 0.01          # two views, half of the surface faces each
    nu_min, nu_max = 2.99792458e10 / (5000. * 1e-4), 2.99792458e10 / (0.01 * 1e-4)
    dnunorm = (nu_max / nu_min) ** (0.5 / 20) - (nu_max / nu_min) ** (-0.5 / 20)
    flux = sed.sum(axis=(0, 1, 3, 4)) * dnunorm / lsun
    assert np.all(np.abs(flux - 1) < 0.01), flux
    assert abs(img.sum() / sed.sum() - 1) < 1e-12


def test_peeloff_with_stellar_source_matches_oracle(golden_car):
    """Imaging + raytracing iterations with a star of finite radius in a dusty Cartesian grid: surface
    peel-off, re-emission peel-offs, lines of sight blocked by the star."""
    from hyperion_b200.flatmodel import FlatSource
    from helpers import lsun
    model = peeloff_model(golden_car, False)
    model.density *= 10.
    model.sources = [FlatSource(type=2, luminosity=lsun, temperature=5000., position=(0.05 * pc, -0.03 * pc, 0.02 * pc),
                                radius=0.25 * pc),
                     FlatSource(type=1, luminosity=0.5 * lsun, temperature=7000., position=(-0.5 * pc, 0.4 * pc, -0.3 * pc))]
    model.specific_energy = _converged_energy(model, 100000)
    gpu, orc = _run_both(model, 12, 60000, True, (20000, 30000))
    print(_compare(gpu, orc))
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    for key in ("n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[2][key] for g in gpu])
        b = np.mean([o[2][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)


@pytest.mark.parametrize("multi", [False, True])
def test_binned_images_match_oracle(golden_car, multi):
    """Binned images (images_binned.f90): every packet that escapes in the final iteration is binned by
    its own direction (2 x 3 direction bins here), with detailed origin tracking so that the source /
    dust ids carried in the packet tag are exercised; forced first interaction must be off."""
    from helpers import FlatPeeledGroup
    model = bitlevel_model(golden_car, False, multi)
    model.conf.forced_first_interaction = False
    model.peeled = [peeloff_groups()[1]]
    model.binned = FlatPeeledGroup(binned=True, n_theta=2, n_phi=3, wavelengths=(4, 0.05, 200.),
                                   image=(3, 3, -pc, pc, -pc, pc), sed=(2, 0.5 * pc, 1.8 * pc), track_origin="detailed",
                                   uncertainties=True)
    model.specific_energy = _converged_energy(model)
    B = 12
    groups = model.peeled + [model.binned]

    def cubes(x):
        return _cubes(x, groups)

    from hyperion_b200.capi import Engine
    from oracle import oracle
    gpu, orc = [], []
    for b in range(B):
        eng = Engine(0)
        eng.load_model(model)
        eng.final_begin()
        eng.final_photons(b * 60000, 60000, False)
        st = eng.final_finish()
        gpu.append((cubes(eng), st.as_dict(), None))
        eng.close()

    def one(b):
        o = oracle.Oracle(model, rank=b)
        o.final_begin()
        o.final_photons(60000, False)
        st = o.final_finish()
        return cubes(o), st.as_dict(), None

    with ThreadPoolExecutor(max_workers=8) as pool:
        orc = list(pool.map(one, range(B)))
    report = _compare(gpu, orc)
    print(report)
    assert "g1_sed" in report and "g1_img" in report and report["g1_sed"][2] > 20
    # the binned cubes hold every escaped packet exactly once
    for key in ("n_escaped", "n_scatterings", "n_absorptions"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)


@pytest.mark.parametrize("geometry,raytracing", [("car", False), ("car", True), ("sph", True)])
def test_inside_observer_matches_oracle(golden_car, golden_sph, geometry, raytracing):
    """Inside observers (images_peeled.f90:131-183): peel-offs travel to an observer inside the grid, stop
    there, are weighted by 1 / (4 pi d^2) and land on a longitude / latitude image."""
    from helpers import FlatPeeledGroup
    from hyperion_b200.flatmodel import FlatSource
    model = peeloff_model(golden_car, False) if geometry == "car" else peeloff_model_sph(golden_car, golden_sph, False)
    if geometry == "car":
        # a star on some lines of sight: peel-offs that would cross it before reaching the observer are dropped
        model.sources.append(FlatSource(type=2, luminosity=2.e33, temperature=6000., position=(0.3 * pc, 0.1 * pc, -0.2 * pc),
                                        radius=0.15 * pc))
    inside = FlatPeeledGroup(theta=[70.], phi=[25.], wavelengths=(4, 0.05, 200.), inside_observer=True,
                             peeloff_origin=(-0.21 * pc, 0.33 * pc, 0.12 * pc),
                             image=(8, 4, 360., 0., -90., 90.), sed=(2, 30., 400.), track_origin="basic",
                             d_min=0.05 * pc, d_max=5 * pc)
    model.peeled = [peeloff_groups()[0], inside]
    model.specific_energy = _converged_energy(model)
    B = 12
    gpu, orc = _run_both(model, B, 60000, raytracing, (20000, 30000) if raytracing else None)
    report = _compare(gpu, orc)
    print(report)
    assert report["g1_img"][2] >= 20 and report["g1_sed"][2] >= 8
    for key in ("n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)


def test_filter_convolution_matches_oracle(golden_car):
    """use_filters (image_type.f90:467-476): two overlapping smooth filters instead of a wavelength grid,
    in a peeled group and in the binned group."""
    from helpers import FlatPeeledGroup
    model = peeloff_model(golden_car, False)
    model.conf.forced_first_interaction = False
    nu = np.logspace(12.5, 15.5, 40)
    f1 = (nu, np.exp(-0.5 * ((np.log10(nu) - 14.5) / 0.3) ** 2), 10 ** 14.5)
    f2 = (nu[::-1].copy(), np.exp(-0.5 * ((np.log10(nu[::-1]) - 13.6) / 0.4) ** 2) * 3.0, 10 ** 13.6)   # decreasing nu
    g = peeloff_groups()[1]
    g.filters = [f1, f2]
    model.peeled = [peeloff_groups()[0], g]
    model.binned = FlatPeeledGroup(binned=True, n_theta=1, n_phi=2, filters=[f2, f1], sed=(1, 0.5 * pc, 1.8 * pc),
                                   track_origin="no", stokes=False)
    model.specific_energy = _converged_energy(model)
    B = 12
    groups = model.peeled + [model.binned]
    from hyperion_b200.capi import Engine
    from oracle import oracle
    gpu = []
    for b in range(B):
        eng = Engine(0)
        eng.load_model(model)
        eng.final_begin()
        eng.final_photons(b * 60000, 60000, False)
        st = eng.final_finish()
        gpu.append((_cubes(eng, groups), st.as_dict(), None))
        eng.close()

    def one(b):
        o = oracle.Oracle(model, rank=b)
        o.final_begin()
        o.final_photons(60000, False)
        st = o.final_finish()
        return _cubes(o, groups), st.as_dict(), None

    with ThreadPoolExecutor(max_workers=8) as pool:
        orc = list(pool.map(one, range(B)))
    report = _compare(gpu, orc)
    print(report)
    assert gpu[0][0]["g1_sed"].shape[-1] == 2 and gpu[0][0]["g2_sed"].shape[-1] == 2
    # raytracing refuses filters with the reference's message
    from hyperion_b200.capi import HyperionError
    eng = Engine(0)
    eng.load_model(model)
    with pytest.raises(HyperionError, match="filter convolution cannot be used with raytracing"):
        eng.raytracing_photons(100, 100)
    eng.close()
