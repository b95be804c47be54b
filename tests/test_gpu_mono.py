"""GPU parity of the monochromatic final iteration (src/main/iter_final_mono.f90, src/grid/grid_monochromatic.f90):
CUDA engine vs the CPU oracle, whose monochromatic mode is pinned by tests/test_oracle_mono.py.

* blackbody point source in an empty grid: exact known answer (every packet carries normalized_B_nu L / N);
* dusty grid: SEDs and images split by origin (source / dust, direct / scattered) of B independent batches on each
  side, per-bin z-scores; source packets (forced scatterings with albedo weights, energy threshold) and thermal
  packets (cells drawn from the emission probability at the frequency) separately and together with raytracing;
* thin grid: every thermal packet carries the same energy, so both sides agree to the noise in WHICH cells emit.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from helpers import bitlevel_model, bitlevel_model_sph
from hyperion_b200.flatmodel import FlatPeeledGroup

pytestmark = pytest.mark.gpu

H, K, C, SIGMA = 6.6260689633e-27, 1.380650424e-16, 2.99792458e10, 5.670400e-5
pc = 3.08568025e18


def _engine(model):
    from hyperion_b200.capi import Engine
    eng = Engine(0)
    eng.load_model(model)
    return eng


def test_blackbody_point_source_in_vacuum_is_exact(golden_car):
    m = bitlevel_model(golden_car, False, False)
    m.density[...] = 0.0
    m.sources = m.sources[:1]
    T, L = m.sources[0].temperature, m.sources[0].luminosity
    m.frequencies = C / (np.array([0.3, 0.55, 1.2, 3.6, 24., 160.]) * 1e-4)
    m.peeled = [FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(1, 0., 8. * pc), stokes=True,
                                inu_min=1, inu_max=6, wavelengths=(6, 1., 2.))]
    eng = _engine(m)
    eng.final_begin()
    for inu in range(1, 7):
        eng.final_mono_photons(inu, 0, 1000, 1000, 0, 0, 0)
    eng.final_finish()
    sed = eng.sed(0)[0, 0, :, 0, :]
    nu = m.frequencies
    want = nu * (2. * H / C ** 2 / SIGMA * np.pi) * nu ** 3 / np.expm1(H * nu / K / T) / T ** 4 * L
    assert np.allclose(sed[0], want, rtol=1e-9) and np.allclose(sed[1], want, rtol=1e-9)
    assert not eng.sed(0)[1:].any()
    eng.close()


def _dusty(golden_car, golden_sph=None, kind="sph"):
    from oracle import oracle
    from helpers import bitlevel_model_oct, bitlevel_model_amr
    make = {"sph": bitlevel_model_sph, "oct": bitlevel_model_oct, "amr": bitlevel_model_amr}[kind]
    m = make(golden_car, golden_sph, False, False) if golden_sph is not None else \
        bitlevel_model(golden_car, False, False)
    m.density *= 10.0
    o = oracle.Oracle(m)
    for _ in range(5):
        o.run_lucy_iteration(200000)
    m.specific_energy = o.get_specific_energy()
    wav = np.array([0.45, 2.2, 40., 110.])
    m.frequencies = C / (wav * 1e-4)
    half = 2.0 * pc
    m.peeled = [FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(1, 0., 8. * pc),
                                image=(4, 4, -half, half, -half, half), stokes=True, track_origin="basic",
                                inu_min=1, inu_max=4, wavelengths=(4, 1., 2.))]
    return m


def _z(a, b, zmax=5.5, z2max=1.8):
    a, b = np.array(a), np.array(b)
    ma, mb = a.mean(0), b.mean(0)
    den = np.sqrt(a.var(0, ddof=1) / len(a) + b.var(0, ddof=1) / len(b))
    ok = ((a != 0).mean(0) > 0.9) & ((b != 0).mean(0) > 0.9) & (den > 0)
    assert ok.sum() > 8
    z = (ma[ok] - mb[ok]) / den[ok]
    print("bins %d  max |z| %.2f  <z^2> %.2f" % (ok.sum(), np.abs(z).max(), (z ** 2).mean()))
    assert np.abs(z).max() < zmax and (z ** 2).mean() < z2max


@pytest.mark.parametrize("case", ["sources", "thermal", "raytracing", "spherical", "octree", "amr"])
def test_mono_matches_oracle(golden_car, golden_sph, golden_oct, golden_amr, case):
    """octree / amr: the thermal packets come from the cells that hold physical quantities only (leaves, cells not
    covered by a finer grid: geo%mask_map in setup_monochromatic_grid_pdfs)."""
    from oracle import oracle
    other = {"spherical": (golden_sph, "sph"), "octree": (golden_oct, "oct"), "amr": (golden_amr, "amr")}.get(case)
    m = _dusty(golden_car, other[0], other[1]) if other else _dusty(golden_car)
    B = 10
    ns, nd = (30000, 0) if case == "sources" else (0, 30000) if case == "thermal" else (20000, 20000)
    ray = case == "raytracing"

    def gpu(b):
        eng = _engine(m)
        eng.final_begin()
        for inu in range(1, 5):
            eng.final_mono_photons(inu, b * ns, ns, ns, b * nd, nd, nd, ray)
        eng.final_finish()
        if ray:
            eng.raytracing_photons(20000, 20000, first_source_id=b * 20000, first_dust_id=b * 20000)
        out = (eng.sed(0).copy(), eng.image(0).copy())
        eng.close()
        return out

    def orc(b):
        o = oracle.Oracle(m, rank=b)
        o.final_begin()
        for inu in range(1, 5):
            o.final_mono_photons(inu, ns, ns, nd, nd, ray)
        o.final_finish()
        if ray:
            o.raytracing_photons(20000, 20000)
        return o.sed(0).copy(), o.image(0).copy()

    G = [gpu(b) for b in range(B)]
    with ThreadPoolExecutor(max_workers=8) as pool:
        O = list(pool.map(orc, range(B)))
    _z([g[0] for g in G], [o[0] for o in O])
    _z([g[1] for g in G], [o[1] for o in O])
    # the totals per frequency agree to the noise of the totals
    tg = np.array([g[0][0].sum((0, 1, 2)) for g in G])
    to = np.array([o[0][0].sum((0, 1, 2)) for o in O])
    err = np.sqrt(tg.var(0, ddof=1) / B + to.var(0, ddof=1) / B)
    assert np.all(np.abs(tg.mean(0) - to.mean(0)) < 5 * err + 1e-300), (tg.mean(0), to.mean(0), err)


def test_thermal_emission_of_a_thin_grid(golden_car):
    from oracle import oracle
    m = bitlevel_model(golden_car, False, False)
    m.density *= 1e-8
    rng = np.random.default_rng(3)
    d = m.dust[0]
    m.specific_energy = 10. ** rng.uniform(np.log10(d.jnu_var[10]), np.log10(d.jnu_var[60]), m.density.shape)
    m.frequencies = C / (np.array([12., 70., 350.]) * 1e-4)
    m.conf.forced_first_interaction = False
    m.peeled = [FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(1, 0., 8. * pc), stokes=False,
                                inu_min=1, inu_max=3, wavelengths=(3, 1., 2.))]
    n = 200000
    o = oracle.Oracle(m)
    o.final_begin()
    eng = _engine(m)
    eng.final_begin()
    for inu in (1, 2, 3):
        o.final_mono_photons(inu, 0, 0, n, n)
        eng.final_mono_photons(inu, 0, 0, 0, 0, n, n)
    o.final_finish()
    eng.final_finish()
    a, b = eng.sed(0)[0, 0, :, 0, :], o.sed(0)[0, 0, :, 0, :]
    assert np.all(b > 0) and np.allclose(a, b, rtol=8e-3), (a, b)
    eng.close()


def test_errors(golden_car):
    from hyperion_b200.capi import HyperionError
    m = bitlevel_model(golden_car, False, False)
    m.frequencies = C / (np.array([1., 10.]) * 1e-4)
    m.peeled = [FlatPeeledGroup(theta=[30.], phi=[40.], sed=(1, 0., 8. * pc), inu_min=1, inu_max=3, wavelengths=(3, 1., 2.))]
    with pytest.raises(HyperionError, match="inu_max value is out of range"):
        _engine(m)
    m.peeled = [FlatPeeledGroup(theta=[30.], phi=[40.], sed=(1, 0., 8. * pc), inu_min=1, inu_max=2, wavelengths=(2, 1., 2.))]
    eng = _engine(m)
    eng.final_begin()
    with pytest.raises(HyperionError, match="incorrect inu"):
        eng.final_mono_photons(3, 0, 10, 10, 0, 0, 0)
    eng.close()
