"""GPU parity for the source types beyond points and spheres: external sphere / box (emitting
inwards), plane-parallel beam, point collection (src/sources/source_type.f90:570-980).  The oracle is
pinned for these emitters by closed-form track densities (tests/test_oracle_sources.py); here the CUDA
engine is compared with it statistically, for the Lucy deposits and for the peeled SEDs / images
(external sources are peeled with the 4 mu weight of emit_from_extern_*_peeloff)."""
import numpy as np
import pytest

from helpers import bitlevel_model, peeloff_groups, pc, lsun
from hyperion_b200.flatmodel import FlatSource
from test_gpu_parity import _gpu_batches, _oracle_batches, _zscores
from test_gpu_imaging import _run_both, _compare, _converged_energy

pytestmark = pytest.mark.gpu


def mixed_sources():
    pts = np.array([[-0.5 * pc, -0.5 * pc, -0.5 * pc], [0.5 * pc, 0.25 * pc, 0.5 * pc], [0.0, 0.5 * pc, -0.25 * pc],
                    [0.9 * pc, -0.9 * pc, 0.1 * pc]])
    return [FlatSource(type=1, luminosity=1.0 * lsun, temperature=6000., position=(0.1 * pc, 0., -0.2 * pc)),
            FlatSource(type=5, luminosity=2.0 * lsun, temperature=5000., position=(0.02 * pc, -0.01 * pc, 0.), radius=0.95 * pc),
            FlatSource(type=6, luminosity=3.0 * lsun, temperature=4000.,
                       bounds=(-0.9 * pc, 0.8 * pc, -0.7 * pc, 0.95 * pc, -0.6 * pc, 0.9 * pc)),
            FlatSource(type=7, luminosity=1.5 * lsun, temperature=7000., position=(-0.3 * pc, 0.2 * pc, -0.8 * pc),
                       radius=0.3 * pc, direction=(25.0, 40.0), peeloff=False),
            FlatSource(type=8, temperature=3000., points=pts, points_luminosity=np.array([1.0, 3.0, 2.0, 0.5]) * lsun),
            FlatSource(type=4, luminosity=2.5 * lsun, temperature=5500., map=_luminosity_map()),
            FlatSource(type=4, luminosity=1.5 * lsun, lte=True, map=_luminosity_map()[::-1].copy())]


def _luminosity_map():
    """A lumpy map on the 7 x 5 x 3 cells of the reference's bit-level grid ([z, y, x]), with dark cells."""
    rng = np.random.default_rng(11)
    m = rng.random((3, 5, 7)) ** 3
    m[1, 2, :3] = 0.0
    return m


@pytest.mark.parametrize("evenly,multi", [(False, False), (True, True)])
def test_deposits_match_oracle_mixed_sources(golden_car, evenly, multi):
    model = bitlevel_model(golden_car, evenly, multi)
    model.sources = mixed_sources()
    B, N = 16, 100000
    g, gst = _gpu_batches(model, N, B)
    o, ost = _oracle_batches(model, N, B)
    z, ok = _zscores(g, o)
    assert ok.all()
    assert np.abs(z).max() < 5.0, np.abs(z).max()
    assert 0.6 < (z ** 2).mean() < 1.5, (z ** 2).mean()
    rel = np.abs(g.mean(0) / o.mean(0) - 1)
    assert np.median(rel) < 0.03
    for key in ("n_crossings", "n_absorptions", "n_scatterings"):
        a = np.mean([s[key] for s in gst])
        b = np.mean([s[key] for s in ost])
        assert abs(a / b - 1) < 0.01, (key, a, b)
    assert all(s["killed_geo"] == 0 and s["killed_int"] == 0 and s["n_photons"] == N for s in gst)


@pytest.mark.parametrize("kind", [4, 5, 6, 7, 8])
def test_each_source_type_alone(golden_car, kind):
    """One source of each new type on its own, so that a wrong emitter cannot hide behind the others."""
    model = bitlevel_model(golden_car, False, False)
    model.sources = [s for s in mixed_sources() if s.type == kind]
    B, N = 12, 100000
    g, gst = _gpu_batches(model, N, B)
    o, ost = _oracle_batches(model, N, B)
    z, ok = _zscores(g, o)
    filled = ok & (o.mean(0) > 0)
    assert filled.sum() > 20
    assert np.abs(z[filled]).max() < 5.0, np.abs(z[filled]).max()
    assert 0.5 < (z[filled] ** 2).mean() < 1.6, (z[filled] ** 2).mean()
    for key in ("n_crossings", "n_absorptions", "n_scatterings"):
        a = np.mean([s[key] for s in gst])
        b = np.mean([s[key] for s in ost])
        assert abs(a / b - 1) < 0.015, (key, a, b)


@pytest.mark.parametrize("raytracing", [False, True])
def test_peeloff_matches_oracle_mixed_sources(golden_car, raytracing):
    model = bitlevel_model(golden_car, False, False)
    model.sources = mixed_sources()
    model.peeled = peeloff_groups()
    model.specific_energy = _converged_energy(model)
    B = 12
    gpu, orc = _run_both(model, B, 60000, raytracing, (20000, 30000) if raytracing else None)
    report = _compare(gpu, orc)
    print(report)
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)


def test_source_argument_errors():
    from hyperion_b200.capi import Engine, HyperionError
    from hyperion_b200 import synthetic as syn
    model = syn.cartesian_point_source_model(n=8, dust=syn.grey_dust())
    for bad, msg in [(FlatSource(type=7, luminosity=lsun, temperature=5000., radius=0.1 * pc, peeloff=True),
                      "Cannot peeloff plane parallel source"),
                     (FlatSource(type=4, luminosity=lsun, temperature=5000.), "luminosity map should have one entry per cell"),
                     (FlatSource(type=4, luminosity=lsun, temperature=5000., map=np.zeros((8, 8, 8))), "all PDF elements are zero"),
                     (FlatSource(type=3, luminosity=lsun, temperature=5000., radius=pc), "unknown type in source list"),
                     (FlatSource(type=1, luminosity=lsun, lte=True), "Point source cannot have LTE spectrum"),
                     (FlatSource(type=6, luminosity=lsun, temperature=5000., bounds=(1., -1., 0., 1., 0., 1.)),
                      "bounds should be increasing"),
                     (FlatSource(type=8, temperature=5000., points=np.zeros((2, 3)), points_luminosity=np.zeros(2)),
                      "all PDF elements are zero")]:
        model.sources = [bad]
        eng = Engine(0)
        with pytest.raises(HyperionError, match=msg):
            eng.load_model(model)
        eng.close()


def _spotted_star_model(golden_car):
    model = bitlevel_model(golden_car, False, False)
    model.density *= 10.
    nu = np.logspace(13.5, 15.2, 12)
    model.sources = [FlatSource(type=2, luminosity=lsun, temperature=5000., position=(0.05 * pc, -0.03 * pc, 0.02 * pc),
                                radius=0.25 * pc,
                                spots=[dict(luminosity=0.8 * lsun, longitude=60., latitude=20., radius=35., temperature=9000.),
                                       dict(luminosity=0.5 * lsun, longitude=140., latitude=250., radius=20.,
                                            spectrum_nu=nu, spectrum_fnu=nu ** -1.5)]),
                     FlatSource(type=1, luminosity=0.2 * lsun, temperature=4000., position=(-0.5 * pc, 0.4 * pc, -0.3 * pc))]
    return model


def test_spotted_star_deposits_match_oracle(golden_car):
    """Source type 3 (source_type.f90:150-188, 421-427, 632-637): spots with their own luminosity and
    spectrum (blackbody, and a table sampled with sample_pdf_log) on a star that also re-absorbs."""
    model = _spotted_star_model(golden_car)
    B, N = 16, 100000
    g, gst = _gpu_batches(model, N, B)
    o, ost = _oracle_batches(model, N, B)
    z, ok = _zscores(g, o)
    assert ok.mean() > 0.9
    assert np.abs(z[ok]).max() < 5.5, np.abs(z[ok]).max()
    assert 0.6 < (z[ok] ** 2).mean() < 1.5, (z[ok] ** 2).mean()
    for key in ("n_crossings", "n_absorptions", "n_scatterings"):
        a = np.mean([s[key] for s in gst])
        b = np.mean([s[key] for s in ost])
        assert abs(a / b - 1) < 0.01, (key, a, b)


def test_spotted_star_peeloff_matches_oracle(golden_car):
    model = _spotted_star_model(golden_car)
    model.peeled = peeloff_groups()
    model.specific_energy = _converged_energy(model)
    gpu, orc = _run_both(model, 12, 60000, True, (20000, 30000))
    report = _compare(gpu, orc)
    print(report)


@pytest.mark.parametrize("geometry", ["sph", "cyl", "oct", "amr"])
def test_map_source_on_other_grids(golden_car, golden_sph, golden_cyl, golden_oct, golden_amr, geometry):
    """emit_from_map uses random_position_cell of each geometry module; an LTE and a blackbody map."""
    from helpers import bitlevel_model_sph, bitlevel_model_oct, bitlevel_model_amr
    if geometry in ("sph", "cyl"):
        model = bitlevel_model_sph(golden_car, golden_sph if geometry == "sph" else golden_cyl, False, True, geometry)
    elif geometry == "oct":
        model = bitlevel_model_oct(golden_car, golden_oct, False, True)
    else:
        model = bitlevel_model_amr(golden_car, golden_amr, False, True)
    rng = np.random.default_rng(5)
    shape = model.density.shape[1:]
    # refined octree nodes / covered AMR cells hold no dust: an LTE source cannot emit from them (the
    # reference stops in find_cdf, as the oracle does)
    from test_gpu_parity import _engine
    eng = _engine(model)
    dusty = (eng.get_density().reshape(model.density.shape) > 0).all(axis=0)
    eng.close()
    assert 0 < dusty.sum() < dusty.size or geometry in ("sph", "cyl")
    model.sources = [FlatSource(type=4, luminosity=2 * lsun, temperature=6000., map=rng.random(shape) ** 2),
                     FlatSource(type=4, luminosity=lsun, lte=True, map=rng.random(shape) * dusty)]
    B, N = 12, 100000
    g, gst = _gpu_batches(model, N, B)
    o, ost = _oracle_batches(model, N, B)
    z, ok = _zscores(g, o)
    filled = ok & (o.mean(0) > 0)
    assert filled.sum() > 20
    assert np.abs(z[filled]).max() < 5.5, np.abs(z[filled]).max()
    assert 0.5 < (z[filled] ** 2).mean() < 1.6, (z[filled] ** 2).mean()
    for key in ("n_crossings", "n_absorptions", "n_scatterings"):
        a = np.mean([s[key] for s in gst])
        b = np.mean([s[key] for s in ost])
        assert abs(a / b - 1) < 0.015, (key, a, b)
