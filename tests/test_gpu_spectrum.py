"""Frequency-resolved specific energy on the GPU: the reference's property tests
(hyperion/model/tests/test_specific_energy_spectrum.py, restated for the oracle in tests/test_oracle_spectrum.py)
through the C ABI, and per-bin parity with the oracle."""
import copy

import numpy as np
import pytest

from helpers import bitlevel_model, bitlevel_model_amr, bitlevel_model_sph, bitlevel_model_vor, pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource
from test_oracle_spectrum import EDGES, _assert_sums, mrw_spectrum_model

pytestmark = pytest.mark.gpu


def _engine(model):
    from hyperion_b200.capi import Engine
    eng = Engine(0)
    eng.load_model(model)
    return eng


def _run(model, n, n_iter=3):
    eng = _engine(model)
    for it in range(n_iter):
        st = eng.run_lucy_iteration(n, it + 1)
        assert st.killed_geo == 0 and st.killed_int == 0
    return eng


@pytest.mark.parametrize("multi", [False, True])
def test_spectrum_is_passive_and_sums_to_specific_energy(golden_car, multi):
    m0 = bitlevel_model(golden_car, False, multi)
    m1 = bitlevel_model(golden_car, False, multi)
    m1.spectrum_bin_edges = EDGES
    e0, e1 = _run(m0, 200000), _run(m1, 200000)
    se, se0 = e1.get_specific_energy(), e0.get_specific_energy()
    se_nu = e1.get_specific_energy_spectrum()
    e0.close()
    e1.close()
    # the same packets (counter RNG per packet) through the generic march instead of the Cartesian kernels
    np.testing.assert_allclose(se, se0, rtol=0.05)
    assert abs(se.sum() / se0.sum() - 1) < 2e-3
    assert se_nu.shape == (12,) + se.shape
    heated = _assert_sums(se, se_nu, se.min(), rtol=1e-9)
    assert heated.mean() > 0.9          # (the coldest cell defines the floor and is left out)


def test_per_bin_deposits_match_oracle(golden_car):
    """B batches on each side; z-scores of the batch means of every (bin, dust, cell) both sides populate."""
    from oracle import oracle
    from concurrent.futures import ThreadPoolExecutor
    model = bitlevel_model(golden_car, False, True)
    model.spectrum_bin_edges = np.logspace(11., 16., 11)
    B, N = 12, 100000
    eng = _engine(model)
    gpu = []
    for b in range(B):
        eng.lucy_begin()
        eng.lucy_photons(b * N, N, 1)
        st = eng.lucy_finish()
        gpu.append(eng.get_specific_energy_spectrum())
        eng.set_specific_energy(eng.ctx, model.specific_energy, model.minimum_specific_energy)
    eng.close()

    def run(r):
        o = oracle.Oracle(model, rank=r)
        o.run_lucy_iteration(N)
        return o.get_specific_energy_spectrum()

    with ThreadPoolExecutor(max_workers=8) as pool:
        orc = list(pool.map(run, range(B)))
    a, b = np.array(gpu), np.array(orc)
    ma, mb = a.mean(0), b.mean(0)
    den = np.sqrt(a.var(0, ddof=1) / B + b.var(0, ddof=1) / B)
    filled = ((a != 0).mean(0) > 0.9) & ((b != 0).mean(0) > 0.9) & (den > 0)
    assert filled.sum() > 1000
    z = (ma[filled] - mb[filled]) / den[filled]
    assert np.abs(z).max() < 5.5, np.abs(z).max()
    assert 0.6 < (z ** 2).mean() < 1.5, (z ** 2).mean()
    # per bin, summed over the grid: a much tighter comparison
    ta, tb = a.sum(axis=(2, 3, 4, 5)), b.sum(axis=(2, 3, 4, 5))
    big = tb.mean(0) > 1e-3 * tb.mean(0).max()
    zt = (ta.mean(0) - tb.mean(0))[big] / np.sqrt(ta.var(0, ddof=1) / B + tb.var(0, ddof=1) / B)[big]
    assert np.abs(zt).max() < 4.5, zt


def test_thin_grey_grid_follows_the_blackbody():
    dust = syn.grey_dust(n_temp=10)
    w = np.linspace(-pc, pc, 3)
    T = 6000.
    model = FlatModel(w, w, w, np.full((1, 2, 2, 2), 1e-30), [dust],
                      [FlatSource(type=1, luminosity=lsun, temperature=T, position=(0., 0., 0.))], FlatConf())
    edges = np.logspace(13.5, 15.5, 9)
    model.spectrum_bin_edges = edges
    eng = _engine(model)
    eng.run_lucy_iteration(4000000)
    se_nu = eng.get_specific_energy_spectrum()
    eng.close()
    per_bin = se_nu.sum(axis=(1, 2, 3, 4))
    nu = np.logspace(12., 16.5, 20001)
    b = syn.B_nu(nu, T)
    cum = np.concatenate([[0.], np.cumsum(0.5 * (b[1:] + b[:-1]) * np.diff(nu))])
    expect = np.diff(np.interp(edges, nu, cum)) / np.diff(np.interp(edges[[0, -1]], nu, cum))[0]
    got = per_bin / per_bin.sum()
    big = expect > 0.02
    np.testing.assert_allclose(got[big], expect[big], rtol=0.01)


def test_spectrum_with_pda():
    dust = syn.grey_dust(n_temp=40)
    w = np.linspace(-pc, pc, 9)
    model = FlatModel(w, w, w, np.full((1, 8, 8, 8), 3e-20), [dust],
                      [FlatSource(type=1, luminosity=lsun, temperature=6000., position=(0., 0., 0.))], FlatConf(use_pda=True))
    model.spectrum_bin_edges = EDGES
    eng = _run(model, 300, n_iter=2)
    se, se_nu = eng.get_specific_energy(), eng.get_specific_energy_spectrum()
    eng.close()
    heated = _assert_sums(se, se_nu, se.min(), exclude_empty=True, rtol=1e-6)
    assert ((se_nu.sum(0) == 0) & (se > se.min() * (1 + 1e-6))).sum() > 0      # cells only the PDA heated
    assert heated.sum() > 50


def test_spectrum_with_mrw():
    eng = _engine(mrw_spectrum_model(EDGES))
    st = eng.run_lucy_iteration(2000)
    se, se_nu = eng.get_specific_energy(), eng.get_specific_energy_spectrum()
    eng.close()
    assert st.n_absorptions < 2000 * 100000      # the random walk is on (without it: 4e5 absorptions per packet)
    _assert_sums(se, se_nu, se.min(), rtol=1e-6)


def test_spectrum_with_sublimation_cap(golden_car):
    m = bitlevel_model(golden_car, False, False)
    eng = _run(bitlevel_model(golden_car, False, False), 100000, n_iter=1)
    cap = float(np.median(eng.get_specific_energy()))
    eng.close()
    d = copy.copy(m.dust[0])
    d.sublimation_mode, d.sublimation_specific_energy = 3, cap
    m.dust = [d]
    m.spectrum_bin_edges = EDGES
    eng = _run(m, 100000)
    se, se_nu = eng.get_specific_energy(), eng.get_specific_energy_spectrum()
    eng.close()
    assert (se == cap).sum() > 10 and (se < cap).sum() > 10
    _assert_sums(se, se_nu, se.min(), rtol=1e-9)


@pytest.mark.parametrize("grid", ["amr", "vor", "sph"])
def test_spectrum_on_other_grids(golden_car, golden_amr, golden_sph, grid):
    m = {"amr": lambda: bitlevel_model_amr(golden_car, golden_amr, False, False),
         "vor": lambda: bitlevel_model_vor(golden_car, False, False),
         "sph": lambda: bitlevel_model_sph(golden_car, golden_sph, False, False)}[grid]()
    m.spectrum_bin_edges = EDGES
    eng = _run(m, 100000)
    se, se_nu = eng.get_specific_energy(), eng.get_specific_energy_spectrum()
    eng.close()
    assert se_nu.shape == (12,) + se.shape
    _assert_sums(se, se_nu, se.min(), rtol=1e-9)


def test_spectrum_through_the_file_boundary(golden_car, tmp_path):
    """output_specific_energy_spectrum = 'last' with bin edges in the .rtin: /iteration_%05d/specific_energy_spectrum
    [n_bins, n_dust, n3, n2, n1] and the 1-D specific_energy_spectrum_bin_edges (test_..._bin_edges_written, :117-127)."""
    from hyperion_b200 import rtin_write, runner
    from hyperion_b200.io import h5min
    m = bitlevel_model(golden_car, False, True)
    m.spectrum_bin_edges = EDGES
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=50000, output_specific_energy_spectrum="last")
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    assert "specific_energy_spectrum" not in r["iteration_00001"]
    se = r["iteration_00002/specific_energy"][...]
    se_nu = r["iteration_00002/specific_energy_spectrum"][...]
    np.testing.assert_allclose(r["iteration_00002/specific_energy_spectrum_bin_edges"][...], EDGES, rtol=1e-12)
    assert se_nu.shape == (12, 3, 3, 5, 7)
    _assert_sums(se, se_nu, se.min(), rtol=1e-9)


def test_spectrum_with_additional_specific_energy(golden_car):
    m = bitlevel_model(golden_car, False, False)
    extra = np.full(m.density.shape, 3.e-2)
    m.specific_energy = extra
    m.conf.specific_energy_additional = True
    m.spectrum_bin_edges = EDGES
    eng = _run(m, 200000, n_iter=2)
    se, se_nu = eng.get_specific_energy(), eng.get_specific_energy_spectrum()
    eng.close()
    mc = se - extra
    heated = mc > 1e-3 * mc.max()
    assert heated.sum() > 50
    np.testing.assert_allclose(se_nu.sum(axis=0)[heated], mc[heated], rtol=1e-8)
