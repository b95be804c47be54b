"""Frequency-resolved specific energy (specific_energy_spectrum) in the oracle.

The reference's tests (hyperion/model/tests/test_specific_energy_spectrum.py) hold no stored outputs; they check
properties of a fresh run, restated here on the oracle: the spectrum is passive (:79-91), sums to the specific
energy over frequency (:95-100), also with the PDA (:104-113), the modified random walk (:219-243, :396-407), capped
sublimation (:247-268), several processes (:373-392) and other grid types (:151-176, :311-333); bins that cover a
window of the spectrum hold part of the energy (:197-215).  One closed form is added: in an optically thin grid of
grey dust the energy per bin follows the source's blackbody.
"""
import numpy as np
import pytest

from helpers import bitlevel_model, bitlevel_model_amr, bitlevel_model_vor, pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource
from oracle import oracle

EDGES = np.logspace(6., 18., 13)      # _DEFAULT_EDGES of the reference's tests


def _assert_sums(se, se_nu, floor, exclude_empty=False, rtol=1e-10):
    """_assert_spectrum_sums_to_specific_energy (:36-54): cells heated above the floor only (the floor is applied to
    the specific energy, not to the spectrum)."""
    nu_sum = se_nu.sum(axis=0)
    heated = se > floor * (1. + 1.e-6)
    if exclude_empty:
        heated &= nu_sum > 0.
    assert np.count_nonzero(heated) >= 1
    np.testing.assert_allclose(nu_sum[heated], se[heated], rtol=rtol)
    return heated


def _run(model, n, n_iter=3):
    o = oracle.Oracle(model)
    for _ in range(n_iter):
        o.run_lucy_iteration(n)
    return o


@pytest.mark.parametrize("multi", [False, True])
def test_spectrum_is_passive_and_sums_to_specific_energy(golden_car, multi):
    m0 = bitlevel_model(golden_car, False, multi)
    m1 = bitlevel_model(golden_car, False, multi)
    m1.spectrum_bin_edges = EDGES
    o0, o1 = _run(m0, 5000), _run(m1, 5000)
    se = o1.get_specific_energy()
    assert np.array_equal(se, o0.get_specific_energy())          # same random stream, same deposits
    se_nu = o1.get_specific_energy_spectrum()
    assert se_nu.shape == (12,) + se.shape
    heated = _assert_sums(se, se_nu, se.min())
    assert heated.mean() > 0.9
    assert (se_nu.sum(axis=(1, 2, 3, 4)) > 0).sum() >= 4         # stellar and re-emitted light: several bins


def test_thin_grey_grid_follows_the_blackbody():
    """Known answer: optically thin grey dust absorbs kappa x B_nu(T); per bin the fraction of the deposited energy
    is the fraction of the blackbody in the bin."""
    dust = syn.grey_dust(n_temp=10)
    w = np.linspace(-pc, pc, 3)
    T = 6000.
    model = FlatModel(w, w, w, np.full((1, 2, 2, 2), 1e-30), [dust],
                      [FlatSource(type=1, luminosity=lsun, temperature=T, position=(0., 0., 0.))], FlatConf())
    edges = np.logspace(13.5, 15.5, 9)
    model.spectrum_bin_edges = edges
    o = oracle.Oracle(model)
    o.lucy_begin()
    o.lucy_photons(200000)
    total = o.get_energy_sum().sum()
    per_bin = o.get_energy_sum_spectrum().sum(axis=(1, 2, 3, 4))
    nu = np.logspace(12., 16.5, 20001)
    b = syn.B_nu(nu, T)
    cum = np.concatenate([[0.], np.cumsum(0.5 * (b[1:] + b[:-1]) * np.diff(nu))])
    expect = np.diff(np.interp(edges, nu, cum)) / cum[-1]
    got = per_bin / total
    assert abs(got.sum() - expect.sum()) < 0.005
    big = expect > 0.02
    assert big.sum() >= 5
    np.testing.assert_allclose(got[big], expect[big], rtol=0.03)


def test_windowed_bins_hold_part_of_the_energy(golden_car):
    full = bitlevel_model(golden_car, False, False)
    full.spectrum_bin_edges = np.logspace(5., 20., 16)
    win = bitlevel_model(golden_car, False, False)
    win.spectrum_bin_edges = np.logspace(14.8, 15.2, 5)
    s = {}
    for key, m in (("full", full), ("window", win)):
        o = _run(m, 5000)
        se, se_nu = o.get_specific_energy(), o.get_specific_energy_spectrum()
        s[key] = se_nu.sum(axis=0)
        assert np.all(s[key] <= se * (1. + 1.e-6))
    assert 0. < s["window"].sum() < 0.9 * s["full"].sum()


def test_bin_edges_are_validated(golden_car):
    m = bitlevel_model(golden_car, False, False)
    m.spectrum_bin_edges = np.array([1e10, 1e12, 1e11])
    with pytest.raises(Exception, match="specific_energy_spectrum_bin_edges should be strictly increasing"):
        oracle.Oracle(m)


def test_spectrum_with_pda():
    """update_specific_energy rescales the bins of a PDA cell to the new specific energy (grid_pda_3d.f90:63-67);
    cells no packet reached have no spectral shape to rescale."""
    dust = syn.grey_dust(n_temp=40)
    w = np.linspace(-pc, pc, 9)
    rho = np.full((1, 8, 8, 8), 3e-20)
    conf = FlatConf(use_pda=True)
    model = FlatModel(w, w, w, rho, [dust], [FlatSource(type=1, luminosity=lsun, temperature=6000., position=(0., 0., 0.))], conf)
    model.spectrum_bin_edges = EDGES
    o = _run(model, 300, n_iter=2)
    se, se_nu = o.get_specific_energy(), o.get_specific_energy_spectrum()
    _assert_sums(se, se_nu, se.min(), exclude_empty=True, rtol=1e-6)


def mrw_spectrum_model(edges, mrw=True):
    """The model of the reference's test_specific_energy_spectrum_with_mrw (:219-243): 2 x 2 x 2 cells of 1 cm of the
    'realistic' test dust, a 6000 K point source, MRW with gamma = 2 -- at a density of 1e9 instead of 1e5, where the
    random walk takes over most of the interactions from the first iteration on."""
    w = np.array([-1., 0., 1.])
    conf = FlatConf(use_mrw=mrw, mrw_gamma=2., n_mrw_max=1000, n_inter_max=1000000000)
    m = FlatModel(w, w, w, np.full((1, 2, 2, 2), 1.e9), [syn.realistic_dust(n_temp=40)],
                  [FlatSource(type=1, luminosity=1., temperature=6000.)], conf)
    m.spectrum_bin_edges = edges
    return m


def test_spectrum_with_mrw():
    """grid_do_mrw deposits through deposit_specific_energy_spectrum (grid_mrw_3d.f90:84-85): spread like the local
    emissivity, which these edges cover completely."""
    o = oracle.Oracle(mrw_spectrum_model(EDGES))
    o0 = oracle.Oracle(mrw_spectrum_model(None))
    o1 = oracle.Oracle(mrw_spectrum_model(None, mrw=False))
    n_with = [x.run_lucy_iteration(50).n_absorptions for x in (o, o0)]
    assert n_with[0] == n_with[1]
    assert o1.run_lucy_iteration(10).n_absorptions / 10 > 5 * n_with[0] / 50     # the random walk replaces most absorptions
    se, se_nu = o.get_specific_energy(), o.get_specific_energy_spectrum()
    assert np.array_equal(se, o0.get_specific_energy())          # passive with the random walk too (:411-430)
    _assert_sums(se, se_nu, se.min(), rtol=1e-6)


def test_spectrum_with_sublimation_cap(golden_car):
    m = bitlevel_model(golden_car, False, False)
    o = _run(bitlevel_model(golden_car, False, False), 5000, n_iter=1)
    cap = float(np.median(o.get_specific_energy()))
    import copy
    d = copy.copy(m.dust[0])
    d.sublimation_mode, d.sublimation_specific_energy = 3, cap
    m.dust = [d]
    m.spectrum_bin_edges = EDGES
    o = _run(m, 5000)
    se, se_nu = o.get_specific_energy(), o.get_specific_energy_spectrum()
    assert (se == cap).sum() > 10 and (se < cap).sum() > 10
    _assert_sums(se, se_nu, se.min())


def test_spectrum_over_two_processes(golden_car):
    m = bitlevel_model(golden_car, False, True)
    m.spectrum_bin_edges = EDGES
    ranks = [oracle.Oracle(m, rank=r) for r in range(2)]
    for o in ranks:
        o.lucy_begin()
        o.lucy_photons(4000)
    tot, tot_nu = sum(o.get_energy_sum() for o in ranks), sum(o.get_energy_sum_spectrum() for o in ranks)
    e_cur = sum(o.energy_current for o in ranks)
    for o in ranks:
        o.set_energy_sum(tot)
        o.set_energy_sum_spectrum(tot_nu)
        o.energy_current = e_cur
        o.lucy_finish()
    se, se_nu = ranks[0].get_specific_energy(), ranks[0].get_specific_energy_spectrum()
    assert np.array_equal(se_nu, ranks[1].get_specific_energy_spectrum())
    _assert_sums(se, se_nu, se.min())


@pytest.mark.parametrize("grid", ["amr", "vor"])
def test_spectrum_on_other_grids(golden_car, golden_amr, grid):
    m = bitlevel_model_amr(golden_car, golden_amr, False, False) if grid == "amr" else bitlevel_model_vor(golden_car, False, False)
    m.spectrum_bin_edges = EDGES
    o = _run(m, 5000)
    se, se_nu = o.get_specific_energy(), o.get_specific_energy_spectrum()
    assert se_nu.shape == (12,) + se.shape
    _assert_sums(se, se_nu, se.min())


def test_spectrum_with_additional_specific_energy(golden_car):
    """specific_energy_type = 'additional': the extra heating is added to the specific energy after every iteration;
    its spectrum is a copy of the (all-zero) spectrum array (grid_physics_3d.f90:143,223-225,543-545), so the bins
    hold the Monte-Carlo part only."""
    m = bitlevel_model(golden_car, False, False)
    extra = np.full(m.density.shape, 3.e-2)
    m.specific_energy = extra
    m.conf.specific_energy_additional = True
    m.spectrum_bin_edges = EDGES
    o = _run(m, 5000, n_iter=2)
    se, se_nu = o.get_specific_energy(), o.get_specific_energy_spectrum()
    mc = se - extra
    heated = mc > 1e-3 * mc.max()
    assert heated.sum() > 50
    np.testing.assert_allclose(se_nu.sum(axis=0)[heated], mc[heated], rtol=1e-9)
