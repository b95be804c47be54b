"""Dust sublimation (sublimate_dust, src/grid/grid_physics_3d.f90:420-498): modes 1 'fast' (dust removed,
specific energy reset to the minimum), 2 'slow' (density scaled down, energy set to the sublimation value)
and 3 'cap' (energy capped) applied after every Lucy iteration.

Known answer: sublimation is a pure function of the specific energy update_energy_abs has just produced, so
a run with mode m must equal a run with mode 0 (same packets) followed by that function."""
import numpy as np
import pytest

from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource


def _model(mode, e_sub):
    dust = syn.make_dust(syn.REALISTIC_NU, syn.REALISTIC_ALBEDO, syn.REALISTIC_CHI, n_temp=60,
                         sublimation_mode=mode, sublimation_specific_energy=e_sub)
    w = np.linspace(-syn.pc, syn.pc, 8)
    rho = np.full((1, 7, 7, 7), 2.0 / (syn.chi_at(dust, 6e14) * syn.pc))
    src = FlatSource(type=1, luminosity=syn.lsun, temperature=6000., position=(0.05 * syn.pc, 0.02 * syn.pc, -0.03 * syn.pc))
    m = FlatModel(w, w, w, rho, [dust], [src], FlatConf(n_initial_iter=1, n_initial_photons=0))
    m.minimum_specific_energy = np.array([dust.specific_energy[0] * 3.0])
    return m


def _chi_rosseland(dust, e):
    # chi_rosseland (src/dust/dust.f90:81-121): log-log interpolation of the mean opacity against specific energy
    return 10. ** np.interp(np.log10(e), np.log10(dust.specific_energy), np.log10(dust.chi_rosseland))


def _expected(mode, e0, rho0, e_sub, e_min, dust):
    e, rho = e0.copy(), rho0.copy()
    hot = e0 > e_sub
    if mode == 1:
        rho[hot] = 0.0
        e[hot] = e_min
    elif mode == 2:
        rho[hot] = rho0[hot] * e_sub / e0[hot] * (_chi_rosseland(dust, e0[hot]) / _chi_rosseland(dust, e_sub)) ** 2
        e[hot] = e_sub
    elif mode == 3:
        e[hot] = e_sub
    return e, rho, hot


def _run(backend, model, n):
    if backend == "oracle":
        from oracle import oracle
        x = oracle.Oracle(model)
        x.run_lucy_iteration(n)
    else:
        from hyperion_b200.capi import Engine
        x = Engine(0)
        x.load_model(model)
        x.run_lucy_iteration(n)
    e, rho = x.get_specific_energy()[0], x.get_density()[0]
    if backend != "oracle":
        x.close()
    return e, rho


def _check(backend, mode, rtol):
    n = 40000
    e0, rho0 = _run(backend, _model(0, 0.0), n)
    e_sub = float(np.sort(e0.ravel())[-40])            # the 39 hottest cells sublimate
    m = _model(mode, e_sub)
    e, rho = _run(backend, m, n)
    want_e, want_rho, hot = _expected(mode, e0, rho0, e_sub, m.minimum_specific_energy[0], m.dust[0])
    assert 30 <= hot.sum() <= 45
    # cells whose energy sits within the run-to-run rounding of the threshold may fall on either side
    clear = np.abs(e0 / e_sub - 1.0) > 1e-6
    assert clear.sum() > e0.size - 5
    assert np.allclose(e[clear], want_e[clear], rtol=rtol, atol=0)
    assert np.allclose(rho[clear], want_rho[clear], rtol=max(rtol, 1e-12), atol=0)
    if mode == 1:
        assert (rho[hot & clear] == 0).all()
    return e, rho


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_oracle_sublimation_modes(mode):
    _check("oracle", mode, 1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_gpu_sublimation_modes(mode):
    """The epilogue kernel (lucy_finish_kernel) applies the same function; and the next iteration marches
    the changed densities (mode 1: the emptied cells take no more deposits)."""
    from hyperion_b200.capi import Engine
    _check("gpu", mode, 1e-9)
    if mode == 1:
        n = 40000
        e0, _ = _run("gpu", _model(0, 0.0), n)
        m = _model(1, float(np.sort(e0.ravel())[-40]))
        eng = Engine(0)
        eng.load_model(m)
        eng.run_lucy_iteration(n, iteration=1)
        emptied = eng.get_density()[0] == 0
        eng.lucy_begin()
        eng.lucy_photons(n, n, 2)
        sums = eng.get_energy_sum()[0]
        eng.lucy_finish()
        eng.close()
        assert emptied.sum() >= 30 and (sums[emptied] == 0).all() and (sums[~emptied] > 0).all()
