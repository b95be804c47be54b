"""Monochromatic final iteration of the oracle (src/main/iter_final_mono.f90, src/grid/grid_monochromatic.f90).

The reference's own fixtures for this mode (test_pascucci / test_pinte) need inputs its Python front end
computes (disk densities, extrapolated dust, mean opacities) and cannot be regenerated here, so the mode is
pinned by known answers:

* a blackbody point source in an empty grid: every packet carries normalized_B_nu(nu, T) L / N, so the peeled
  nu F_nu is nu B_nu(T) pi / (sigma T^4) L exactly, at every frequency and from every direction;
* in a dusty grid the monochromatic SED must agree with the polychromatic one (the path pinned bit for bit by
  the reference's test_peeloff files) at the same wavelengths, within the Monte-Carlo noise of both: source
  light, scattered light and thermal emission all go through different code in the two modes.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from helpers import bitlevel_model
from hyperion_b200.flatmodel import FlatPeeledGroup

# lib_constants.f90:52-96, the values normalized_B_nu is built from
H, K, C, SIGMA = 6.6260689633e-27, 1.380650424e-16, 2.99792458e10, 5.670400e-5
pc = 3.08568025e18


def _group(model, n, **kw):
    half = 2.0 * pc
    return FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(1, 0., 4. * half), stokes=True,
                           inu_min=1, inu_max=n, wavelengths=(n, 1., 2.), **kw)


def test_blackbody_point_source_in_vacuum_is_exact(golden_car):
    from oracle import oracle
    m = bitlevel_model(golden_car, False, False)
    m.density[...] = 0.0
    m.sources = m.sources[:1]
    T, L = m.sources[0].temperature, m.sources[0].luminosity
    m.frequencies = C / (np.array([0.3, 0.55, 1.2, 3.6, 24., 160.]) * 1e-4)
    m.peeled = [_group(m, 6)]
    o = oracle.Oracle(m)
    o.final_begin()
    for inu in range(1, 7):
        o.final_mono_photons(inu, 50, 50, 0, 0)
    o.final_finish()
    sed = o.sed(0)[0, 0, :, 0, :]                        # Stokes I: [view, nu]
    nu = m.frequencies
    want = nu * (2. * H / C ** 2 / SIGMA * np.pi) * nu ** 3 / np.expm1(H * nu / K / T) / T ** 4 * L
    assert np.allclose(sed[0], want, rtol=1e-12) and np.allclose(sed[1], want, rtol=1e-12)
    assert not o.sed(0)[1:].any()                        # unpolarised


def test_monochromatic_sed_agrees_with_polychromatic(golden_car):
    """SEDs split by origin (track_origin = 'basic': source direct / source scattered / dust direct / dust
    scattered).  Light that starts on the sources goes through forced scatterings with albedo weights in one
    mode and through absorb-or-scatter in the other and must agree within the noise.  Thermal light is only
    asked to agree to 12 %: the two modes interpolate the emissivity between two temperature states
    differently (nu from the two states with one random number, dust_type_4elem.f90:379-398, against the
    log-interpolated probability, :356-377) and take the thermal luminosity from the specific energy in one
    case and from the packets absorbed during the run in the other -- in the reference as well."""
    import copy
    from oracle import oracle
    m = bitlevel_model(golden_car, False, False)
    m.density *= 10.0                                    # tau ~ 1: scattered and thermal light matter
    base = oracle.Oracle(m)
    for _ in range(5):
        base.run_lucy_iteration(200000)
    m.specific_energy = base.get_specific_energy()
    wav = np.array([0.45, 2.2, 40., 110.])
    n_ranks = 8
    models = [copy.deepcopy(m) for _ in range(2 * n_ranks)]       # the .npz reader is not thread-safe
    views = dict(theta=[30., 110.], phi=[40., 250.], sed=(1, 0., 8. * pc), stokes=False, track_origin="basic")

    def poly(r):
        mm = models[r]
        # bins 10 % wide centred (in the log) on the wavelengths: nu F_nu of the polychromatic run
        mm.peeled = [FlatPeeledGroup(wavelengths=(1, w / 1.05, w * 1.05), **views) for w in wav]
        o = oracle.Oracle(mm, rank=r)
        o.final_begin()
        o.final_photons(600000, False)
        o.final_finish()
        return np.array([o.sed(k)[0, :, :, 0, 0] for k in range(len(wav))])      # [nu, origin, view]

    def mono(r):
        mm = models[n_ranks + r]
        mm.frequencies = C / (wav * 1e-4)
        mm.peeled = [FlatPeeledGroup(inu_min=1, inu_max=len(wav), wavelengths=(len(wav), 1., 2.), **views)]
        o = oracle.Oracle(mm, rank=100 + r)
        o.final_begin()
        for inu in range(1, len(wav) + 1):
            o.final_mono_photons(inu, 20000, 20000, 20000, 20000)
        o.final_finish()
        return np.moveaxis(o.sed(0)[0, :, :, 0, :], -1, 0)                         # [nu, origin, view]

    with ThreadPoolExecutor(max_workers=8) as pool:
        P = np.array(list(pool.map(poly, range(n_ranks))))
        M = np.array(list(pool.map(mono, range(n_ranks))))
    pm, ps = P.mean(0), P.std(0, ddof=1) / np.sqrt(n_ranks)
    mm_, ms = M.mean(0), M.std(0, ddof=1) / np.sqrt(n_ranks)
    err = np.sqrt(ps ** 2 + ms ** 2)
    total = pm.sum(1, keepdims=True)
    # (wavelength, origin, view) entries that carry at least 2 % of the flux at their wavelength
    ok = pm > 0.02 * total
    z = np.where(ok, (mm_ - pm) / np.where(err > 0, err, 1.0), 0.0)
    rel = np.where(ok, mm_ / np.where(pm > 0, pm, 1.0) - 1.0, 0.0)
    print("mono / poly - 1 [nu, origin, view]:\n", np.round(rel, 3), "\nz:\n", np.round(z, 2))
    # orig() (image_type.f90:117-134): 1 source direct, 2 dust direct, 3 source scattered, 4 dust scattered
    S, D = [0, 2], [1, 3]
    src = ok[:, S]
    assert src.sum() >= 4
    assert np.abs(z[:, S][src]).max() < 4.5 and (z[:, S][src] ** 2).mean() < 2.5
    w = np.where(src, pm[:, S] ** 2 / np.where(err[:, S] > 0, err[:, S] ** 2, 1.0), 0.0)
    assert abs((rel[:, S] * w).sum() / w.sum()) < 0.02
    dust = ok[:, D]
    assert dust.sum() >= 2
    assert np.abs(rel[:, D][dust]).max() < 0.12


def test_thermal_emission_of_a_thin_grid_has_closed_form(golden_car):
    """emit_from_monochromatic_grid_pdf (grid_monochromatic.f90:120-174) in a grid too thin to absorb anything:
    the thermal nu F_nu at frequency nu is nu * sum over cells of prob_c(nu) * E_c rho_c V_c, with prob_c the
    emission probability per unit frequency of the cell's emissivity state (dust_sample_emit_probability)."""
    from oracle import oracle
    m = bitlevel_model(golden_car, False, False)
    m.density *= 1e-8
    rng = np.random.default_rng(3)
    d = m.dust[0]
    e_lo, e_hi = d.jnu_var[10], d.jnu_var[60]
    m.specific_energy = 10. ** rng.uniform(np.log10(e_lo), np.log10(e_hi), m.density.shape)
    wav = np.array([12., 70., 350.])
    m.frequencies = C / (wav * 1e-4)
    m.conf.forced_first_interaction = False
    m.peeled = [FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(1, 0., 8. * pc), stokes=False,
                                inu_min=1, inu_max=3, wavelengths=(3, 1., 2.))]
    o = oracle.Oracle(m)
    o.final_begin()
    n = 200000
    for inu in (1, 2, 3):
        o.final_mono_photons(inu, 0, 0, n, n)
    o.final_finish()
    sed = o.sed(0)[0, 0, :, 0, :]                            # [view, nu]
    # closed form
    x = np.asarray(d.emiss_nu)
    vol = (np.diff(m.w3)[:, None, None] * np.diff(m.w2)[None, :, None] * np.diff(m.w1)[None, None, :])
    erv = m.specific_energy[0] * m.density[0] * vol
    ljv = np.log10(np.asarray(d.jnu_var))
    le = np.log10(m.specific_energy[0])
    k = np.clip(np.searchsorted(ljv, le, side="right") - 1, 0, len(ljv) - 2)
    f = (le - ljv[k]) / (ljv[k + 1] - ljv[k])

    def pdf_at(state, nu):
        y = np.asarray(d.emiss_jnu)[:, state]
        with np.errstate(divide="ignore", invalid="ignore"):
            b = np.log(y[1:] / y[:-1]) / np.log(x[1:] / x[:-1])
            seg = np.where(np.abs(b + 1.) < 1e-10, x[:-1] * y[:-1] * np.log(x[1:] / x[:-1]),
                           (y[1:] * x[1:] - y[:-1] * x[:-1]) / (b + 1.))
        norm = np.nansum(np.where((y[1:] > 0) & (y[:-1] > 0), seg, 0.0))
        j = np.searchsorted(x, nu) - 1
        if not (y[j] > 0 and y[j + 1] > 0):
            return 0.0                   # interp1d_loglog gives zero if either end is zero (lib_array.f90:605-614)
        val = y[j] * (nu / x[j]) ** (np.log(y[j + 1] / y[j]) / np.log(x[j + 1] / x[j]))
        return val / norm

    states = np.unique(np.concatenate([k.ravel(), k.ravel() + 1]))
    for i, nu in enumerate(m.frequencies):
        table = {s_: pdf_at(s_, nu) for s_ in states}
        p1 = np.vectorize(table.get)(k)
        p2 = np.vectorize(table.get)(k + 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            prob = np.where((p1 > 0) & (p2 > 0), 10. ** (np.log10(p1) + f * (np.log10(p2) - np.log10(p1))), 0.0)
        want = nu * (prob * erv).sum()
        got = sed[:, i]
        # Monte-Carlo noise only in WHICH cells emit: every packet carries the same energy
        assert np.allclose(got, want, rtol=5e-3), (wav[i], got, want)
