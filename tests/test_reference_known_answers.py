"""Known answers of the reference's own tests that run its binary on one-cell models
(hyperion/model/tests/test_helpers.py:45-56 get_test_model_noimaging: cube [-1, 1]^3, one packet, one
iteration, a 1000 K point source of unit luminosity, the 2-point test dust of get_test_dust :14-18).
The oracle reproduces them on the CPU; the CUDA engine must as well (marked gpu)."""
import numpy as np
import pytest

from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource


def _test_dust():
    return syn.make_dust([3.e9, 3.e16], [0.5, 0.5], [1., 1.], n_temp=10, temp_min=0.1, temp_max=1600.)


def one_cell_model(n_dust=1, minimum=None, specific_energy=None, additional=False):
    dust = _test_dust()
    w = np.array([-1., 1.])
    return FlatModel(w, w, w, np.full((n_dust, 1, 1, 1), 1.0), [dust] * n_dust,
                     [FlatSource(type=1, luminosity=1., temperature=1000.)],
                     FlatConf(specific_energy_additional=additional),
                     specific_energy=specific_energy, minimum_specific_energy=minimum)


def _run(model, n_iter, backend):
    if backend == "oracle":
        from oracle import oracle
        x = oracle.Oracle(model)
    else:
        from hyperion_b200.capi import Engine
        x = Engine(0)
        x.load_model(model)
    for it in range(n_iter):
        if backend == "oracle":
            x.run_lucy_iteration(1)
        else:
            x.run_lucy_iteration(1, iteration=it + 1)
    e = x.get_specific_energy().reshape(len(model.dust))
    if backend != "oracle":
        x.close()
    return e


BACKENDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("backend", BACKENDS)
def test_minimum_specific_energy(backend):
    """test_minimum_energy.py:150-260: one packet cannot lift the cell above the requested minimum, so the
    output is the minimum itself, per dust type (10 ulp)."""
    e = _run(one_cell_model(1, np.array([2.])), 1, backend)
    assert abs(e[0] - 2.) <= 10 * np.spacing(2.)
    e = _run(one_cell_model(2, np.array([2., 3.])), 1, backend)
    assert abs(e[0] - 2.) <= 10 * np.spacing(2.) and abs(e[1] - 3.) <= 10 * np.spacing(3.)


@pytest.mark.parametrize("backend", BACKENDS)
def test_minimum_temperature(backend):
    """test_minimum_energy.py:25-135: set_minimum_temperature converts through the dust's
    temperature <-> specific energy table (hyperion/dust/dust_type.py:479-545); 10 K and 8 K come back."""
    dust = _test_dust()
    t2e = lambda t: 10 ** np.interp(np.log10(t), np.log10(dust.temperature), np.log10(dust.specific_energy))
    e2t = lambda e: 10 ** np.interp(np.log10(e), np.log10(dust.specific_energy), np.log10(dust.temperature))
    e = _run(one_cell_model(2, np.array([t2e(10.), t2e(8.)])), 1, backend)
    assert np.isclose(e2t(e[0]), 10., rtol=1e-13) and np.isclose(e2t(e[1]), 8., rtol=1e-13)


def test_specific_energy_type_oracle():
    """test_specific_energy_type.py:30-58 (three iterations of one packet): 'initial' ends on the minimum,
    'additional' on 2.08583984422 -- the oracle's sequential stream reproduces the reference's packet."""
    two = np.full((1, 1, 1, 1), 2.0)
    assert _run(one_cell_model(1, np.array([0.5]), two), 3, "oracle")[0] == 0.5
    e = _run(one_cell_model(1, np.array([0.5]), two, additional=True), 3, "oracle")[0]
    assert np.isclose(e, 2.08583984422, rtol=1e-11, atol=0)


@pytest.mark.gpu
def test_specific_energy_type_gpu():
    """Same model on the engine: other random numbers, so the deposit of the single packet differs, but
    it is 2 + (what 'initial' mode deposits) and lies in the range one packet can deposit."""
    two = np.full((1, 1, 1, 1), 2.0)
    assert _run(one_cell_model(1, np.array([0.5]), two), 3, "gpu")[0] == 0.5
    e = _run(one_cell_model(1, np.array([0.5]), two, additional=True), 3, "gpu")[0]
    assert 2.0 < e < 2.5


# ---- inside observers: hyperion/model/tests/test_image.py:729-916 -------------------------------------
from hyperion_b200.flatmodel import FlatPeeledGroup  # noqa: E402


def _inside_image(model, n_photons, backend, seed=-1):
    model.conf.seed = seed
    if backend == "oracle":
        from oracle import oracle
        x = oracle.Oracle(model)
    else:
        from hyperion_b200.capi import Engine
        x = Engine(0)
        x.load_model(model)
    x.final_begin()
    if backend == "oracle":
        x.final_photons(n_photons, False)
    else:
        x.final_photons(0, n_photons, False)
    x.final_finish()
    img = x.image(0)[0, 0, :, :, :, 0]          # [n_view, n_y, n_x]
    if backend != "oracle":
        x.close()
    return img


def _inside_model(half, positions, views, d_range=(-np.inf, np.inf), rho=0.0, dust=None):
    dust = dust or _test_dust()
    w = np.array([-half, half])
    srcs = [FlatSource(type=1, luminosity=1., temperature=6000., position=tuple(p)) for p in positions]
    m = FlatModel(w, w, w, np.full((1, 1, 1, 1), rho), [dust], srcs, FlatConf())
    m.peeled = [FlatPeeledGroup(theta=[v[0] for v in views], phi=[v[1] for v in views], wavelengths=(1, 1., 1000.),
                                inside_observer=True, peeloff_origin=(0., 0., 0.), image=(360, 180, 180., -180., -90., 90.),
                                stokes=False, d_min=d_range[0], d_max=d_range[1])]
    return m


@pytest.mark.parametrize("backend", BACKENDS)
def test_inside_observer_flux_dilution(backend):
    """test_image.py:835-871: two equal sources at d1 = 2 and d2 = 5 give fluxes in the ratio (d2 / d1)^2,
    also with a nonzero depth minimum."""
    d1, d2 = 2., 5.
    m = _inside_model(8., [(d1, 0., 0.), (0., d2, 0.)], [(90., 0.)], d_range=(1., 20.))
    val = _inside_image(m, 200000, backend)[0]
    brightest = np.sort(val[val > 0])[::-1]
    assert len(brightest) >= 2
    assert np.isclose(brightest[0] / brightest[1], (d2 / d1) ** 2, rtol=0.08)


@pytest.mark.parametrize("backend", BACKENDS)
def test_inside_observer_peeloff_optical_depth(backend):
    """test_image.py:874-916: purely absorbing flat dust, tau = 1 from the source to the observer; the
    dust / no-dust ratio of the direct light is exp(-1) whatever the depth minimum."""
    d, chi, rho = 1.e16, 1., 1.e-16
    dust = syn.make_dust([3.e9, 3.e16], [0., 0.], [chi, chi], n_temp=10, temp_min=0.1, temp_max=1.e4)
    peak = []
    for r in (rho, 0.0):
        m = _inside_model(2.e16, [(d, 0., 0.)], [(90., 0.)], d_range=(0.5e16, 3.e16), rho=r, dust=dust)
        peak.append(_inside_image(m, 100000, backend).max())
    assert np.isclose(peak[0] / peak[1], np.exp(-chi * rho * d), rtol=0.05)


@pytest.mark.parametrize("backend", BACKENDS)
def test_inside_observer_sky_coordinates(backend):
    """test_image.py:752-820: the source appears where its arrival direction, expressed in the local
    spherical frame (r, phi, -theta) of the viewing direction, says it should, for nine viewing directions."""
    def direction(theta, phi):
        t, p = np.radians(theta), np.radians(phi)
        return np.array([np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)])

    def expected_lonlat(d, theta_v, phi_v):
        t, p = np.radians(theta_v), np.radians(phi_v)
        st, ct, cp, sp = np.sin(t), np.cos(t), np.cos(p), np.sin(p)
        vx, vy, vz = np.dot([st * cp, st * sp, ct], d), np.dot([-sp, cp, 0.], d), np.dot([-ct * cp, -ct * sp, st], d)
        return np.degrees(np.arctan2(vy, vx)), np.degrees(np.arctan2(np.hypot(vx, vy), vz)) - 90.

    photon_dir = direction(90., 0.)
    views = [(90., 0.), (90., 90.), (90., 150.), (60., 0.), (120., 0.), (89., 0.), (91., 0.), (45., 30.), (135., 200.)]
    m = _inside_model(1., [tuple(-0.5 * photon_dir)], views)
    val = _inside_image(m, 20000, backend)
    n_y, n_x = val.shape[1], val.shape[2]
    xmin, xmax, ymin, ymax = 180., -180., -90., 90.
    for iv, (theta_v, phi_v) in enumerate(views):
        ys, xs = np.nonzero(val[iv] > 0)
        assert len(xs) > 0
        wgt = val[iv][ys, xs]
        lon = ((xmin + (xs + 0.5) * (xmax - xmin) / n_x) * wgt).sum() / wgt.sum()
        lat = ((ymin + (ys + 0.5) * (ymax - ymin) / n_y) * wgt).sum() / wgt.sum()
        exp_lon, exp_lat = expected_lonlat(photon_dir, theta_v, phi_v)
        dlon = (lon - exp_lon + 180.) % 360. - 180.
        assert abs(dlon) < 2. and abs(lat - exp_lat) < 2., (theta_v, phi_v, lon, lat, exp_lon, exp_lat)


# ---- filters: hyperion/model/tests/test_filters.py:19-99 ----------------------------------------------
def _reference_filter(wav_um, tr, alpha, detector, wav0_um):
    """Filter.to_hdf5_group (hyperion/filter/filter.py:90-126): the normalised transmission the front end
    writes into the .rtin ('tn' column) and the central frequency nu0."""
    c = 29979245800.
    nu = c / (np.asarray(wav_um, dtype=float) * 1.e-4)
    tr = np.asarray(tr, dtype=float)
    order = np.argsort(nu)
    nu, tr = nu[order], tr[order]
    nu0 = c / (wav0_um * 1.e-4)
    beta = -1 if detector == "energy" else 0
    y = tr / nu ** (1. + alpha + beta)
    integral = np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(nu))
    tn = tr / nu ** (1 + beta) / nu0 ** alpha / integral
    return nu, tn * nu, nu0


def _filter_model():
    f1 = _reference_filter([1, 1.1, 1.2, 1.3], [0., 1.0, 0.5, 0.], 0., "photons", 1.15)
    f2 = _reference_filter([2, 2.1, 2.2, 2.3, 2.4], [0., 0.5, 1.0, 0.6, 0.], 1., "energy", 2.15)
    w = np.array([-1., 1.])
    srcs = [FlatSource(type=1, luminosity=1., temperature=6000.), FlatSource(type=1, luminosity=1., temperature=6000.)]
    m = FlatModel(w, w, w, np.full((1, 1, 1, 1), 1.0), [_test_dust()], srcs, FlatConf())
    m.peeled = [FlatPeeledGroup(theta=[1., 2., 3.], phi=[1., 2., 3.], filters=[f1, f2], image=(10, 20, -1., 1., -1., 1.),
                                sed=(1, 1e-30, 1e30))]
    return m, np.array([f1[2], f2[2]])


def _filter_sums(m, nu0, n_photons, backend):
    if backend == "oracle":
        from oracle import oracle
        x = oracle.Oracle(m)
    else:
        from hyperion_b200.capi import Engine
        x = Engine(0)
        x.load_model(m)
    x.final_begin()
    if backend == "oracle":
        x.final_photons(n_photons, False)
    else:
        x.final_photons(0, n_photons, False)
    x.final_finish()
    img = x.image(0)[0, 0]            # [n_view, n_y, n_x, n_filt], stokes I, no origin tracking
    if backend != "oracle":
        x.close()
    # ModelOutput.get_image(units='MJy/sr', distance=1) (hyperion/model/model_output.py:733-797)
    pix_area_sr = (np.arctan(1.) - np.arctan(-1.)) / 10. * (np.arctan(1.) - np.arctan(-1.)) / 20.
    val = img * 1.e17 / nu0 / pix_area_sr / (4. * np.pi)
    return np.array([val[..., 0].sum(), val[..., 1].sum()])


def test_filter_image_values():
    """test_filters.py:94-99: two 6000 K unit sources in a tau = 1 cube of the test dust, three views,
    10 x 20 pixels, filters F1 (photon detector, alpha 0, 1.15 micron) and F2 (energy detector, alpha 1,
    2.15 micron), 1000 packets: sum of the image in MJy/sr at distance 1 = 3438.06 / 2396.48.  The
    reference's criterion is 10 %; the oracle, drawing the reference's own random numbers, reproduces the
    two sums the real Fortran binary produced to the last digit."""
    m, nu0 = _filter_model()
    sums = _filter_sums(m, nu0, 1000, "oracle")
    assert np.allclose(sums, [3438.059082285024, 2396.4803378036186], rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_filter_image_values_gpu():
    """The engine draws other random numbers, and 1000 packets put only a few dozen into each filter: the
    comparison is made at 2e5 packets against the oracle at 2e5 packets (about 1.5 % of noise each)."""
    m, nu0 = _filter_model()
    g = _filter_sums(m, nu0, 200000, "gpu")
    o = _filter_sums(m, nu0, 200000, "oracle")
    assert np.allclose(g, o, rtol=0.06), (g, o)


@pytest.mark.parametrize("backend", BACKENDS)
def test_sed_uncertainty_sum_of_squares(backend):
    """test_sed.py:472-501: one isotropic source, no dust, one SED bin: every packet has the same weight, so
    the sum-of-squares estimator gives sigma / flux = 1 / sqrt(N)."""
    n = 10000
    w = np.array([-1., 1.])
    m = FlatModel(w, w, w, np.zeros((1, 1, 1, 1)), [_test_dust()], [FlatSource(type=1, luminosity=1., temperature=6000.)],
                  FlatConf(seed=-1))
    m.peeled = [FlatPeeledGroup(theta=[45.], phi=[45.], wavelengths=(1, 0.01, 5000.), sed=(1, 1.e10, 1.e10),
                                uncertainties=True)]
    if backend == "oracle":
        from oracle import oracle
        x = oracle.Oracle(m)
    else:
        from hyperion_b200.capi import Engine
        x = Engine(0)
        x.load_model(m)
    x.final_begin()
    if backend == "oracle":
        x.final_photons(n, False)
    else:
        x.final_photons(0, n, False)
    x.final_finish()
    val, unc = x.sed(0, True)
    if backend != "oracle":
        x.close()
    flux, sigma = np.nansum(val[0]), np.sqrt(np.nansum(unc[0] ** 2))
    assert np.isclose(sigma / flux, 1. / np.sqrt(n), rtol=0.03)
