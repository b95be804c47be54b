"""Filter convolution (image_bin with use_filters, src/images/image_type.f90:467-476): every packet adds
energy x transmission(nu) to each filter channel.  Pinned by the identity with a finely binned SED: the
filter channel equals sum over wavelength bins of (bin flux x mean transmission in the bin) when the
transmission is piecewise constant on those bins (a top-hat aligned with the bin edges)."""
import numpy as np

from helpers import pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatPeeledGroup, FlatSource

C_CGS = 2.99792458e10
MICRON = float(np.float32(1.e-4))


def _model(filters):
    dust = syn.grey_dust(n_temp=10)
    n = 4
    w = np.linspace(-pc, pc, n + 1)
    src = FlatSource(type=1, luminosity=lsun, temperature=5000., position=(0.1 * pc, 0., 0.))
    m = FlatModel(w, w, w, np.zeros((1, n, n, n)), [dust], [src], FlatConf())
    m.peeled = [FlatPeeledGroup(theta=[40.], phi=[70.], wavelengths=(8, 0.1, 10.), sed=(1, 0.01 * pc, 3 * pc), stokes=False),
                FlatPeeledGroup(theta=[40.], phi=[70.], filters=filters, sed=(1, 0.01 * pc, 3 * pc),
                                image=(2, 2, -pc, pc, -pc, pc), stokes=False, uncertainties=True)]
    return m


def tophat_filters():
    """Two top-hats spanning bins 3-4 and bin 6 of the 8 log-spaced wavelength bins of 0.1-10 micron, with
    transmission 0.5 and 2.0; edges nudged inwards so that the ramps hold no packet of this test."""
    nu_min, nu_max = C_CGS / (10. * MICRON), C_CGS / (0.1 * MICRON)
    edges = nu_min * (nu_max / nu_min) ** (np.arange(9) / 8.0)      # frequency-bin edges, increasing
    out = []
    for (b0, b1, t) in ((2, 4, 0.5), (5, 6, 2.0)):
        lo, hi = edges[b0], edges[b1]
        nu = np.array([lo * (1 - 1e-12), lo, hi, hi * (1 + 1e-12)])
        out.append((nu, np.array([0.0, t, t, 0.0]), np.sqrt(lo * hi)))
    return out, ((2, 4, 0.5), (5, 6, 2.0))


def test_filter_channels_equal_transmission_weighted_bins():
    from oracle import oracle
    filters, spec = tophat_filters()
    m = _model(filters)
    o = oracle.Oracle(m)
    o.final_begin()
    o.final_photons(50000, False)
    o.final_finish()
    fine = o.sed(0)[0, 0, 0, 0, :]            # nu F_nu per bin = bin sum / dnunorm
    nu_min, nu_max = C_CGS / (10. * MICRON), C_CGS / (0.1 * MICRON)
    dnunorm = (nu_max / nu_min) ** (0.5 / 8) - (nu_max / nu_min) ** (-0.5 / 8)
    sums = fine * dnunorm                     # energy per bin
    filt, unc = o.sed(1, True)
    assert filt.shape == (1, 1, 1, 1, 2)
    for k, (b0, b1, t) in enumerate(spec):
        assert np.isclose(filt[0, 0, 0, 0, k], t * sums[b0:b1].sum(), rtol=1e-10), k
        assert unc[0, 0, 0, 0, k] > 0
    img = o.image(1)
    assert img.shape == (1, 1, 1, 2, 2, 2) and np.isclose(img.sum(), filt.sum(), rtol=1e-12)


def test_filters_and_raytracing_are_exclusive():
    import pytest
    from oracle import oracle
    m = _model(tophat_filters()[0])
    o = oracle.Oracle(m)
    with pytest.raises(Exception, match="filter convolution cannot be used with raytracing"):
        o.raytracing_photons(100, 100)
