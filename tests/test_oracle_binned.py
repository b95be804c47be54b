"""Binned images (src/images/images_binned.f90): escaping packets of the final iteration are binned by
their own direction.  The reference has no golden file with a binned group, so the oracle is pinned by
the identity the method rests on: for a source seen the same from every direction, the SED of every
direction bin equals the peeled SED (both estimate the flux an observer would measure)."""
import numpy as np

from helpers import pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatPeeledGroup, FlatSource


def _model(rho, n=6):
    dust = syn.grey_dust(n_temp=10)
    w = np.linspace(-pc, pc, n + 1)
    src = FlatSource(type=1, luminosity=lsun, temperature=5000., position=(0., 0., 0.))
    m = FlatModel(w, w, w, np.full((1, n, n, n), rho), [dust], [src], FlatConf(forced_first_interaction=False))
    m.peeled = [FlatPeeledGroup(theta=[40.], phi=[70.], wavelengths=(6, 0.1, 10.), sed=(1, 0.01 * pc, 3 * pc),
                                image=(3, 3, -pc, pc, -pc, pc), stokes=False)]
    m.binned = FlatPeeledGroup(binned=True, n_theta=3, n_phi=4, wavelengths=(6, 0.1, 10.), sed=(1, 0.01 * pc, 3 * pc),
                               image=(3, 3, -pc, pc, -pc, pc), stokes=False, uncertainties=True)
    return m


def test_binned_sed_equals_peeled_sed_in_vacuum():
    from oracle import oracle
    m = _model(0.0)
    o = oracle.Oracle(m)
    o.final_begin()
    o.final_photons(240000, False)
    st = o.final_finish().as_dict()
    assert st["n_escaped"] == 240000
    peeled = o.sed(0)[0, 0, 0, 0, :]
    binned, unc = o.sed(1, True)
    assert binned.shape == (1, 1, 12, 1, 6)
    b = binned[0, 0, :, 0, :]
    assert peeled.sum() > 0
    # 20000 packets per direction bin: a few per cent of noise in the populated wavelength bins
    good = peeled > 0.02 * peeled.max()
    rel = b[:, good] / peeled[None, good] - 1.0
    assert np.abs(rel).max() < 0.12, np.abs(rel).max()
    assert abs(b.sum() / (12 * peeled.sum()) - 1.0) < 0.01
    # the uncertainty cube is the standard error of the same sum
    assert np.all(unc[0, 0, :, 0, :][:, good] > 0)
    # every packet sits at the centre of its image: the central pixel holds everything
    img = o.image(1)[0, 0]
    assert img.shape == (12, 3, 3, 6)
    assert np.isclose(img[:, 1, 1, :].sum(), img.sum())


def test_binned_group_rules():
    """setup_final_iteration (src/main/setup_rt.f90:318-331)."""
    import pytest
    from oracle import oracle
    m = _model(0.0)
    m.conf.forced_first_interaction = True
    with pytest.raises(Exception, match="can't use binned images with forced first interaction"):
        oracle.Oracle(m)
