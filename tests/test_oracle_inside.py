"""Inside observers (peeloff_photon, src/images/images_peeled.f90:95-270, the inside_observer branches):
the peel-off direction points at the observer, the march stops at the observer, the image axes are
longitude / latitude in the observer's frame and the weight is the flux 1 / (4 pi d^2).  No golden file
of the reference has an inside observer, so the oracle is pinned by what the geometry dictates."""
import numpy as np

from helpers import pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatPeeledGroup, FlatSource


OBS = np.array([0.05, -0.03, 0.02]) * pc      # not on a cell wall


def _model(rho, src_pos, n=6):
    dust = syn.grey_dust(n_temp=10)
    w = np.linspace(-pc, pc, n + 1)
    src = FlatSource(type=1, luminosity=lsun, temperature=5000., position=tuple(src_pos))
    m = FlatModel(w, w, w, np.full((1, n, n, n), rho), [dust], [src], FlatConf())
    common = dict(wavelengths=(6, 0.01, 1000.), stokes=False)      # wide enough to hold every packet
    m.peeled = [
        FlatPeeledGroup(theta=[90.], phi=[0.], inside_observer=True, peeloff_origin=tuple(OBS),
                        sed=(1, 1.0, 400.), image=(36, 18, 360., 0., -90., 90.), **common),
        FlatPeeledGroup(theta=[40.], phi=[70.], sed=(1, 0.01 * pc, 3 * pc), **common)]
    return m, dust


def test_inside_observer_sees_the_inverse_square_flux_at_the_right_place():
    from oracle import oracle
    l, b, D = np.radians(33.), np.radians(23.), 0.7 * pc
    pos = OBS + D * np.array([np.cos(b) * np.cos(l), np.cos(b) * np.sin(l), np.sin(b)])
    m, _ = _model(0.0, pos)
    o = oracle.Oracle(m)
    o.final_begin()
    o.final_photons(20000, False)
    o.final_finish()
    s_in = o.sed(0)[0, 0, 0, 0, :]
    s_out = o.sed(1)[0, 0, 0, 0, :]
    # vacuum: every packet is peeled with weight 1 towards both observers
    assert np.allclose(s_in, s_out / (4.0 * np.pi * D * D), rtol=1e-12, atol=0)
    img = o.image(0)[0, 0, 0]          # [n_y, n_x, n_wav]
    iy, ix = np.unravel_index(np.argmax(img.sum(axis=-1)), img.shape[:2])
    # arrival direction = - direction to the source: longitude l + 180 deg, latitude b; the x axis runs
    # from 360 down to 0 over 36 pixels, the y axis from -90 to 90 over 18
    assert (ix, iy) == (14, 11)
    assert np.isclose(img[iy, ix].sum(), img.sum())


def test_inside_observer_march_stops_at_the_observer():
    """Uniform grey dust: the attenuation is exp(-chi rho D) over the distance source -> observer only."""
    from oracle import oracle
    D = 0.5 * pc
    m, dust = _model(0.0, OBS + np.array([D, 0., 0.]))
    chi = float(dust.chi[0])
    tau = 0.8
    m.density[...] = tau / (chi * D)
    m.conf.forced_first_interaction = False
    # direct light only: kill packets at their first interaction so that nothing scattered is peeled
    m.conf.kill_on_absorb = m.conf.kill_on_scatter = True
    o = oracle.Oracle(m)
    o.final_begin()
    o.final_photons(20000, False)
    o.final_finish()
    s_in = o.sed(0)[0, 0, 0, 0, :]
    m0, _ = _model(0.0, OBS + np.array([D, 0., 0.]))
    o0 = oracle.Oracle(m0)
    o0.final_begin()
    o0.final_photons(20000, False)
    o0.final_finish()
    s_vac = o0.sed(0)[0, 0, 0, 0, :]
    # the two runs draw different frequencies (interactions consume random numbers), but every packet is
    # peeled once, at emission, with the same grey attenuation: the bolometric sums differ by exp(-tau)
    assert np.isclose(s_in.sum() / s_vac.sum(), np.exp(-tau), rtol=1e-9, atol=0)
