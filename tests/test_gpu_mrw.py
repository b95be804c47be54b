"""GPU parity of the modified random walk (src/grid/grid_mrw_3d.f90) against the reference's known
answers (hyperion/model/tests/test_mrw.py) and against the oracle."""
import numpy as np
import pytest

from test_oracle_mrw import D_REF, T_REF, mrw_model, temperature_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("density_ref,temperature_ref", list(zip(D_REF, T_REF))[::2] + [(D_REF[-1], T_REF[-1])])
def test_single_temperature_gpu(density_ref, temperature_ref):
    """test_single_temperature through the C ABI: same one-cell model, 30 Lucy iterations with the MRW;
    the reference's criterion is 10 %."""
    from hyperion_b200 import synthetic as syn
    from hyperion_b200.capi import Engine
    dust = syn.realistic_dust(n_temp=40)
    eng = Engine(0)
    eng.load_model(mrw_model(density_ref, dust))
    for it in range(12):
        st = eng.run_lucy_iteration(20000, iteration=it + 1)
    se = eng.get_specific_energy()
    eng.close()
    assert st.killed_int == 0 and st.n_photons == 20000
    t = temperature_of(dust, se.ravel()[0])
    assert temperature_ref / t < 1.1 and t / temperature_ref < 1.1, (t, temperature_ref)
    assert abs(t / temperature_ref - 1) < 0.04, (t, temperature_ref)


def _dense_model(density=200., n=3, peeled=None):
    """n^3 cells over a 2 cm box of the realistic dust at ~50 K, dense enough that the random walk
    replaces a large share of the interactions (alpha_inv_planck * R0 > gamma away from the walls)."""
    from hyperion_b200 import synthetic as syn
    from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource
    dust = syn.realistic_dust(n_temp=40)
    w = np.linspace(-1., 1., n + 1)
    rng = np.random.default_rng(3)
    conf = FlatConf(use_mrw=True, mrw_gamma=2., n_mrw_max=100000, n_inter_max=1000000000)
    rho = density * (1 + rng.random((1, n, n, n)))
    k = int(np.argmin(np.abs(dust.temperature - 50.)))
    se = np.full((1, n, n, n), dust.specific_energy[k])
    m = FlatModel(w, w, w, rho, [dust], [FlatSource(type=1, luminosity=1., temperature=6000., position=(0.1, 0.05, -0.2))],
                  conf, specific_energy=se)
    if peeled:
        m.peeled = peeled
    return m


def test_mrw_deposits_match_oracle():
    """Deposits of one Lucy iteration with the MRW engaged (about a third of the interactions are
    replaced by random-walk steps at this density) against the oracle, batch z-scores per cell."""
    import test_gpu_parity as T
    model = _dense_model()
    B, N = 12, 2000
    g, gst = T._gpu_batches(model, N, B)
    o, ost = T._oracle_batches(model, N, B)
    z, ok = T._zscores(g, o)
    assert ok.all()
    assert np.abs(z).max() < 5.0, np.abs(z).max()
    assert 0.4 < (z ** 2).mean() < 1.8, (z ** 2).mean()
    for key in ("n_absorptions", "n_scatterings"):
        a = np.mean([s[key] for s in gst])
        b = np.mean([s[key] for s in ost])
        assert abs(a / b - 1) < 0.05, (key, a, b)
    assert all(s["killed_int"] == 0 and s["n_escaped"] == N for s in gst)


def test_mrw_imaging_matches_oracle():
    """Imaging iteration with the MRW: every random-walk step is peeled off as an isotropic emission
    (grid_do_mrw_noenergy + peeloff_photon, iter_final.f90:166-185)."""
    from hyperion_b200.flatmodel import FlatPeeledGroup
    import test_gpu_imaging as TI
    grp = FlatPeeledGroup(theta=[40., 130.], phi=[20., 250.], wavelengths=(6, 1., 3000.),
                          image=(3, 3, -1.5, 1.5, -1.5, 1.5), sed=(2, 0.5, 2.), stokes=False, track_origin="basic")
    model = _dense_model(density=60., peeled=[grp])
    gpu, orc = TI._run_both(model, 16, 4000, False, None)
    print(TI._compare(gpu, orc))
    for key in ("n_absorptions", "n_peeloffs"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.05, (key, a, b)


def test_interrupted_random_walks_give_the_same_images(monkeypatch):
    """The peel-off queue of a round holds two jobs per random-walk step; a walk that finds it nearly full is
    interrupted, the packet takes a flight of optical depth zero and the walk resumes in the next round.  With the
    smallest queue (HYPERION_B200_JOBS=4 pools instead of 8) and a medium in which most interactions are random
    walks of many steps, walks ARE interrupted, and since every packet keeps its own random-number stream the cubes
    must equal those of the run with the large queue to rounding."""
    from hyperion_b200.capi import Engine
    from hyperion_b200.flatmodel import FlatPeeledGroup
    grp = FlatPeeledGroup(theta=[40., 130.], phi=[20., 250.], wavelengths=(6, 1., 3000.),
                          image=(3, 3, -1.5, 1.5, -1.5, 1.5), sed=(2, 0.5, 2.), stokes=False, track_origin="basic")
    model = _dense_model(density=600., peeled=[grp])

    def run(jobs):
        monkeypatch.setenv("HYPERION_B200_JOBS", str(jobs))
        eng = Engine(0)
        eng.load_model(model)
        eng.final_begin()
        eng.final_photons(0, 20000, False)
        st = eng.final_finish()
        out = eng.sed(0).copy(), eng.image(0).copy(), st.as_dict()
        eng.close()
        return out

    big, small = run(64), run(4)
    print("peel-offs per packet %.1f, rounds %d (large queue) / %d (small queue)" %
          (big[2]["n_peeloffs"] / 20000. / 2., big[2]["n_rounds"], small[2]["n_rounds"]))
    assert big[2]["n_peeloffs"] > 20 * 20000          # long walks: tens of peeled steps per packet and view
    assert small[2]["n_rounds"] > big[2]["n_rounds"]  # walks were interrupted
    for k in ("n_peeloffs", "n_absorptions", "n_scatterings", "killed_int"):
        assert big[2][k] == small[2][k], k
    for a, b in zip(big[:2], small[:2]):
        nz = a != 0
        assert np.array_equal(nz, b != 0) and np.abs(b[nz] / a[nz] - 1.).max() < 1e-9
