"""GPU parity of the partial diffusion approximation and of the per-cell packet counter (src/grid/grid_pda_3d.f90,
src/grid/grid_propagate_3d.f90:90-95,175-180): CUDA engine vs the CPU oracle, whose restatement is pinned by the
known answers of tests/test_oracle_pda.py.

The reference's inner solvers (dense Gauss elimination / Gauss-Seidel sweeps in cell order) are sequential; the
engine relaxes the same linear system with Jacobi sweeps to 1e-9 per sweep, so for the same packet counts it must
land on the exact solver's answer to that solver's own tolerance (1e-5 in the outer loop).
"""
import numpy as np
import pytest

from test_oracle_pda import _model, _pocket, sph_shell_model

pytestmark = pytest.mark.gpu


def _engine(m):
    from hyperion_b200.capi import Engine
    eng = Engine(0)
    eng.load_model(m)
    return eng


def _noisy_energy(m, counts, seed):
    d = m.dust[0]
    rng = np.random.default_rng(seed)
    lo, hi = np.log10(d.specific_energy[3]), np.log10(d.specific_energy[-4])
    return 10. ** rng.uniform(lo, hi, m.density.shape)


@pytest.mark.parametrize("grid", ["car", "sph2d", "sph3d", "cyl"])
def test_solver_matches_oracle(golden_car, golden_sph, grid):
    """solve_pda on random energies and the same packet counts, Cartesian and polar grids (geometrical factors,
    periodic phi, two-dimensional grids), two dust types on the Cartesian grid."""
    from oracle import oracle
    if grid == "car":
        m = _model(golden_car)
        m.density = np.concatenate([m.density, 0.3 * m.density])
        m.dust = [m.dust[0], m.dust[0]]
        m.density *= np.random.default_rng(5).uniform(0.2, 5., m.density.shape)
    else:
        m = sph_shell_model(golden_car, golden_sph, n1=14, n2=9, n3=1 if grid == "sph2d" else 5)
        m.density[...] = 1e-16 * np.random.default_rng(6).uniform(0.2, 5., m.density.shape)
        if grid == "cyl":
            m.grid_type = "cyl"
            m.w2 = np.linspace(-1e17, 1e17, len(m.w2))
    shape = m.density.shape[1:]
    counts = np.random.default_rng(7).integers(0, 80, shape).astype(np.int64)
    se = _noisy_energy(m, counts, 8)
    o = oracle.Oracle(m)
    o.put_specific_energy(se)
    o.set_n_photons(counts)
    n_pda = o.solve_pda()
    want = o.get_specific_energy()
    eng = _engine(m)
    eng.set_specific_energy_array(se)
    got_n = eng.solve_pda(counts)
    got = eng.get_specific_energy()
    assert got_n == n_pda and n_pda > 20
    changed = want != np.asarray(se)
    assert np.array_equal(got != np.asarray(se), changed)
    assert np.allclose(got, want, rtol=2e-4), np.abs(got / want - 1).max()
    eng.close()


def test_known_answer_on_the_gpu(golden_car):
    m = _model(golden_car, flat_opacities=True)
    x = 0.5 * (m.w1[1:] + m.w1[:-1])
    profile = 2.0 * (1e-3 + 4e-4 * (x - x[0]) / (x[-1] - x[0]))
    want = np.broadcast_to(profile, m.density.shape).copy()
    counts = _pocket(want.shape[1:])
    se = want.copy()
    se[0][counts == 0] *= np.random.default_rng(2).uniform(0.3, 3., int((counts == 0).sum()))
    eng = _engine(m)
    eng.set_specific_energy_array(se)
    assert eng.solve_pda(counts) == int((counts == 0).sum())
    assert np.allclose(eng.get_specific_energy(), want, rtol=2e-5)
    eng.close()


def test_n_photons_matches_oracle(golden_car, golden_sph):
    """The packet counter: every packet starts in the source's cell, and the per-cell counts of independent packet
    sets agree within their Poisson noise, on a Cartesian and a spherical grid."""
    from oracle import oracle
    for m in (_model(golden_car, n=(7, 6, 5)), sph_shell_model(golden_car, golden_sph, n1=10, n2=6, n3=4)):
        m.conf.use_pda = False
        m.conf.count_photons = True
        if m.grid_type == "car":
            m.density[...] = 3e-18
        else:
            m.density[...] = 1e-17
        n = 200000
        o = oracle.Oracle(m)
        o.run_lucy_iteration(n)
        a = o.get_n_photons().astype(float)
        eng = _engine(m)
        eng.run_lucy_iteration(n)
        b = eng.get_n_photons().astype(float)
        eng.close()
        assert a.max() <= n and (m.grid_type != "car" or (a.max() == n and b.max() >= n))
        # The oracle counts distinct packets per cell.  On the GPU all packets are in flight together: a packet that
        # comes back to a cell after another packet has passed through it is counted again, which matters only in
        # the busiest cells (far from the PDA threshold of 30 packets).
        z = (a - b) / np.sqrt(a + b + 1.)
        rel = b / a - 1.
        print("n_photons GPU / oracle - 1: max %.4f, total %.4f, max |z| %.2f" % (rel.max(), b.sum() / a.sum() - 1., np.abs(z).max()))
        # never fewer than the oracle beyond the noise, at most half more in the busiest cell (the source's, where packets keep coming back), and the cells that
        # fewer than 2 % of the packets reach (where the PDA threshold lives) agree within the noise
        assert z.max() < 6. and rel.max() < 0.5 and b.sum() / a.sum() - 1. < 0.10
        quiet = a < 0.02 * n
        assert quiet.sum() > 20 and np.abs(z[quiet]).max() < 6.


def test_lucy_iteration_with_pda(golden_car, golden_sph):
    """do_lucy with use_pda on the engine: poorly sampled interior cells are replaced by values inside the range of
    the sampled ones, and the counts travel through the reduction buffer."""
    m = sph_shell_model(golden_car, golden_sph)
    eng = _engine(m)
    eng.lucy_begin()
    eng.lucy_photons(0, 150, 1)
    buf = eng.reduction_buffer()
    nc = m.density[0].size
    assert buf.numel() == nc + 11 + nc
    eng.lucy_finish()
    n = eng.get_n_photons()[0]
    se = eng.get_specific_energy()[0, 0]
    deep = (n < 30) & (m.density[0, 0] > 0)
    deep[:, [0, -1]] = False
    deep[[0, -1], :] = False
    assert 10 < deep.sum() < deep.size - 10
    assert se[deep].max() <= se[~deep].max() * 1.01 and se[deep].min() >= se[~deep].min() * 0.99
    m.conf.use_pda = False
    m.conf.count_photons = True
    e2 = _engine(m)
    e2.run_lucy_iteration(150)
    raw = e2.get_specific_energy()[0, 0]
    # (which of a packet's return visits count depends on how the packets interleave: a few counts may differ)
    assert np.abs(e2.get_n_photons()[0] - n).max() <= 3
    assert np.allclose(raw[~deep], se[~deep], rtol=1e-12) and np.all(raw[deep] != se[deep])
    e2.close()
    eng.close()
