"""The partial diffusion approximation of the oracle (src/grid/grid_pda_3d.f90) and the per-cell packet counter it
needs (n_photons, src/grid/grid_propagate_3d.f90:90-95,175-180).

The reference's fixture for this path (test_pinte_specific_energy) needs inputs of its Python front end that cannot
be regenerated here, so the restatement is pinned by known answers of the diffusion equation:

* a pocket of unsampled cells surrounded by cells of one mean energy takes exactly that energy;
* with opacities that do not depend on the energy the discrete equation is Laplace's: a profile linear in x on the
  sampled cells is continued linearly through the pocket, for the dense solver (to its 1e-5) and the Gauss-Seidel
  solver (to the accuracy its stopping rule of 1e-4 per sweep gives);
* n_photons counts packets, not crossings: one packet per cell it visits, whatever its path.
"""
import copy

import numpy as np
import pytest

from helpers import bitlevel_model, bitlevel_model_sph


def _model(golden_car, n=(12, 7, 6), flat_opacities=False):
    m = bitlevel_model(golden_car, False, False)
    n1, n2, n3 = n
    m.w1, m.w2, m.w3 = [np.linspace(-1e18, 1e18, k + 1) for k in (n1, n2, n3)]
    m.density = np.full((1, n3, n2, n1), 1e-17)
    m.sources = m.sources[:1]
    m.sources[0].position = (0., 0., 0.)
    if flat_opacities:
        d = copy.deepcopy(m.dust[0])
        d.kappa_planck = np.full_like(d.kappa_planck, 2.0)
        d.chi_rosseland = np.full_like(d.chi_rosseland, 5.0)
        m.dust = [d]
    m.conf.use_pda = True
    return m


def _pocket(shape):
    n3, n2, n1 = shape
    counts = np.full(shape, 1000, dtype=np.int64)
    counts[2:n3 - 2, 2:n2 - 2, 2:n1 - 2] = 0
    return counts


def test_pocket_takes_the_energy_of_its_surroundings(golden_car):
    from oracle import oracle
    m = _model(golden_car)
    o = oracle.Oracle(m)
    d = m.dust[0]
    s0 = float(np.sqrt(d.specific_energy[5] * d.specific_energy[6]))
    se = np.full(m.density.shape, s0)
    counts = _pocket(se.shape[1:])
    rng = np.random.default_rng(1)
    se[0][counts == 0] *= 10. ** rng.uniform(-1, 1, int((counts == 0).sum()))
    o.put_specific_energy(se)
    o.set_n_photons(counts)
    assert o.solve_pda() == int((counts == 0).sum())
    assert np.allclose(o.get_specific_energy(), s0, rtol=3e-5)


@pytest.mark.parametrize("solver", ["exact", "iterative"])
def test_linear_profile_is_continued_through_the_pocket(golden_car, solver):
    from oracle import oracle
    m = _model(golden_car, flat_opacities=True)
    o = oracle.Oracle(m)
    if solver == "iterative":
        o.set_pda_exact_limit(0)
    x = 0.5 * (m.w1[1:] + m.w1[:-1])
    profile = 2.0 * (1e-3 + 4e-4 * (x - x[0]) / (x[-1] - x[0]))          # s = kappa_planck * e_mean
    se = np.broadcast_to(profile, m.density.shape).copy()
    counts = _pocket(se.shape[1:])
    want = se.copy()
    se[0][counts == 0] *= np.random.default_rng(2).uniform(0.3, 3., int((counts == 0).sum()))
    o.put_specific_energy(se)
    o.set_n_photons(counts)
    o.solve_pda()
    got = o.get_specific_energy()
    assert np.array_equal(got[0][counts > 0], want[0][counts > 0])        # sampled cells are left alone
    assert np.allclose(got, want, rtol=2e-5 if solver == "exact" else 5e-3)


def test_threshold_and_edge_cells(golden_car):
    """do_pda = n_photons < max(30, ceiling(0.005 mean)) and density > 0, never on the edge of the grid
    (grid_pda_3d.f90:124-131, grid_pda_cartesian_3d.f90:24-50)."""
    from oracle import oracle
    m = _model(golden_car)
    m.density[0, 3, 3, 5] = 0.0
    o = oracle.Oracle(m)
    shape = m.density.shape[1:]
    counts = np.full(shape, 29, dtype=np.int64)
    counts[3, 3, 6] = 30
    o.put_specific_energy(np.full(m.density.shape, float(m.dust[0].specific_energy[5])))
    o.set_n_photons(counts)
    n3, n2, n1 = shape
    assert o.solve_pda() == (n1 - 2) * (n2 - 2) * (n3 - 2) - 2      # the empty cell and the one with 30 packets
    counts[...] = 20000                                              # mean 20000 -> limit 100
    counts[2, 2, 2:5] = (99, 100, 101)
    o.set_n_photons(counts)
    assert o.solve_pda() == 1


def test_n_photons_counts_packets_per_cell(golden_car):
    from oracle import oracle
    m = _model(golden_car, n=(5, 5, 5))
    m.conf.use_pda = False
    m.conf.count_photons = True
    m.density[...] = 1e-30                                             # straight lines from the centre
    o = oracle.Oracle(m)
    o.run_lucy_iteration(20000)
    n = o.get_n_photons()
    assert n[2, 2, 2] == 20000                                         # every packet starts in the central cell
    assert n.sum() > 3 * 20000 and n[0, 0, 0] > 0
    # a packet is counted once per cell however it scatters: optically thick, isotropic scattering
    m.density[...] = 3e-15
    o = oracle.Oracle(m)
    st = o.run_lucy_iteration(2000)
    n = o.get_n_photons()
    assert st.n_scatterings + st.n_absorptions > 1.5 * 2000 and n.max() <= 2000 and n[2, 2, 2] == 2000


def sph_shell_model(golden_car, golden_sph, n1=16, n2=8, n3=1):
    """A spherical polar grid (2-D for n3 = 1) with an opaque shell that packets from the central star hardly enter."""
    m = bitlevel_model_sph(golden_car, golden_sph, False, False)
    m.sources = m.sources[:1]
    m.sources[0].position = (0., 0., 0.)
    m.w1 = np.concatenate([[0.], np.logspace(15., 17., n1)])
    m.w2 = np.linspace(0., np.pi, n2 + 1)
    m.w3 = np.linspace(0., 2. * np.pi, n3 + 1)
    m.density = np.full((1, n3, n2, n1), 1e-22)
    m.density[0, :, :, n1 // 2:n1 - 1] = 1e-13
    m.conf.use_pda = True
    return m


def test_pda_in_the_lucy_iteration_obeys_the_maximum_principle(golden_sph, golden_car):
    """do_lucy with the PDA on a 2-D spherical grid and so few packets that many cells see fewer than 30: their
    energies are replaced, the others are not, and the replaced values lie between the extremes of the sampled
    cells (the discrete diffusion equation has no interior extrema)."""
    from oracle import oracle
    m = sph_shell_model(golden_car, golden_sph)
    o = oracle.Oracle(m)
    o.run_lucy_iteration(150)
    n = o.get_n_photons()[0]
    se = o.get_specific_energy()[0, 0]
    deep = (n < 30) & (m.density[0, 0] > 0)
    deep[:, [0, -1]] = False
    deep[[0, -1], :] = False
    assert 10 < deep.sum() < deep.size - 10
    m.conf.use_pda = False
    m.conf.count_photons = True
    o2 = oracle.Oracle(m)
    o2.run_lucy_iteration(150)
    assert np.array_equal(o2.get_n_photons()[0], n)               # the same packets
    raw = o2.get_specific_energy()[0, 0]
    assert np.array_equal(raw[~deep], se[~deep]) and np.all(raw[deep] != se[deep])
    assert se[deep].max() <= se[~deep].max() * 1.01 and se[deep].min() >= se[~deep].min() * 0.99
