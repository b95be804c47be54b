"""Pin the oracle's Voronoi geometry (src/grid/grid_geometry_voronoi.f90).

The reference stores no golden file for this grid type (hyperion/model/tests/test_model.py:306-343 only
checks that a 1000-site run completes, and needs voro++ through the Python front end).  The oracle is
therefore pinned in two ways:

* a Voronoi mesh with its sites at the cell centres of an equidistant Cartesian grid HAS that grid's
  cells, so the oracle -- same random stream, other geometry code -- must land on the reference's
  golden Cartesian outputs (test_specific_energy / test_peeloff, grid_type=car) to rounding;
* on a random mesh an isotropic uniform radiation field deposits the same track length per volume in
  every cell, whatever its shape (a closed form), and every packet stays in the cell its position says
  (in_correct_cell at every step).
"""
import numpy as np
import pytest

from helpers import bitlevel_model, bitlevel_model_vor, peeloff_model, pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource
from oracle import oracle


def _on_lattice_mesh(m):
    mesh = syn.lattice_voronoi(m.w1, m.w2, m.w3)
    v = FlatModel(None, None, None, m.density.reshape(m.density.shape[0], -1), m.dust, m.sources, m.conf,
                  grid_type="vor", voronoi=mesh)
    v.peeled = m.peeled
    return v


@pytest.mark.parametrize("evenly,multi", [(False, False), (True, False), (False, True), (True, True)])
def test_lattice_mesh_reproduces_golden_cartesian_specific_energy(golden_car, evenly, multi):
    z = golden_car
    o = oracle.Oracle(_on_lattice_mesh(bitlevel_model(z, evenly, multi)))
    expected = z["expected_evenly=%s_multi=%s" % (evenly, multi)]
    for it in range(5):
        st = o.run_lucy_iteration(10000)
        got = o.get_specific_energy().reshape(expected[it].shape)
        assert st.killed_geo == 0 and st.killed_int == 0
        # bisecting planes instead of stored walls: path lengths differ in the last bits only
        np.testing.assert_allclose(got, expected[it], rtol=1e-12)


@pytest.mark.parametrize("raytracing", [False, True])
def test_lattice_mesh_reproduces_golden_cartesian_peeloff(golden_car, raytracing):
    """Peel-off marches (grid_escape_tau), forced first interaction and, with raytracing, thermal packets
    from random positions of random cells (random_position_cell: rejection sampling in the bounding box,
    which for these cells never rejects and so draws like the Cartesian routine)."""
    z = golden_car
    o = oracle.Oracle(_on_lattice_mesh(peeloff_model(z, False)))
    for it in range(5):
        o.run_lucy_iteration(1000)
    o.final_begin()
    o.final_photons(5000, peeloff_scattering_only=raytracing)
    st = o.final_finish()
    assert st.killed_geo == 0 and st.killed_int == 0
    if raytracing:
        o.raytracing_photons(2000, 3000)
    for ig in (1, 2, 3):
        for kind, get in (("seds", o.sed), ("images", o.image)):
            expected = z["peeloff_ray=%s_evenly=False_g%d_%s" % (raytracing, ig, kind)]
            got = get(ig - 1)
            assert got.shape == expected.shape
            assert np.array_equal(got == 0, expected == 0), (ig, kind)
            nz = expected != 0
            np.testing.assert_allclose(got[nz], expected[nz], rtol=1e-11)


def _random_mesh(n, seed, box):
    rng = np.random.default_rng(seed)
    sites = np.stack([rng.uniform(box[2 * a], box[2 * a + 1], n) for a in range(3)], axis=1)
    return syn.voronoi_mesh(sites, box)


def test_mesh_helper_is_a_tessellation_of_the_box():
    box = np.array([-pc, pc, -0.8 * pc, 0.9 * pc, -0.7 * pc, pc])
    mesh = _random_mesh(200, 1, box)
    assert np.isclose(mesh["volume"].sum(), np.prod(box[1::2] - box[::2]), rtol=1e-12)
    idx, nb = mesh["sparse_idx"], mesh["sparse_neighs"]
    pairs = {(i, int(j)) for i in range(200) for j in nb[idx[i]:idx[i + 1]] if j >= 0}
    assert all((j, i) in pairs for i, j in pairs)                      # neighbours are mutual
    assert set(np.unique(nb[nb < 0])) == {-1, -2, -3, -4, -5, -6}      # every wall of the box is touched
    assert np.all(mesh["bb_min"] >= box[::2] - 1e-9 * pc) and np.all(mesh["bb_max"] <= box[1::2] + 1e-9 * pc)


def test_uniform_field_gives_every_cell_the_same_track_density():
    """A box emitting inwards by the cosine law fills its interior with a uniform isotropic field:
    track length per volume and packet = 4 / S (S: surface of the emitting box) in EVERY cell inside it.
    propagation_check_frequency = 1 runs in_correct_cell (:274-283) before every step."""
    box = np.array([-pc, pc, -0.8 * pc, 0.9 * pc, -0.7 * pc, pc])
    n = 300
    mesh = _random_mesh(n, 3, box)
    dust = syn.grey_dust(n_temp=10)
    b = tuple(0.98 * box)
    src = FlatSource(type=6, luminosity=lsun, temperature=5000., bounds=b)
    model = FlatModel(None, None, None, np.full((1, n), 1e-30), [dust], [src],
                      FlatConf(propagation_check_frequency=1.0), grid_type="vor", voronoi=mesh)
    o = oracle.Oracle(model)
    o.lucy_begin()
    o.lucy_photons(200000)
    sums = o.get_energy_sum()[0]
    st = o.lucy_finish().as_dict()
    assert st["killed_geo"] == 0 and st["n_escaped"] == 200000
    kappa = float(dust.chi[0] * (1.0 - dust.albedo[0]))
    t = sums / kappa / mesh["volume"] / o.energy_current
    L = np.array(b[1::2]) - np.array(b[::2])
    S = 2 * (L[0] * L[1] + L[1] * L[2] + L[0] * L[2])
    inside = np.all(mesh["bb_min"] > np.array(b[::2]), axis=1) & np.all(mesh["bb_max"] < np.array(b[1::2]), axis=1)
    assert inside.sum() > 80
    rel = t[inside] * S / 4 - 1
    assert abs(rel.mean()) < 0.006, rel.mean()
    assert rel.std() < 0.03 and np.abs(rel).max() < 0.1, (rel.std(), np.abs(rel).max())


def test_scattering_run_keeps_packets_in_their_cells(golden_car):
    """The bit-level dust and sources on a random mesh with the self-check at every step: interactions,
    re-emission and scattering restart flights inside cells (on_wall = no wall), none is lost."""
    model = bitlevel_model_vor(golden_car, False, True)
    model.conf.propagation_check_frequency = 1.0
    o = oracle.Oracle(model)
    st = o.run_lucy_iteration(20000)
    assert st.killed_geo == 0 and st.killed_int == 0 and st.n_absorptions > 1000 and st.n_scatterings > 1000
    assert np.all(o.get_specific_energy() > 0)


def test_masked_cells_and_mrw(golden_car):
    """Cells whose volume the front end marked invalid (-1, hyperion/grid/voronoi_grid.py:468-470) carry no
    dust and are never chosen for thermal emission (geo%mask); the modified random walk stops with the
    reference's message (distance_to_closest_wall, :314-320)."""
    model = bitlevel_model_vor(golden_car, False, False)
    model.voronoi["volume"][[3, 17]] = -1.0
    o = oracle.Oracle(model)
    o.run_lucy_iteration(5000)
    se = o.get_specific_energy()
    assert np.all(o.get_density()[0][[3, 17]] == 0)
    assert np.all(se[0][np.delete(np.arange(se.shape[1]), [3, 17])] > 0)
    model = bitlevel_model_vor(golden_car, False, False)
    model.conf.use_mrw = True
    with pytest.raises(Exception, match="not implemented for Voronoi grid"):
        o = oracle.Oracle(model)
        o.run_lucy_iteration(2000)


def test_source_outside_the_box_and_single_cell(golden_car):
    """find_cell outside the box kills the packet at emission (place_in_cell, :230-241 -> emit, source.f90:177);
    a mesh of one site is the box itself: all six walls, no neighbour."""
    m = bitlevel_model_vor(golden_car, False, False)
    m.sources[0].position = (10 * m.voronoi["box"][1], 0., 0.)
    with pytest.raises(Exception, match="photon was not emitted inside a cell"):
        oracle.Oracle(m).run_lucy_iteration(2000)
    one = bitlevel_model(golden_car, False, False)
    w = [np.array([one.w1[0], one.w1[-1]]), np.array([one.w2[0], one.w2[-1]]), np.array([one.w3[0], one.w3[-1]])]
    rho = np.full((1, 1, 1, 1), float(one.density.mean()))
    car = FlatModel(w[0], w[1], w[2], rho, one.dust, one.sources, one.conf)
    vor = FlatModel(None, None, None, rho.reshape(1, 1), one.dust, one.sources, one.conf, grid_type="vor",
                    voronoi=syn.lattice_voronoi(*w))
    assert sorted(vor.voronoi["sparse_neighs"]) == [-6, -5, -4, -3, -2, -1]
    a, b = oracle.Oracle(car), oracle.Oracle(vor)
    sa, sb = a.run_lucy_iteration(5000), b.run_lucy_iteration(5000)
    assert sa.n_crossings == sb.n_crossings and sa.n_absorptions == sb.n_absorptions
    np.testing.assert_allclose(b.get_specific_energy().ravel(), a.get_specific_energy().ravel(), rtol=1e-12)
