"""GPU parity on Voronoi meshes (csrc/geometry_vor.cuh) against the oracle, which tests/test_oracle_voronoi.py
pins to the reference's golden Cartesian files and to a closed form.  Statistical, as in test_gpu_parity.py."""
import numpy as np
import pytest

from helpers import bitlevel_model, bitlevel_model_vor, peeloff_model_vor, pc, lsun
from hyperion_b200 import synthetic as syn
from hyperion_b200.flatmodel import FlatConf, FlatModel, FlatSource
from test_gpu_parity import _gpu_batches, _oracle_batches, _zscores
from test_gpu_imaging import _compare, _converged_energy, _run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("evenly,multi", [(False, False), (True, True)])
def test_deposits_match_oracle_random_mesh(golden_car, evenly, multi):
    """The bit-level dust and sources on a seeded random mesh of 160 cells (8 to 25 neighbours per cell, cells cut
    by all six walls of the box)."""
    model = bitlevel_model_vor(golden_car, evenly, multi)
    B, N = 16, 100000
    g, gst = _gpu_batches(model, N, B)
    o, ost = _oracle_batches(model, N, B)
    z, ok = _zscores(g, o)
    assert ok.all()
    assert np.abs(z).max() < 5.5, np.abs(z).max()
    assert 0.6 < (z ** 2).mean() < 1.5, (z ** 2).mean()
    for key in ("n_crossings", "n_absorptions", "n_scatterings"):
        a = np.mean([s[key] for s in gst])
        b = np.mean([s[key] for s in ost])
        assert abs(a / b - 1) < 0.01, (key, a, b)
    assert all(s["killed_geo"] == 0 and s["killed_int"] == 0 and s["n_photons"] == N for s in gst)


def test_uniform_field_gives_every_cell_the_same_track_density_on_the_gpu():
    """The closed form of tests/test_oracle_voronoi.py on the device: 4 / S per volume and packet in every cell
    inside a box that emits inwards, on a mesh of 2000 cells."""
    from hyperion_b200.capi import Engine
    box = np.array([-pc, pc, -0.8 * pc, 0.9 * pc, -0.7 * pc, pc])
    n = 2000
    rng = np.random.default_rng(5)
    sites = np.stack([rng.uniform(box[2 * a], box[2 * a + 1], n) for a in range(3)], axis=1)
    mesh = syn.voronoi_mesh(sites, box)
    dust = syn.grey_dust(n_temp=10)
    b = tuple(0.98 * box)
    model = FlatModel(None, None, None, np.full((1, n), 1e-30), [dust],
                      [FlatSource(type=6, luminosity=lsun, temperature=5000., bounds=b)], FlatConf(),
                      grid_type="vor", voronoi=mesh)
    eng = Engine(0)
    eng.load_model(model)
    eng.lucy_begin()
    N = 4000000
    eng.lucy_photons(0, N, 1)
    sums = eng.get_energy_sum()[0]
    st = eng.lucy_finish()
    eng.close()
    assert st.killed_geo == 0 and st.n_escaped == N
    kappa = float(dust.chi[0] * (1.0 - dust.albedo[0]))
    t = sums / kappa / mesh["volume"] / st.energy_emitted      # unit-energy packets: per packet
    L = np.array(b[1::2]) - np.array(b[::2])
    S = 2 * (L[0] * L[1] + L[1] * L[2] + L[0] * L[2])
    inside = np.all(mesh["bb_min"] > np.array(b[::2]), axis=1) & np.all(mesh["bb_max"] < np.array(b[1::2]), axis=1)
    assert inside.sum() > 1000
    rel = t[inside] * S / 4 - 1
    assert abs(rel.mean()) < 0.002, rel.mean()
    assert rel.std() < 0.02 and np.abs(rel).max() < 0.08, (rel.std(), np.abs(rel).max())


def test_lattice_mesh_matches_cartesian_grid_on_the_gpu(golden_car):
    """The mesh whose cells are those of the bit-level Cartesian grid: same packets (counter RNG per packet), other
    geometry code -- the deposit grids agree far inside the Monte-Carlo noise."""
    from hyperion_b200.capi import Engine
    m = bitlevel_model(golden_car, False, True)
    v = FlatModel(None, None, None, m.density.reshape(3, -1), m.dust, m.sources, m.conf, grid_type="vor",
                  voronoi=syn.lattice_voronoi(m.w1, m.w2, m.w3))
    out = []
    for model in (m, v):
        eng = Engine(0)
        eng.load_model(model)
        eng.lucy_begin()
        eng.lucy_photons(0, 200000, 1)
        out.append(eng.get_energy_sum().reshape(3, -1))
        st = eng.lucy_finish()
        assert st.killed_geo == 0
        eng.close()
    np.testing.assert_allclose(out[1], out[0], rtol=2e-3)
    assert abs(out[1].sum() / out[0].sum() - 1) < 1e-4


def test_peeloff_and_raytracing_match_oracle_random_mesh(golden_car):
    """Final iteration with forced first interaction, peel-off marches through the mesh, raytracing with thermal
    packets from random positions of random cells (rejection sampling in the cell's bounding box)."""
    model = peeloff_model_vor(golden_car, False)
    model.specific_energy = _converged_energy(model)
    gpu, orc = _run_both(model, 12, 60000, True, (20000, 30000))
    print(_compare(gpu, orc))
    for key in ("n_crossings", "n_absorptions", "n_scatterings", "n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[1][key] for g in gpu])
        b = np.mean([o[1][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    for key in ("n_peeloffs", "n_peel_crossings"):
        a = np.mean([g[2][key] for g in gpu])
        b = np.mean([o[2][key] for o in orc])
        assert abs(a / b - 1) < 0.02, (key, a, b)
    assert all(g[1]["killed_geo"] == 0 for g in gpu)


def test_mrw_is_refused_with_the_reference_message(golden_car):
    from hyperion_b200.capi import Engine, HyperionError
    model = bitlevel_model_vor(golden_car, False, False)
    model.conf.use_mrw = True
    eng = Engine(0)
    with pytest.raises(HyperionError, match="not implemented for Voronoi grid"):
        eng.load_model(model)
    eng.close()


def test_source_outside_the_box_reports_reference_message(golden_car):
    from hyperion_b200.capi import Engine, HyperionError
    m = bitlevel_model_vor(golden_car, False, False)
    m.sources[0].position = (10 * m.voronoi["box"][1], 0., 0.)
    eng = Engine(0)
    eng.load_model(m)
    with pytest.raises(HyperionError, match="photon was not emitted inside a cell"):
        eng.run_lucy_iteration(20000)
    eng.close()


def test_masked_cells_stay_empty(golden_car):
    """Cells the front end marked invalid (volume -1) hold no dust and take no thermal raytracing packets."""
    from hyperion_b200.capi import Engine
    m = peeloff_model_vor(golden_car, False)
    m.voronoi["volume"][[3, 17]] = -1.0
    eng = Engine(0)
    eng.load_model(m)
    eng.run_lucy_iteration(100000)
    assert np.all(eng.get_density()[0][[3, 17]] == 0)
    eng.final_begin()
    eng.final_photons(0, 20000, True)
    eng.final_finish()
    st = eng.raytracing_photons(5000, 20000)
    assert st.killed_geo == 0 and eng.sed(0).sum() > 0
    eng.close()
