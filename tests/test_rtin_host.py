"""Host side of the drop-in boundary: .rtin parsing, the convergence test, error wording."""
import numpy as np
import pytest

from helpers import bitlevel_model
from hyperion_b200 import rtin, rtin_write, runner
from hyperion_b200.io import h5min, h5write


def test_rtin_roundtrip(golden_car, tmp_path):
    m = bitlevel_model(golden_car, True, True)
    m.minimum_specific_energy = np.array([1.0, 2.0, 3.0])
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=5, n_initial_photons=1e4, check_convergence=(2., 1.02, 99.))
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.n_initial_iter == 5 and rs.n_initial_photons == 10000 and rs.grid_type == "car"
    assert rs.check_convergence and rs.convergence_percentile == 99.
    assert got.conf.sample_sources_evenly and got.conf.seed == -124902
    for k in ("w1", "w2", "w3", "density"):
        assert np.array_equal(getattr(got, k), getattr(m, k))
    assert len(got.dust) == 3 and len(got.sources) == 5
    for k in ("nu", "chi", "albedo", "P2", "emiss_jnu", "jnu_var", "chi_inv_planck"):
        assert np.array_equal(getattr(got.dust[2], k), getattr(m.dust[0], k)), k
    assert got.dust[0].version == m.dust[0].version
    assert np.array_equal(got.minimum_specific_energy, [1., 2., 3.])
    for a, b in zip(got.sources, m.sources):
        assert a.luminosity == b.luminosity and a.temperature == b.temperature and tuple(a.position) == tuple(b.position)


def test_reference_error_phrases(golden_car, tmp_path):
    """hyperion/model/tests/test_fortran.py: the log is searched for these phrases."""
    m = bitlevel_model(golden_car, False, False)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=0)
    with pytest.raises(rtin.ModelError, match="number of specific_energy photons is zero"):
        rtin.read_rtin(fn)
    m.sources[0].temperature = None
    m.sources[0].spectrum_nu = np.array([3e10, 2e10, 1e11])
    m.sources[0].spectrum_fnu = np.ones(3)
    rtin_write.write_rtin(fn, m)
    with pytest.raises(rtin.ModelError, match="spectrum frequency should be monotonically increasing"):
        rtin.read_rtin(fn)
    m.sources.clear()
    rtin_write.write_rtin(fn, m)
    with pytest.raises(rtin.ModelError, match="no sources set up"):
        rtin.read_rtin(fn)


def test_convergence_check_follows_reference():
    """specific_energy_converged (grid_physics_3d.f90:637-689): first call stores, second call
    has no previous value, third may converge; quantile uses nint(p/100*(n-1))+1."""
    q = runner.ConvergenceCheck.quantile
    x = np.arange(1., 11.)
    assert q(x, 100.) == 10. and q(x, 0.) == 1. and q(x, 50.) == 6. and q(x, 99.) == 10.
    c = runner.ConvergenceCheck(absolute=2., relative=1.5, percentile=100.)
    a = np.array([1., 2., 4.])
    assert c(a) is False
    assert c(a * np.array([1.1, 1., 1.])) is False          # value 1.1, no previous value
    assert c(a * np.array([1.1, 1.2, 1.])) is True           # value 1.2, ratio 1.09
    assert c(a * 100.) is False                               # value way above 'absolute'
    same = a * 100.
    assert c(same) is True                                    # exact convergence


@pytest.mark.gpu
def test_runner_end_to_end_writes_reference_layout(golden_car, tmp_path):
    """python -m hyperion_b200 input output: the .rtout has what ModelOutput reads
    (hyperion/model/helpers.py:10 find_last_iteration, model_output.py:975 get_quantities)."""
    m = bitlevel_model(golden_car, False, True)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    gid = rtin_write.write_rtin(fin, m, n_initial_iter=3, n_initial_photons=20000, output_specific_energy="all",
                                copy_input=False)
    assert runner.main([fin, fout]) == 0
    assert runner.main([fin, fout]) == 1                      # exists, no -f
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    assert "date_ended" in r.attrs and "date_started" in r.attrs and "cpu_time" in r.attrs
    assert int(np.asarray(r.attrs["iterations"]).ravel()[0]) == 3
    assert bytes(np.asarray(r.attrs["converged"]).ravel()[0]).strip(b"\x00") == b"no"
    for k in ("killed_photons_geo_final", "killed_photons_int_final", "killed_photons_geo_raytracing",
              "killed_photons_int_raytracing"):
        assert k in r.attrs
    assert isinstance(r.get_link("Input"), h5min.ExternalLink)
    for it in (1, 2, 3):
        g = r["iteration_%05d" % it]
        se = g["specific_energy"]
        assert se.shape == (3, 3, 5, 7) and se.dtype == np.float64
        assert bytes(np.asarray(se.attrs["geometry"]).ravel()[0]).decode().strip("\x00") == gid
        assert int(np.asarray(g.attrs["killed_photons_geo"]).ravel()[0]) == 0
        assert np.all(se[...] > 0)


def test_rtin_roundtrip_peeled_groups(golden_car, tmp_path):
    """Output/Peeled/group_%05i as PeeledImageConf.write lays it out (hyperion/conf/conf_files.py)."""
    from helpers import peeloff_model
    m = peeloff_model(golden_car, False)
    m.peeled[1].track_origin, m.peeled[1].track_n_scat = "scatterings", 3
    m.peeled[2].uncertainties, m.peeled[2].io_bytes, m.peeled[2].stokes = True, 4, False
    m.peeled[0].d_min, m.peeled[0].peeloff_origin = -1e18, (1e17, 2e17, 3e17)
    m.conf.forced_first_interaction_algorithm, m.conf.baes16_xi = "baes16", 0.3
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_last_photons=5000, raytracing=True, n_ray_photons=(2000, 3000))
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.n_last_photons == 5000 and rs.raytracing and rs.n_ray_photons_sources == 2000 and rs.n_ray_photons_dust == 3000
    assert got.conf.forced_first_interaction and got.conf.forced_first_interaction_algorithm == "baes16"
    assert got.conf.baes16_xi == 0.3
    assert len(got.peeled) == 3
    for a, b in zip(got.peeled, m.peeled):
        assert np.array_equal(a.theta, b.theta) and np.array_equal(a.phi, b.phi)
        for k in ("wavelengths", "image", "sed", "track_origin", "track_n_scat", "uncertainties", "stokes", "io_bytes",
                  "inside_observer", "ignore_optical_depth", "d_min", "d_max"):
            assert getattr(a, k) == getattr(b, k), k
        assert tuple(a.peeloff_origin) == tuple(b.peeloff_origin)


@pytest.mark.gpu
def test_runner_writes_peeled_groups(golden_car, tmp_path):
    """End to end through the file boundary: Lucy iterations, imaging iteration, raytracing;
    /Peeled/group_%05d/{seds,images}[_unc] with the attributes ModelOutput.get_sed/get_image read
    (hyperion/model/model_output.py:212,539), in the shapes of the reference's golden .rtout."""
    from helpers import peeloff_model
    m = peeloff_model(golden_car, False)
    m.peeled[0].uncertainties = True
    m.peeled[1].io_bytes = 4
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=20000, n_last_photons=50000, raytracing=True,
                          n_ray_photons=(20000, 30000))
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    assert "date_ended" in r.attrs
    shapes = {1: ((4, 1, 2, 5, 5), (4, 1, 2, 5, 4, 5)), 2: ((4, 4, 1, 2, 4), (4, 4, 1, 6, 6, 4)),
              3: ((4, 12, 1, 2, 4), (4, 12, 1, 6, 6, 4))}
    for ig, (ssed, simg) in shapes.items():
        g = r["Peeled/group_%05d" % ig]
        assert g["seds"].shape == ssed and g["images"].shape == simg
        assert g["seds"].dtype == (np.float32 if ig == 2 else np.float64)
        for k in ("numin", "numax", "apmin", "apmax", "track_origin"):
            assert k in g["seds"].attrs
        for k in ("numin", "numax", "xmin", "xmax", "ymin", "ymax", "track_origin"):
            assert k in g["images"].attrs
        for k in ("inside_observer", "d_min", "d_max"):
            assert k in g.attrs
        sed = g["seds"][...]
        assert np.all(np.isfinite(sed)) and sed[0].sum() > 0
        # apertures are cumulative (image_type.f90:683-686)
        assert np.all(np.diff(sed[0], axis=-2) >= -1e-30)
    assert "seds_unc" in r["Peeled/group_00001"] and "images_unc" in r["Peeled/group_00001"]
    assert "n_sources" in r["Peeled/group_00003/seds"].attrs
    # total flux: with raytracing the direct + thermal light comes from do_raytracing, the scattered
    # light from do_final; the largest aperture of group 2 (basic origin tracking) holds all four
    sed2 = r["Peeled/group_00002/seds"][...]
    # (re-emitted IR light is practically not scattered by this dust: the oracle leaves slice 4 empty too)
    assert (sed2[0, :3, 0, -1, :].sum(axis=-1) > 0).all()


def test_rtin_roundtrip_spherical_grid(golden_car, golden_sph, tmp_path):
    """Grid/Geometry of a 'sph_pol' grid: walls_1 'r', walls_2 't', walls_3 'p'
    (hyperion/grid/spherical_polar_grid.py:371-380, grid_geometry_spherical_3d.f90:111-128)."""
    from helpers import bitlevel_model_sph
    m = bitlevel_model_sph(golden_car, golden_sph, False, True)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m)
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.grid_type == "sph_pol" and got.grid_type == "sph"
    for k in ("w1", "w2", "w3", "density"):
        assert np.array_equal(getattr(got, k), getattr(m, k))
    assert np.allclose(got.volumes().sum(), 4. / 3. * np.pi * m.w1[-1] ** 3, rtol=1e-12)
    m.w2[0] = -0.1
    rtin_write.write_rtin(fn, m)
    with pytest.raises(rtin.ModelError, match="theta walls should be between 0 and pi"):
        rtin.read_rtin(fn)


@pytest.mark.gpu
def test_runner_spherical_grid(golden_car, golden_sph, tmp_path):
    """hyperion_sph: the same binary drives spherical polar grids (scripts/hyperion:44-92 picks the
    back end from grid_type)."""
    from helpers import peeloff_model_sph
    m = peeloff_model_sph(golden_car, golden_sph, False)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=20000, n_last_photons=20000)
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    se = r["iteration_00002/specific_energy"][...]
    assert se.shape == (1, 3, 7, 5) and np.all(se > 0)
    assert r["Peeled/group_00001/seds"][...][0].sum() > 0


def test_rtin_roundtrip_cylindrical_grid(golden_car, golden_cyl, tmp_path):
    """'cyl_pol' grids: walls_1 'w', walls_2 'z', walls_3 'p' (grid_geometry_cylindrical_3d.f90:109-121)."""
    from helpers import bitlevel_model_sph
    m = bitlevel_model_sph(golden_car, golden_cyl, False, False, "cyl")
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m)
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.grid_type == "cyl_pol" and got.grid_type == "cyl"
    for k in ("w1", "w2", "w3", "density"):
        assert np.array_equal(getattr(got, k), getattr(m, k))
    assert np.allclose(got.volumes().sum(), np.pi * m.w1[-1] ** 2 * (m.w2[-1] - m.w2[0]), rtol=1e-12)


def test_rtin_roundtrip_octree(golden_car, golden_oct, tmp_path):
    """'oct' grids: table 'cells' with the depth-first 'refined' flags, root cell attributes x, y, z,
    dx, dy, dz (hyperion/grid/octree_grid.py:426-436); quantities are [n_dust, n_nodes]."""
    from helpers import bitlevel_model_oct
    m = bitlevel_model_oct(golden_car, golden_oct, False, True)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m)
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.grid_type == "oct" and got.grid_type == "oct"
    assert np.array_equal(got.refined, m.refined) and got.oct_half == m.oct_half and got.oct_center == m.oct_center
    assert np.array_equal(got.density, m.density) and got.density.shape == (3, 25)
    v = got.volumes()
    assert np.allclose(v[got.refined == 0].sum(), v[0], rtol=1e-14)


@pytest.mark.gpu
def test_runner_octree(golden_car, golden_oct, tmp_path):
    from helpers import peeloff_model_oct
    m = peeloff_model_oct(golden_car, golden_oct, False)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=20000, n_last_photons=20000, raytracing=True,
                          n_ray_photons=(5000, 5000))
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    se = r["iteration_00002/specific_energy"][...]
    assert se.shape == (1, 25)
    assert np.all(se[0][m.refined == 0] > 0)
    assert r["Peeled/group_00001/seds"][...][0].sum() > 0


def test_rtin_roundtrip_voronoi(golden_car, tmp_path):
    """'vor' grids: table 'cells' (coordinates, volume, bb_min, bb_max), 'sparse_neighs' / 'sparse_idx', the box as
    attributes (hyperion/grid/voronoi_grid.py:417-478); quantities are [n_dust, n_cells]."""
    from helpers import bitlevel_model_vor
    m = bitlevel_model_vor(golden_car, False, True)
    m.voronoi["volume"][5] = 0.0
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m)
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.grid_type == "vor" and got.grid_type == "vor"
    for k in ("coordinates", "bb_min", "bb_max", "sparse_neighs", "sparse_idx", "box"):
        assert np.array_equal(got.voronoi[k], m.voronoi[k]), k
    assert got.voronoi["volume"][5] == -1.0 and np.array_equal(np.delete(got.voronoi["volume"], 5), np.delete(m.voronoi["volume"], 5))
    assert np.array_equal(got.density, m.density) and got.density.shape == (3, 160)


@pytest.mark.gpu
def test_runner_voronoi(golden_car, tmp_path):
    from helpers import peeloff_model_vor
    m = peeloff_model_vor(golden_car, False)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=20000, n_last_photons=20000, raytracing=True,
                          n_ray_photons=(5000, 5000))
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    se = r["iteration_00002/specific_energy"][...]
    assert se.shape == (1, 160) and np.all(se > 0)
    assert r["Peeled/group_00001/seds"][...][0].sum() > 0


def test_rtin_spectrum_bin_edges(golden_car, tmp_path):
    """setup_initial (src/main/setup_rt.f90:78-104): the option, the table of bin edges, and the reference's messages."""
    m = bitlevel_model(golden_car, False, False)
    m.spectrum_bin_edges = np.logspace(6., 18., 13)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, output_specific_energy_spectrum="all")
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.output_specific_energy_spectrum == "all" and np.array_equal(got.spectrum_bin_edges, m.spectrum_bin_edges)
    rtin_write.write_rtin(fn, m)                       # edges given, option off: nothing is computed
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.output_specific_energy_spectrum == "none" and got.spectrum_bin_edges is None
    m.spectrum_bin_edges = None
    rtin_write.write_rtin(fn, m, output_specific_energy_spectrum="last")
    with pytest.raises(rtin.ModelError, match="specific_energy_spectrum_bin_edges should be present in the input"):
        rtin.read_rtin(fn)
    m.spectrum_bin_edges = np.array([1e10, 1e12, 1e11])
    rtin_write.write_rtin(fn, m, output_specific_energy_spectrum="last")
    with pytest.raises(rtin.ModelError, match="should be strictly increasing"):
        rtin.read_rtin(fn)


def test_rtin_roundtrip_amr(golden_car, golden_amr, tmp_path):
    """'amr' grids: Grid/Geometry/level_%05d/grid_%05d attributes, one density dataset per grid
    (hyperion/grid/amr_grid.py:372-412)."""
    from helpers import bitlevel_model_amr
    m = bitlevel_model_amr(golden_car, golden_amr, False, True)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m)
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.grid_type == "amr" and got.grid_type == "amr"
    assert got.amr_levels == m.amr_levels
    assert np.array_equal(got.density, m.density) and got.density.shape == (3, 8 * 6 * 4 + 4 * 6 * 20)


@pytest.mark.gpu
def test_runner_amr(golden_car, golden_amr, tmp_path):
    from helpers import peeloff_model_amr
    m = peeloff_model_amr(golden_car, golden_amr, False)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=20000, n_last_photons=20000)
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    a = r["iteration_00002/level_00001/grid_00001/specific_energy"][...]
    b = r["iteration_00002/level_00002/grid_00001/specific_energy"][...]
    assert a.shape == (1, 4, 6, 8) and b.shape == (1, 20, 6, 4) and np.all(b > 0)
    assert r["Peeled/group_00001/seds"][...][0].sum() > 0


def test_rtin_roundtrip_all_source_types(golden_car, tmp_path):
    """External sphere / box, plane-parallel and point-collection sources in the layout
    hyperion/sources/source.py writes (attrs xmin..zmax, theta/phi, datasets position/luminosity)."""
    from helpers import pc, lsun
    from hyperion_b200.flatmodel import FlatSource
    m = bitlevel_model(golden_car, False, False)
    pts = np.array([[-0.5 * pc, 0.1 * pc, 0.2 * pc], [0.5 * pc, 0.25 * pc, -0.5 * pc]])
    m.sources = [FlatSource(type=5, luminosity=2.0 * lsun, temperature=5000., position=(0.1 * pc, 0., 0.), radius=0.9 * pc),
                 FlatSource(type=6, luminosity=3.0 * lsun, temperature=4000., bounds=(-0.9 * pc, 0.8 * pc, -0.7 * pc, 0.95 * pc, -0.6 * pc, 0.9 * pc)),
                 FlatSource(type=7, luminosity=1.5 * lsun, temperature=7000., position=(0., 0.2 * pc, -0.9 * pc), radius=0.3 * pc,
                            direction=(25.0, 40.0), peeloff=False),
                 FlatSource(type=8, temperature=3000., points=pts, points_luminosity=np.array([1.0, 3.0]) * lsun),
                 FlatSource(type=4, luminosity=lsun, lte=True, map=np.arange(105.).reshape(3, 5, 7))]
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100)
    got, rs, _ = rtin.read_rtin(fn)
    assert [s.type for s in got.sources] == [5, 6, 7, 8, 4]
    a, b, c, d, e = got.sources
    assert e.lte and e.temperature is None and np.array_equal(e.map, np.arange(105.).reshape(3, 5, 7))
    assert a.radius == 0.9 * pc and tuple(a.position) == (0.1 * pc, 0., 0.) and a.peeloff
    assert tuple(b.bounds) == tuple(m.sources[1].bounds) and b.luminosity == 3.0 * lsun
    assert tuple(c.direction) == (25.0, 40.0) and c.radius == 0.3 * pc and not c.peeloff
    assert np.array_equal(d.points, pts) and np.array_equal(d.points_luminosity, np.array([1.0, 3.0]) * lsun)
    assert d.luminosity == 4.0 * lsun
    # the reference's wording for what it rejects
    f = h5min.File(fn)
    del f
    m.sources = [FlatSource(type=1, luminosity=lsun, temperature=5000.)]
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100)


def test_rtin_roundtrip_binned_group(golden_car, tmp_path):
    """Output/Binned/group_00001 (BinnedImageConf): n_theta / n_phi instead of viewing angles."""
    from hyperion_b200.flatmodel import FlatPeeledGroup
    from helpers import pc
    m = bitlevel_model(golden_car, False, False)
    m.conf.forced_first_interaction = False
    m.binned = FlatPeeledGroup(binned=True, n_theta=2, n_phi=3, wavelengths=(4, 0.05, 200.),
                               image=(3, 3, -pc, pc, -pc, pc), sed=(2, 0.5 * pc, 1.8 * pc), track_origin="basic")
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100, n_last_photons=100)
    got, rs, _ = rtin.read_rtin(fn)
    b = got.binned
    assert b is not None and b.binned and (b.n_theta, b.n_phi) == (2, 3)
    assert b.image == m.binned.image and b.sed == m.binned.sed and b.track_origin == "basic"
    assert got.peeled == []
    m.conf.forced_first_interaction = True
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100, n_last_photons=100)
    with pytest.raises(rtin.ModelError, match="can't use binned images with forced first interaction"):
        rtin.read_rtin(fn)


@pytest.mark.gpu
def test_runner_writes_binned_group(golden_car, tmp_path):
    """main.f90:263,326: the binned cubes go into /Binned/{seds,images} of the .rtout
    (ModelOutput.get_sed(group=0) reads them, hyperion/model/model_output.py:349)."""
    from hyperion_b200.flatmodel import FlatPeeledGroup
    from helpers import pc
    m = bitlevel_model(golden_car, False, False)
    m.conf.forced_first_interaction = False
    m.binned = FlatPeeledGroup(binned=True, n_theta=2, n_phi=3, wavelengths=(4, 0.05, 200.),
                               image=(3, 3, -pc, pc, -pc, pc), sed=(2, 0.5 * pc, 1.8 * pc), track_origin="basic")
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=1, n_initial_photons=20000, n_last_photons=50000)
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    assert "date_ended" in r.attrs
    assert r["Binned/seds"].shape == (4, 4, 6, 2, 4) and r["Binned/images"].shape == (4, 4, 6, 3, 3, 4)
    sed = r["Binned/seds"][...]
    assert np.all(np.isfinite(sed)) and (sed[0, 0].sum(axis=(-1, -2)) > 0).all()


def test_rtin_roundtrip_filters_and_inside_observer(golden_car, tmp_path):
    """use_filters / n_filt / filter_%05i tables (hyperion/conf/conf_files.py:862-885) and inside_observer."""
    from helpers import peeloff_model, pc
    m = peeloff_model(golden_car, False)
    nu = np.logspace(13, 15, 10)
    m.peeled[0].filters = [(nu, np.linspace(0.1, 1.0, 10), 3e14), (nu[:5], np.ones(5), 5e13)]
    m.peeled[1].inside_observer = True
    m.peeled[1].peeloff_origin = (0.1 * pc, 0.2 * pc, -0.3 * pc)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100, n_last_photons=100)
    got, rs, _ = rtin.read_rtin(fn)
    f = got.peeled[0].filters
    assert len(f) == 2 and np.array_equal(f[0][0], nu) and np.array_equal(f[1][1], np.ones(5)) and f[0][2] == 3e14
    assert got.peeled[0].wavelengths[0] == 2
    assert got.peeled[1].inside_observer and got.peeled[1].peeloff_origin == (0.1 * pc, 0.2 * pc, -0.3 * pc)
    assert got.peeled[1].filters is None and got.peeled[2].filters is None


def test_rtin_roundtrip_spots(golden_car, tmp_path):
    """Sub-groups of a spherical source are its spots (source_read, src/sources/source_type.f90:150-188)."""
    from helpers import pc, lsun
    from hyperion_b200.flatmodel import FlatSource
    m = bitlevel_model(golden_car, False, False)
    nu = np.logspace(13.5, 15.2, 12)
    m.sources = [FlatSource(type=2, luminosity=lsun, temperature=5000., radius=0.25 * pc,
                            spots=[dict(luminosity=0.8 * lsun, longitude=60., latitude=20., radius=35., temperature=9000.),
                                   dict(luminosity=0.5 * lsun, longitude=140., latitude=250., radius=20.,
                                        spectrum_nu=nu, spectrum_fnu=nu ** -1.5)])]
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100)
    got, rs, _ = rtin.read_rtin(fn)
    sp = got.sources[0].spots
    assert len(sp) == 2 and sp[0]["temperature"] == 9000. and sp[0]["radius"] == 35.
    assert np.array_equal(sp[1]["spectrum_nu"], nu) and sp[1]["longitude"] == 140.
    assert got.sources[0].luminosity == lsun      # the star's own; the engine adds the spots


def test_rtin_roundtrip_map_source_on_amr_grid(golden_car, golden_amr, tmp_path):
    """A luminosity map on an AMR grid is stored per level / grid like the density."""
    from helpers import bitlevel_model_amr, lsun
    from hyperion_b200.flatmodel import FlatSource
    m = bitlevel_model_amr(golden_car, golden_amr, False, False)
    lm = np.arange(float(m.n_cells)) + 1.0
    m.sources = [FlatSource(type=4, luminosity=lsun, temperature=4000., map=lm)]
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=1, n_initial_photons=100)
    got, rs, _ = rtin.read_rtin(fn)
    assert np.array_equal(got.sources[0].map, lm)


def test_model_without_dust_skips_initial_iterations(golden_car, tmp_path, capsys):
    """setup_initial (src/main/setup_rt.f90:164-168): no dust -> warning, n_initial_iter = 0, no thermal
    raytracing; the sources are still imaged.  The reader hands the engine a vacuum."""
    m = bitlevel_model(golden_car, False, False)
    m.dust = []
    m.density = np.zeros((0,) + m.density.shape[1:])
    m.specific_energy = None
    m.minimum_specific_energy = None
    fn = str(tmp_path / "nodust.rtin")
    rtin_write.write_rtin(fn, m, n_initial_iter=3, n_initial_photons=1e4, n_last_photons=1e3)
    got, rs, _ = rtin.read_rtin(fn)
    assert "no dust present, so skipping initial iterations" in capsys.readouterr().out
    assert got.no_dust and rs.n_initial_iter == 0 and rs.n_ray_photons_dust == 0 and rs.n_last_photons == 1000
    assert len(got.dust) == 1 and got.density.shape[0] == 1 and not got.density.any()
    nu = got.dust[0].nu
    assert nu.min() < 1e3 and nu.max() > 1e25      # every source frequency finds an opacity


def test_process_layouts(monkeypatch):
    """hyperion_b200.launch: rank / world / local rank from the launcher's environment
    (scripts/hyperion:65-92 starts `mpirun -n N hyperion_<grid>_mpi`)."""
    from hyperion_b200 import launch
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "PMI_RANK", "PMI_SIZE",
              "SLURM_PROCID", "SLURM_NTASKS", "PMIX_RANK", "PMIX_SIZE", "HYPERION_B200_NGPU"):
        monkeypatch.delenv(k, raising=False)
    assert launch.layout() == (0, 1, 0, None)
    assert launch.layout({"OMPI_COMM_WORLD_RANK": "3", "OMPI_COMM_WORLD_SIZE": "8", "OMPI_COMM_WORLD_LOCAL_RANK": "3"}) == \
        (3, 8, 3, "OMPI_COMM_WORLD_RANK")
    assert launch.layout({"PMI_RANK": "1", "PMI_SIZE": "2"}) == (1, 2, 1, "PMI_RANK")
    assert launch.layout({"SLURM_PROCID": "5", "SLURM_NTASKS": "6", "SLURM_LOCALID": "1"}) == (5, 6, 1, "SLURM_PROCID")
    assert launch.layout({"RANK": "2", "WORLD_SIZE": "4", "LOCAL_RANK": "2", "OMPI_COMM_WORLD_RANK": "0",
                          "OMPI_COMM_WORLD_SIZE": "1"})[:3] == (2, 4, 2)      # torchrun's names win
    with pytest.raises(ValueError):
        launch.layout({"RANK": "4", "WORLD_SIZE": "4"})
    a = launch.rendezvous("/tmp/x.rtout", {})
    assert a == launch.rendezvous("/tmp/x.rtout", {}) and a[0] == "127.0.0.1" and 20000 <= a[1] < 40000
    assert launch.rendezvous("/tmp/x.rtout", {"MASTER_ADDR": "10.0.0.1", "MASTER_PORT": "1234"}) == ("10.0.0.1", 1234)
    assert launch.wanted_gpus({"HYPERION_B200_NGPU": "4"}) == 4 and launch.wanted_gpus({}) == 1


def test_hyperion_mpirun_starts_coordinated_ranks(tmp_path):
    """bin/hyperion_mpirun -n N prog: N processes with RANK / WORLD_SIZE / LOCAL_RANK set; the first failure is
    the exit status and stops the others."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "ranks"
    out.mkdir()
    prog = "import os; open(os.path.join(r'%s', os.environ['RANK']), 'w').write(os.environ['WORLD_SIZE'] + ' ' + " \
           "os.environ['LOCAL_RANK'] + ' ' + os.environ['MASTER_ADDR'] + ' ' + os.environ['MASTER_PORT'])" % out
    rc = subprocess.call([os.path.join(root, "bin", "hyperion_mpirun"), "-n", "3", sys.executable, "-c", prog])
    assert rc == 0
    got = sorted(os.listdir(out))
    assert got == ["0", "1", "2"]
    ports = {open(out / r).read().split()[3] for r in got}
    assert len(ports) == 1 and all(open(out / r).read().split()[:3] == ["3", r, "127.0.0.1"] for r in got)
    rc = subprocess.call([os.path.join(root, "bin", "hyperion_mpirun"), "-n", "2", sys.executable, "-c",
                          "import os, sys, time; time.sleep(0 if os.environ['RANK'] == '1' else 30); sys.exit(7)"])
    assert rc == 7


def _mono_model(golden_car):
    from helpers import bitlevel_model
    from hyperion_b200.flatmodel import FlatPeeledGroup
    m = bitlevel_model(golden_car, False, False)
    m.frequencies = 2.99792458e10 / (np.array([0.45, 2.2, 40., 110.]) * 1e-4)
    m.monochromatic_energy_threshold = 1e-8
    pc = 3.08568025e18
    m.peeled = [FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(2, 1e-3 * pc, 8. * pc),
                                image=(4, 4, -2 * pc, 2 * pc, -2 * pc, 2 * pc), track_origin="basic",
                                inu_min=2, inu_max=4, wavelengths=(3, 1., 1.))]
    return m


def test_rtin_roundtrip_monochromatic(golden_car, tmp_path):
    """monochromatic = yes: /frequencies, monochromatic_energy_threshold, n_last_photons_sources / _dust
    (hyperion/model/model.py:133-137, hyperion/conf/conf_files.py:260-268) and inu_min / inu_max of the image groups
    (conf_files.py:1076-1079)."""
    m = _mono_model(golden_car)
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, n_last_photons_mono=(4000, 6000), raytracing=True, n_ray_photons=(2000, 3000))
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.monochromatic and rs.n_last_photons == 0
    assert (rs.n_last_photons_sources, rs.n_last_photons_dust) == (4000, 6000)
    assert np.array_equal(got.frequencies, m.frequencies) and got.monochromatic_energy_threshold == 1e-8
    p = got.peeled[0]
    assert (p.inu_min, p.inu_max, p.wavelengths[0]) == (2, 4, 3)


@pytest.mark.gpu
def test_runner_monochromatic(golden_car, tmp_path):
    """main.f90:271-272 -> do_final_mono: the .rtout holds nu F_nu at the exact frequencies, the table
    /Peeled/group/frequencies and no numin / numax (image_type.f90:700-705,781-784); same cubes as driving the engine
    directly with the same packet ids."""
    from hyperion_b200.capi import Engine
    m = _mono_model(golden_car)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=20000, n_last_photons_mono=(20000, 30000))
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    g = r["Peeled/group_00001"]
    assert g["seds"].shape == (4, 4, 2, 2, 3) and g["images"].shape == (4, 4, 2, 4, 4, 3)
    assert "numin" not in g["seds"].attrs and "numin" not in g["images"].attrs
    assert np.array_equal(np.asarray(g["frequencies"][...]["nu"]), m.frequencies[1:4])
    sed = g["seds"][...]
    assert np.all(np.isfinite(sed)) and (sed[0, 0, :, -1, :] > 0).all() and (sed[0, 1, :, -1, 1:] > 0).all()
    eng = Engine(0)
    eng.load_model(m)
    for it in (1, 2):
        eng.run_lucy_iteration(20000, it)
    eng.final_begin()
    for inu in (1, 2, 3, 4):
        eng.final_mono_photons(inu, 0, 20000, 20000, 0, 30000, 30000)
    eng.final_finish()
    assert np.allclose(eng.sed(0), sed, rtol=1e-9, atol=1e-300)
    eng.close()


def test_rtin_roundtrip_pda_and_n_photons(golden_car, tmp_path):
    """'pda' (setup_rt.f90:75) and Output/output_n_photons (:273-278)."""
    from helpers import bitlevel_model
    m = bitlevel_model(golden_car, False, False)
    m.conf.use_pda = True
    fn = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fn, m, output_n_photons="last")
    got, rs, _ = rtin.read_rtin(fn)
    assert rs.pda and rs.output_n_photons == "last"


@pytest.mark.gpu
def test_runner_pda_and_n_photons(golden_car, tmp_path):
    """main.f90 with pda = yes and output_n_photons = all: every iteration group holds the n_photons dataset
    (output_grid, grid_generic.f90:40-46), one value per cell, and the run ends normally."""
    from helpers import bitlevel_model
    import copy
    m = bitlevel_model(golden_car, False, False)
    m.conf.use_pda = True
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    kw = dict(n_initial_iter=2, n_initial_photons=3000, output_n_photons="all", output_specific_energy="all")
    # the reference's kmh_lite.hdf5 is a version-1 dust file: refused with the PDA (setup_rt.f90:296-302)
    assert m.dust[0].version == 1
    rtin_write.write_rtin(fin, m, **kw)
    assert runner.main(["-f", fin, fout]) == 1
    d = copy.deepcopy(m.dust[0])
    d.version = 2
    m.dust = [d]
    rtin_write.write_rtin(fin, m, **kw)
    assert runner.main(["-f", fin, fout]) == 0
    r = h5min.File(fout)
    for it in (1, 2):
        g = r["iteration_%05d" % it]
        n = g["n_photons"][...]
        assert n.shape == m.density.shape[1:] and n.dtype == np.int64
        assert n.max() <= 3000 * 1.05 and n.sum() > 3000
        assert g["specific_energy"].shape == m.density.shape
