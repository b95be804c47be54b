"""Multi-GPU runs through the drop-in boundary (needs 2 GPUs; skipped otherwise).

`mpirun -n N hyperion_car_mpi in out` of the reference (scripts/hyperion:65-92) must give N coordinated GPU
ranks and ONE output file.  Packets are keyed by id (counter RNG), so sharding them over ranks changes only the
order of the floating-point additions: specific_energy, SEDs and images of the N-rank run must equal the
1-rank run to rounding (the semantics of mp_collect_physical_arrays / mp_collect_images,
src/mpi/mpi_routines.f90:272-471)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _compare(out1, out2):
    from hyperion_b200.io import h5min
    a, b = h5min.File(out1), h5min.File(out2)
    worst = 0.0
    for path in ["iteration_%05d/specific_energy" % i for i in (1, 2, 3)] + \
            ["Peeled/group_%05d/%s" % (g, k) for g in (1, 2, 3) for k in ("seds", "images")]:
        x, y = a[path][...], b[path][...]
        assert x.shape == y.shape
        nz = (x != 0) | (y != 0)
        rel = np.abs(x[nz] - y[nz]) / np.maximum(np.abs(x[nz]), np.abs(y[nz]))
        # a bin fed by very few packets can differ by a large ULP count only through cancellation in Q/U/V
        tol = 1e-9 if "specific_energy" in path else 1e-6
        assert rel.max() < tol, (path, rel.max())
        worst = max(worst, rel.max())
    for k in ("killed_photons_geo_final", "killed_photons_int_final", "killed_photons_geo_raytracing",
              "killed_photons_int_raytracing", "iterations"):
        assert int(np.asarray(a.attrs[k]).ravel()[0]) == int(np.asarray(b.attrs[k]).ravel()[0]), k
    assert "date_ended" in b.attrs
    return worst


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("launcher", ["hyperion_mpirun", "openmpi_env", "ngpu_env"])
def test_two_rank_run_equals_one_rank_run(golden_car, tmp_path, launcher):
    from helpers import peeloff_model
    from hyperion_b200 import rtin_write
    m = peeloff_model(golden_car, False)
    fin = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fin, m, n_initial_iter=3, n_initial_photons=200000, n_last_photons=100000, raytracing=True,
                          n_ray_photons=(40000, 60000), output_specific_energy="all")
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "HYPERION_B200_NGPU")}
    out1, out2 = str(tmp_path / "one.rtout"), str(tmp_path / "two.rtout")
    car, car_mpi = os.path.join(ROOT, "bin", "hyperion_car"), os.path.join(ROOT, "bin", "hyperion_car_mpi")
    subprocess.check_call([car, "-f", fin, out1], env=env, stdout=subprocess.DEVNULL)
    if launcher == "hyperion_mpirun":
        # ~/.hyperionrc [mpi] command = bin/hyperion_mpirun: what `hyperion -m 2 in out` executes
        subprocess.check_call([os.path.join(ROOT, "bin", "hyperion_mpirun"), "-n", "2", car_mpi, "-f", fin, out2],
                              env=env, stdout=subprocess.DEVNULL)
    elif launcher == "ngpu_env":
        subprocess.check_call([car, "-f", fin, out2], env=dict(env, HYPERION_B200_NGPU="2"), stdout=subprocess.DEVNULL)
    else:
        # what Open MPI's mpirun leaves in the environment of its two processes
        procs = [subprocess.Popen([car_mpi, "-f", fin, out2], stdout=subprocess.DEVNULL,
                                  env=dict(env, OMPI_COMM_WORLD_RANK=str(r), OMPI_COMM_WORLD_SIZE="2",
                                           OMPI_COMM_WORLD_LOCAL_RANK=str(r))) for r in range(2)]
        assert [p.wait(timeout=600) for p in procs] == [0, 0]
    worst = _compare(out1, out2)
    print("1-GPU and 2-GPU runs agree, worst relative difference %.2e" % worst)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_failure_on_one_rank_stops_all(golden_car, tmp_path):
    """error() on one rank ends the job (mpi_abort in the reference): a model whose source lies outside the grid
    fails in the engine of every rank... and a rank-0-only failure (output exists, no -f) must not leave rank 1
    waiting in a collective."""
    from helpers import bitlevel_model
    from hyperion_b200 import rtin_write
    m = bitlevel_model(golden_car, False, False)
    fin, fout = str(tmp_path / "m.rtin"), str(tmp_path / "m.rtout")
    rtin_write.write_rtin(fin, m, n_initial_iter=1, n_initial_photons=20000)
    open(fout, "w").write("occupied")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    p = subprocess.run([os.path.join(ROOT, "bin", "hyperion_mpirun"), "-n", "2", os.path.join(ROOT, "bin", "hyperion_car_mpi"),
                        fin, fout], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert "File exists" in p.stderr
    assert open(fout).read() == "occupied"


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_rank_monochromatic_and_pda_run(golden_car, tmp_path):
    """The round-2 paths through two ranks: the monochromatic final iteration (every packet already scaled, the cubes
    of the ranks add up, iter_final_mono.f90:118,187) with raytracing, and Lucy iterations with the PDA, whose packet
    counts travel with the reduced grid (mpi_routines.f90:303-311).  SEDs and images equal the 1-rank run to rounding;
    n_photons and the PDA cells may differ by a few counts, so the specific energy is
    compared on the cells both runs sampled."""
    import copy
    from helpers import bitlevel_model
    from hyperion_b200 import rtin_write
    from hyperion_b200.flatmodel import FlatPeeledGroup
    from hyperion_b200.io import h5min
    m = bitlevel_model(golden_car, False, False)
    d = copy.deepcopy(m.dust[0])
    d.version = 2
    m.dust = [d]
    m.conf.use_pda = True
    m.frequencies = 2.99792458e10 / (np.array([0.45, 2.2, 40.]) * 1e-4)
    pc = 3.08568025e18
    m.peeled = [FlatPeeledGroup(theta=[30., 110.], phi=[40., 250.], sed=(2, 1e-3 * pc, 8. * pc),
                                image=(4, 4, -2 * pc, 2 * pc, -2 * pc, 2 * pc), track_origin="basic",
                                inu_min=1, inu_max=3, wavelengths=(3, 1., 1.))]
    fin = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=100000, n_last_photons_mono=(40000, 60000),
                          raytracing=True, n_ray_photons=(20000, 30000), output_n_photons="last")
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "HYPERION_B200_NGPU")}
    out1, out2 = str(tmp_path / "one.rtout"), str(tmp_path / "two.rtout")
    car = os.path.join(ROOT, "bin", "hyperion_car")
    subprocess.check_call([car, "-f", fin, out1], env=env, stdout=subprocess.DEVNULL)
    subprocess.check_call([car, "-f", fin, out2], env=dict(env, HYPERION_B200_NGPU="2"), stdout=subprocess.DEVNULL)
    a, b = h5min.File(out1), h5min.File(out2)
    for k in ("seds", "images"):
        x, y = a["Peeled/group_00001/" + k][...], b["Peeled/group_00001/" + k][...]
        nz = (x != 0) | (y != 0)
        # the images start from specific energies that agree to the PDA's tolerance, not to rounding
        assert nz.sum() > 20 and (np.abs(x[nz] - y[nz]) / np.maximum(np.abs(x[nz]), np.abs(y[nz]))).max() < 2e-2, k
    n1, n2 = a["iteration_00002/n_photons"][...], b["iteration_00002/n_photons"][...]
    assert abs(n1.sum() / n2.sum() - 1.) < 1e-2
    e1, e2 = a["iteration_00002/specific_energy"][...][0], b["iteration_00002/specific_energy"][...][0]
    both = (n1 >= 40) & (n2 >= 40)
    assert both.sum() > 20 and np.abs(e1[both] / e2[both] - 1.).max() < 1e-9


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_rank_spectrum_run_on_a_voronoi_mesh(golden_car, tmp_path):
    """The frequency-resolved sums travel behind the scalars of the one reduction per iteration
    (mp_collect_physical_arrays, mpi_routines.f90:292-301; the reference's test_..._mpi_matches_serial,
    test_specific_energy_spectrum.py:373-392), here on a Voronoi mesh through bin/hyperion_vor[_mpi]."""
    from helpers import bitlevel_model_vor
    from hyperion_b200 import rtin_write
    from hyperion_b200.io import h5min
    m = bitlevel_model_vor(golden_car, False, True)
    m.spectrum_bin_edges = np.logspace(6., 18., 13)
    fin = str(tmp_path / "m.rtin")
    rtin_write.write_rtin(fin, m, n_initial_iter=2, n_initial_photons=100000, output_specific_energy="all",
                          output_specific_energy_spectrum="all")
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "HYPERION_B200_NGPU")}
    out1, out2 = str(tmp_path / "one.rtout"), str(tmp_path / "two.rtout")
    subprocess.check_call([os.path.join(ROOT, "bin", "hyperion_vor"), "-f", fin, out1], env=env, stdout=subprocess.DEVNULL)
    subprocess.check_call([os.path.join(ROOT, "bin", "hyperion_mpirun"), "-n", "2", os.path.join(ROOT, "bin", "hyperion_vor_mpi"),
                           "-f", fin, out2], env=env, stdout=subprocess.DEVNULL)
    a, b = h5min.File(out1), h5min.File(out2)
    for it in (1, 2):
        for name in ("specific_energy", "specific_energy_spectrum"):
            x, y = a["iteration_%05d/%s" % (it, name)][...], b["iteration_%05d/%s" % (it, name)][...]
            assert x.shape == y.shape == ((3, 160) if name == "specific_energy" else (12, 3, 160))
            nz = (x != 0) | (y != 0)
            assert (np.abs(x[nz] - y[nz]) / np.maximum(np.abs(x[nz]), np.abs(y[nz]))).max() < 1e-9, (it, name)
        se, se_nu = b["iteration_%05d/specific_energy" % it][...], b["iteration_%05d/specific_energy_spectrum" % it][...]
        heated = se > se.min() * (1 + 1e-6)
        np.testing.assert_allclose(se_nu.sum(axis=0)[heated], se[heated], rtol=1e-9)
