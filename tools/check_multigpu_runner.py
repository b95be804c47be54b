"""2-GPU check of the file-boundary runner: the same .rtin run on 1 GPU and under torchrun on 2 GPUs.

Packets are keyed by id (counter RNG), so sharding them over ranks changes only the order of the
floating-point additions: specific_energy, SEDs and images must agree to rounding.

    python tools/check_multigpu_runner.py            (needs 2 GPUs)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import peeloff_model  # noqa: E402
from hyperion_b200 import rtin_write  # noqa: E402
from hyperion_b200.io import h5min  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_car.npz"))
m = peeloff_model(z, False)
tmp = tempfile.mkdtemp()
fin = os.path.join(tmp, "m.rtin")
rtin_write.write_rtin(fin, m, n_initial_iter=3, n_initial_photons=200000, n_last_photons=100000, raytracing=True,
                      n_ray_photons=(40000, 60000), output_specific_energy="all")
env = dict(os.environ, PYTHONPATH=ROOT)
out1, out2 = os.path.join(tmp, "one.rtout"), os.path.join(tmp, "two.rtout")
subprocess.check_call([sys.executable, "-m", "hyperion_b200", "-f", fin, out1], env=env, stdout=subprocess.DEVNULL)
subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                       "--master-addr", "127.0.0.1", "--master-port", "29533", "-m", "hyperion_b200", "-f", fin, out2],
                      env=env, stdout=subprocess.DEVNULL)
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
    print("%-40s max rel diff %.2e  (%d values)" % (path, rel.max(), nz.sum()))
    assert rel.max() < tol, path
    worst = max(worst, rel.max())
print("OK: 1-GPU and 2-GPU runs agree, worst relative difference %.2e" % worst)
