"""Debug: wave engine vs direct kernels on the 64^3 cube: counters and where the sums differ."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hyperion_b200 import synthetic as syn
from hyperion_b200.capi import Engine

model = syn.cartesian_point_source_model(n=64, tau_edge=3.0, dust=syn.realistic_dust(n_temp=40))
N = 500000
res = {}
for engine in ("rounds", "wave"):
    os.environ["HYPERION_B200_ENGINE"] = engine
    os.environ["HYPERION_B200_WAVE_TAIL"] = "0"
    eng = Engine(0)
    eng.load_model(model)
    eng.lucy_begin()
    eng.lucy_photons(0, N, 1)
    sums = eng.get_energy_sum()
    st = eng.lucy_finish().as_dict()
    eng.close()
    res[engine] = (sums[0], st)
a, sa = res["rounds"]
b, sb = res["wave"]
for k in sa:
    print("%-18s %22s %22s" % (k, sa[k], sb[k]))
print("total", a.sum(), b.sum(), b.sum() / a.sum() - 1)
d = b - a
z, y, x = np.indices(a.shape)
r = np.sqrt((x - 31.5) ** 2 + (y - 31.5) ** 2 + (z - 31.5) ** 2)
for lo, hi in [(0, 1), (1, 2), (2, 4), (4, 8), (8, 16), (16, 32), (32, 64)]:
    m = (r >= lo) & (r < hi)
    print("shell %2d-%2d: a %.6e  b-a %.3e  rel %.3e" % (lo, hi, a[m].sum(), d[m].sum(), d[m].sum() / a[m].sum()))
i = np.unravel_index(np.argmax(np.abs(d)), d.shape)
print("largest |b-a| at", i, a[i], b[i], d[i])
idx = np.argsort(-np.abs(d).ravel())[:12]
for j in idx:
    k = np.unravel_index(j, d.shape)
    print(k, "a=%.6e b=%.6e d=%.3e" % (a[k], b[k], d[k]))
# per tile-plane sums of the difference (22-cell tiles)
for ax, name in ((2, "x"), (1, "y"), (0, "z")):
    other = tuple(k for k in range(3) if k != ax)
    prof = d.sum(axis=other)
    print(name, "planes with largest |diff|:", np.argsort(-np.abs(prof))[:8], np.sort(-np.abs(prof))[:3])
