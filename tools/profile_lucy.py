"""Small driver for ncu: one warm-up and a few Lucy iterations on a synthetic grid.

    ncu --set full --clock-control none --import-source on -k regex:lucy_photon -s 1 -c 1 \
        -o gpurun_out/prof python tools/profile_lucy.py --grid 256 --photons 2e6
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hyperion_b200 import synthetic as syn  # noqa: E402
from hyperion_b200.capi import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--photons", type=float, default=2e6)
ap.add_argument("--tau", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--n-temp", type=int, default=1200)
a = ap.parse_args()
model = syn.cartesian_point_source_model(n=a.grid, tau_edge=a.tau, dust=syn.realistic_dust(n_temp=a.n_temp))
eng = Engine(0)
eng.load_model(model)
for it in range(a.iters):
    st = eng.run_lucy_iteration(int(a.photons), iteration=it + 1)
    alg = 24 * st.n_crossings + 12 * st.n_absorptions
    print("iter %d: %.3f ms photon loop (%.3f ms flight kernels, %d rounds, %d on the wave engine), %.3e packets/s, "
          "%.1f crossings/packet, alg GB/s %.1f (flight kernels alone %.1f)" % (
              it + 1, st.kernel_ms, st.flight_ms, st.n_rounds, st.n_wave_rounds, a.photons / (st.kernel_ms * 1e-3),
              st.n_crossings / a.photons, alg / (st.kernel_ms * 1e-3) / 1e9, alg / (st.flight_ms * 1e-3) / 1e9))
eng.close()
