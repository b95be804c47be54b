"""Small driver for ncu: one imaging iteration (do_final + peel-off) on a synthetic Cartesian or
spherical grid.

    ncu --set full --clock-control none --import-source on -k regex:peel_kernel -s 2 -c 1 \
        -o gpurun_out/prof python tools/profile_final.py --grid 256 --photons 1e6
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hyperion_b200 import synthetic as syn  # noqa: E402
from hyperion_b200.capi import Engine  # noqa: E402
from hyperion_b200.flatmodel import FlatPeeledGroup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--photons", type=float, default=1e6)
ap.add_argument("--tau", type=float, default=1.0)
ap.add_argument("--geometry", default="car", choices=["car", "sph"])
ap.add_argument("--lucy", type=float, default=0, help="also run a Lucy iteration with this many packets")
a = ap.parse_args()
if a.geometry == "car":
    model = syn.cartesian_point_source_model(n=a.grid, tau_edge=a.tau, dust=syn.realistic_dust(n_temp=200))
else:
    model = syn.spherical_disk_model(n_r=a.grid, n_theta=a.grid // 2, tau_edge=a.tau, dust=syn.realistic_dust(n_temp=200))
half = float(model.w1[-1])
model.peeled = [FlatPeeledGroup(theta=[30., 60., 90., 140.], phi=[10., 80., 200., 300.], wavelengths=(50, 0.1, 1000.),
                                image=(256, 256, -half, half, -half, half), sed=(1, 2 * half, 2 * half))]
eng = Engine(0)
eng.load_model(model)
if a.lucy:
    st = eng.run_lucy_iteration(int(a.lucy), 1)
    print("lucy: %.3f ms, %.3e packets/s, %.1f crossings/packet, alg GB/s %.1f" % (
        st.kernel_ms, a.lucy / (st.kernel_ms * 1e-3), st.n_crossings / a.lucy,
        24 * st.n_crossings / (st.kernel_ms * 1e-3) / 1e9))
eng.final_begin()
eng.final_photons(0, int(a.photons), False)
st = eng.final_finish()
cr = st.n_crossings + st.n_peel_crossings
print("final: %.3f ms (%d rounds), %.3e packets/s, %.1f peel-offs/packet, %.1f crossings/packet, %.1f GB/s at 8 B/crossing" % (
    st.kernel_ms, st.n_rounds, a.photons / (st.kernel_ms * 1e-3), st.n_peeloffs / a.photons, cr / a.photons,
    8 * cr / (st.kernel_ms * 1e-3) / 1e9))
print("sed total", float(eng.sed(0)[0].sum()))
eng.close()
