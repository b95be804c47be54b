"""One BASELINE.json configuration (hyperion_b200/workloads.py) with per-iteration engine statistics.

    HYPERION_B200_TIMING=1 python tools/profile_config.py c3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hyperion_b200 import workloads  # noqa: E402
from hyperion_b200.capi import Engine  # noqa: E402

name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m, info = workloads.build(name)
eng = Engine(0)
eng.load_model(m)
for it in range(iters):
    st = eng.run_lucy_iteration(info["photons"], iteration=it + 1)
    print("%s iter %d: %.2f ms photon loop (%.2f ms flights), %d rounds, %d launches, %.1f crossings/packet, %.2f interactions/packet, "
          "killed %d" % (name, it + 1, st.kernel_ms, st.flight_ms, st.n_rounds, st.n_launches, st.n_crossings / info["photons"],
                         (st.n_absorptions + st.n_scatterings) / info["photons"], st.killed_int))
eng.close()
