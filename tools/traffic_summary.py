"""Turn an ncu launch list (gpu__time_duration.sum + dram__bytes_read.sum + dram__bytes_write.sum per launch,
CSV) of `bench.py --steps K --warmup W ...` into profiles/*_traffic.json: DRAM bytes the flight kernels move
per step and every kernel's share of the kernel time.

    python tools/traffic_summary.py gpurun_out/r02m_launches.csv STEPS_INCLUDING_WARMUP PHOTONS > profiles/r02m_traffic.json
"""
import collections
import csv
import json
import sys

fn, n_steps, photons = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3]))
alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
rows = [r for r in csv.reader(open(fn)) if len(r) > 5]
hdr = None
t = collections.defaultdict(float)
by = collections.defaultdict(float)
n = collections.defaultdict(int)
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
    metric = r[hdr.index("Metric Name")]
    unit = r[hdr.index("Metric Unit")]
    val = float(r[hdr.index("Metric Value")].replace(",", ""))
    if metric == "gpu__time_duration.sum":
        t[name] += val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        n[name] += 1
    elif metric.startswith("dram__bytes"):
        by[name] += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
flight = [k for k in t if k in ("wave_tile_kernel", "flight_kernel", "flight_beam_kernel", "flight_geo_kernel")]
tot = sum(t.values())
out = {
    "workload": "cartesian_256^3_point_source_6000K_isotropic_dust_tau1",
    "photons_per_gpu_per_step": photons,
    "flight_dram_bytes_per_step": sum(by[k] for k in flight) / n_steps,
    "algorithmic_bytes_per_step": alg,
    "source": "%s: sum of dram__bytes_read.sum + dram__bytes_write.sum over the flight kernels (%s) of %d steps "
              "(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; "
              "ncu serialises the streams), divided by %d" % (fn, ", ".join(sorted(flight)), n_steps, n_steps),
    "per_kernel_GB_per_step": {k: by[k] / n_steps / 1e9 for k in sorted(by, key=lambda k: -by[k]) if by[k] > 0},
    "kernel_ms_per_step_under_ncu": {k: t[k] / n_steps for k in sorted(t, key=lambda k: -t[k])},
    "kernel_time_share_under_ncu": {k: t[k] / tot for k in sorted(t, key=lambda k: -t[k])},
    "launches_per_step": {k: n[k] / n_steps for k in sorted(n, key=lambda k: -t[k])},
}
print(json.dumps(out, indent=1))
