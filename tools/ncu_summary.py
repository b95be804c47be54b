"""Summarise an ncu report (raw metrics + SASS hot spots) into text for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r1_lucy.txt
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "sm__cycles_elapsed.avg", "smsp__pcsamp_sample_buffer_full"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:100])
    for i, h in enumerate(hdr):
        if h in KEYS:
            print("  %-70s %14s %s" % (h, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h2 = rd[0]
ix = {h: i for i, h in enumerate(h2)}
stall_cols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
recs = []
for r in rd[1:]:
    if len(r) != len(h2):
        continue
    try:
        ns = int(r[ix["# Samples"]])
    except ValueError:
        continue
    for s in stall_cols:
        tot[s] += int(r[ix[s]] or 0)
    recs.append((ns, r))
total = sum(n for n, _ in recs) or 1
print("\n== stall reasons (all samples, %d total)" % total)
for s, n in tot.most_common(10):
    print("  %-28s %6.2f %%" % (s, 100.0 * n / total))
print("\n== top %d SASS instructions by samples (cumulative %% | instr executed | avg threads)" % top)
cum = 0
for ns, r in sorted(recs, key=lambda x: -x[0])[:top]:
    cum += ns
    why = max(stall_cols, key=lambda s: int(r[ix[s]] or 0))
    print("  %5.2f%% %5.1f%% %12s %5s  %-18s %s" % (100.0 * ns / total, 100.0 * cum / total,
          r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]], why, r[ix["Source"]].strip()[:70]))
