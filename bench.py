#!/usr/bin/env python
"""Benchmark of the hot path: photon packets/sec for one Lucy iteration on a 256^3
Cartesian grid (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one Lucy iteration (src/main/iter_lucy.f90:66-237) over a fixed batch of
synthetic photon packets per GPU (weak scaling): reset + jnu_var precompute, the photon
kernel, the all-reduce of the deposit grid (N > 1), the scale/clamp epilogue.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is derived.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hyperion_b200 import synthetic as syn  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="cells per axis")
    ap.add_argument("--photons", type=float, default=2.0e7, help="packets per GPU per step")
    ap.add_argument("--tau", type=float, default=1.0, help="centre-to-face optical depth at 0.5 micron")
    ap.add_argument("--cpu-photons", type=float, default=0, help="packets for the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-thin", action="store_true", help="skip the optically thin variant")
    ap.add_argument("--thin-tau", type=float, default=0.01)
    ap.add_argument("--no-moderate", action="store_true", help="skip the tau = 5 variant")
    ap.add_argument("--moderate-tau", type=float, default=5.0)
    ap.add_argument("--no-imaging", action="store_true", help="skip the imaging-iteration variant")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE.json configurations c1, c3, c4, c5")
    ap.add_argument("--workload", default=None, choices=["c1", "c3", "c4", "c5", "tau5"],
                    help="run ONE configuration of hyperion_b200/workloads.py as the measured workload of the line")
    ap.add_argument("--imaging-photons", type=float, default=2.0e6)
    return ap.parse_args()


def workload_name(a):
    return "cartesian_%d^3_point_source_6000K_isotropic_dust_tau%.3g" % (a.grid, a.tau)


def build_model(a):
    dust = syn.realistic_dust(n_temp=1200)
    return syn.cartesian_point_source_model(n=a.grid, tau_edge=a.tau, dust=dust, seed=1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clock/throttle sampling during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_contexts(model, cores):
    """`cores` emulated reference processes (oracle/), built ONCE: the set-up of a context (1200-state
    emissivity tables, 256^3 arrays) is not part of the photon loop and takes longer than a sample."""
    from oracle import oracle
    oracle.build()
    t0 = time.time()
    ranks = [oracle.Oracle(model, rank=r) for r in range(cores)]
    return ranks, time.time() - t0


def cpu_run(ranks, n_photons):
    """Time the CPU restatement of the Fortran path (oracle/) the way the reference runs under
    MPI: rank r seeded seed+r, equal split, deposit grids summed (src/mpi/mpi_routines.f90)."""
    from concurrent.futures import ThreadPoolExecutor
    cores = len(ranks)
    split = [n_photons // cores + (1 if r < n_photons % cores else 0) for r in range(cores)]
    for o in ranks:
        o.lucy_begin()
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=cores) as pool:
        list(pool.map(lambda x: x[0].lucy_photons(x[1]), zip(ranks, split)))
    dt = time.time() - t0
    cross = 0
    e_cur = sum(o.energy_current for o in ranks)
    for o in ranks:
        o.energy_current = e_cur
        cross += o.lucy_finish().n_crossings
    return n_photons / dt, dt, cross


def oracle_flags():
    try:
        with open(os.path.join(ROOT, "oracle", "Makefile")) as f:
            for line in f:
                if line.startswith("CXXFLAGS"):
                    return line.split("=", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


REFERENCE_BUDGET_S = 240.0     # wall-clock cap of the --impl reference arm (set-up included)


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.time()
    cores = os.cpu_count() or 1
    model = build_model(a)
    ranks, t_setup = cpu_contexts(model, cores)
    n = int(a.cpu_photons) if a.cpu_photons else 0
    if n == 0:
        rate, dt, _ = cpu_run(ranks, 20000 * cores)
        # steps of about (budget left) / (steps + warm-up), never more than 15 s each
        left = max(20.0, REFERENCE_BUDGET_S - (time.time() - t_start))
        per_step = min(15.0, left / (a.warmup + a.steps + 1))
        n = int(max(2000 * cores, min(rate * per_step, 5e7)))
    times, cross = [], 0
    for i in range(a.warmup + a.steps):
        rate, dt, cr = cpu_run(ranks, n)
        if i >= a.warmup:
            times.append(dt)
            cross += cr
        if time.time() - t_start > REFERENCE_BUDGET_S and times:
            break
    steps = len(times)
    T = sum(times)
    value = n * steps / T
    line = {
        "impl": "reference", "metric": "photon_packets_per_sec", "value": value, "unit": "packets/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": a.warmup, "ms_per_step": 1e3 * T / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "grid": [a.grid] * 3, "n_dust": 1,
                   "photons_per_step": n, "note": "bounded CPU sample of the same model"},
        "cpu_baseline": {"value": value, "unit": "packets/s", "cores": cores, "kind": "port",
                         "sample": "%d packets/step x %d steps on %d threads (rank r seeded seed+r, grids summed); "
                                   "contexts built once in %.1f s, not timed" % (n, steps, cores, t_setup),
                         "compile_flags": oracle_flags()},
        "e2e": {"value": value, "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "crossings_per_packet": cross / (n * steps),
        "wall_s": time.time() - t_start,
    }
    print(json.dumps(line), flush=True)


def measure(eng, drv, stream, model, P, world, rank, a, sync_all, e2e=True):
    """Warm-up + K timed Lucy iterations of P packets per GPU; returns (total_ms, stats, e2e dict)."""
    import torch
    import torch.distributed as dist

    def step(it):
        # every step draws fresh packet ids (it * world * P is the id offset of the step)
        return drv.iteration(world * P, it + 1, id_offset=it * world * P)

    for it in range(a.warmup):
        step(it)
    sync_all()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stats = []
    for it in range(a.steps):
        stats.append(step(a.warmup + it))
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())

    out = None
    if e2e:
        # end to end through the public API with HOST buffers: density in (pinned H2D), one
        # iteration, specific_energy out (D2H).  With N > 1 rank 0 owns the host side as in the
        # reference (src/mpi/mpi_io.f90:213-242): it uploads, NCCL broadcasts the device buffer,
        # every rank installs it from device memory, and only rank 0 reads the result back.
        n_el = eng.n_dust * eng.n_cells
        out_np = None
        h_rho = None
        if rank == 0:
            h_rho = torch.empty(n_el, dtype=torch.float64).pin_memory()
            h_rho.numpy()[:] = model.density.ravel()
            h_out = torch.empty(n_el, dtype=torch.float64).pin_memory()
            out_np = h_out.numpy().reshape((eng.n_dust,) + tuple(eng.shape))
        d_rho = torch.empty(n_el, dtype=torch.float64, device="cuda") if world > 1 else None

        def e2e_step(it):
            if world > 1:
                with torch.cuda.stream(stream):
                    if rank == 0:
                        d_rho.copy_(h_rho, non_blocking=True)
                    dist.broadcast(d_rho, src=0)
                stream.synchronize()
                eng.update_density_device(d_rho.data_ptr())
            else:
                eng.update_density(h_rho.numpy())
            step(it)
            if rank == 0:
                eng.get_specific_energy(out_np)

        e2e_step(10_000)
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        k2 = max(1, min(a.steps, 3))
        t0.record(stream)
        for it in range(k2):
            e2e_step(20_000 + it)
        t1.record(stream)
        sync_all()
        ms2 = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        out = {"value": P * world * k2 / (float(ms2.item()) * 1e-3), "unit": "packets/s",
               "h2d_bytes_per_step": int(n_el * 8), "d2h_bytes_per_step": int(n_el * 8), "steps": k2,
               "host_io": "rank 0 (upload + NCCL broadcast, download on rank 0)" if world > 1 else "rank 0"}
    return total_ms, stats, out


def roofline_of(stats, n_dust, world, steps):
    """Algorithmic bytes (24 B per crossing per dust type + 12 B per absorption)
    over the device time of the flight kernels, summed over the rounds of each step."""
    div = world if world > 1 else 1     # with N > 1 the counters were all-reduced with the grid
    cross = sum(s.n_crossings for s in stats) / div
    nabs = sum(s.n_absorptions for s in stats) / div
    nscat = sum(s.n_scatterings for s in stats) / div
    flight_ms = sum(s.flight_ms for s in stats)
    loop_ms = sum(s.kernel_ms for s in stats)
    alg_bytes = 24.0 * n_dust * cross + 12.0 * nabs
    peak, peak_kind = peaks()
    achieved = alg_bytes / (flight_ms * 1e-3) / 1e9 if flight_ms > 0 else 0.0
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "algorithmic_bytes_per_step": alg_bytes / steps,
            "peak_kind": peak_kind, "bytes_per_crossing": 24 * n_dust,
            "kernel": "wave_tile_kernel (all rounds of a step; + flight_kernel for the tail)"
                      if any(getattr(st, "n_wave_rounds", 0) for st in stats)
                      else "flight_beam_kernel + flight_kernel (all rounds of a step)",
            "kernel_ms_per_step": flight_ms / steps, "photon_loop_ms_per_step": loop_ms / steps,
            "achieved_whole_photon_loop": alg_bytes / (loop_ms * 1e-3) / 1e9 if loop_ms > 0 else 0.0,
            "crossings_per_step": cross / steps}, cross, nabs, nscat


def measured_traffic(workload, photons):
    """DRAM bytes (ncu dram__bytes_read.sum + dram__bytes_write.sum) the flight kernels move per step of
    this workload, from the newest committed capture profiles/*_traffic.json; None if no capture matches.
    Like `achieved` it is summed over all flight launches of one step."""
    import glob
    for fn in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            with open(fn) as f:
                t = json.load(f)
        except (OSError, ValueError):
            continue
        if t.get("workload") == workload and int(t.get("photons_per_gpu_per_step", 0)) == int(photons):
            return float(t["flight_dram_bytes_per_step"]), os.path.relpath(fn, ROOT)
    return None, None


def run_config(name, make, Engine, local, rank, world, sync_all):
    """One BASELINE.json configuration (hyperion_b200/workloads.py) on this rank's GPU: per-GPU photon counts,
    packets sharded by id, one all-reduce per Lucy iteration and one of the image cubes.  Returns the
    other_workloads row."""
    import torch
    from hyperion_b200 import workloads as wl
    from hyperion_b200.multigpu import shard
    model, plan = wl.build(name)
    peak, _ = peaks()
    row = {"workload": plan["description"], "config": name, "unit": "packets/s"}
    P = plan["photons"]
    eng, drv, stream = make(model)
    if "lucy" in plan["kind"]:
        drv.iteration(world * P, 1)                       # warm-up (allocations, first-touch)
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stats = [drv.iteration(world * P, 2 + it, id_offset=(1 + it) * world * P) for it in range(plan["iterations"])]
        e1.record(stream)
        sync_all()
        ms = e0.elapsed_time(e1)
        r, cross, _, _ = roofline_of(stats, eng.n_dust, world, len(stats))
        row.update({"value": P * world * len(stats) / (ms * 1e-3), "ms_per_step": ms / len(stats),
                    "photons_per_gpu_per_step": P, "lucy_iterations": len(stats),
                    "crossings_per_packet": cross / (P * len(stats)), "bytes_per_crossing": 24 * eng.n_dust,
                    "roofline_achieved": r["achieved"], "roofline_frac": r["frac"],
                    "roofline_achieved_whole_photon_loop": r["achieved_whole_photon_loop"],
                    "wave_engine_rounds": int(sum(getattr(st, "n_wave_rounds", 0) for st in stats))})
    if "final" in plan["kind"]:
        Pf = plan.get("final_photons", P)
        st = None
        for rep in range(2):       # first pass is the warm-up
            sync_all()
            t0 = time.time()
            eng.final_begin()
            first, count = shard(world * Pf, rank, world)
            eng.final_photons(rep * world * Pf + first, count, False)
            st = eng.final_finish()
            sync_all()
            wall = time.time() - t0
        cr = st.n_crossings + st.n_peel_crossings - st.n_peel_cached
        key = "final_" if "lucy" in plan["kind"] else ""
        rate_ms = st.kernel_ms if world == 1 else wall * 1e3     # N > 1: wall clock between two barriers
        row.update({key + "value": Pf * world / (rate_ms * 1e-3),
                    key + "ms_per_step": rate_ms, key + "photons_per_gpu_per_step": Pf,
                    key + "peeloffs_per_packet": st.n_peeloffs / Pf, key + "crossings_per_packet": cr / Pf,
                    key + "bytes_per_crossing": 8,
                    key + "roofline_achieved": 8.0 * cr / (st.kernel_ms * 1e-3) / 1e9,
                    key + "roofline_frac": 8.0 * cr / (st.kernel_ms * 1e-3) / 1e9 / peak})
    eng.close()
    return row


def multi_gpu_check(make, Engine, local, rank, world, sync_all):
    """32^3 model, 2e6 packets, two iterations: specific_energy of the N-rank run (id shards + one all-reduce per
    iteration) against a 1-rank replay of the same ids on rank 0.

    * direct kernels (HYPERION_B200_ENGINE=rounds, fp64 atomics): the two runs differ only in the order of
      floating-point additions -> 1e-9;
    * wave engine (the product path: 32-bit fixed-point sums per tile visit): a packet's deposits are rounded per
      visit, and which visits the tail hand-off takes over depends on how many packets a rank holds -> the runs
      agree to the fixed-point resolution: total energy to 1e-6, cells to 2e-3.
    Returns "ok" or a description of the mismatch."""
    import torch
    import torch.distributed as dist
    m = syn.cartesian_point_source_model(n=32, tau_edge=2.0, dust=syn.realistic_dust(n_temp=60), seed=3)
    n = 2_000_000
    saved = os.environ.get("HYPERION_B200_ENGINE")
    problems = []
    for engine, tol_cell, tol_total in (("rounds", 1e-9, 1e-9), ("wave", 2e-3, 1e-6)):
        os.environ["HYPERION_B200_ENGINE"] = engine
        eng, drv, _ = make(m)
        for it in range(2):
            drv.iteration(n, it + 1, id_offset=it * n)
        got = eng.get_specific_energy()
        eng.close()
        verdict = torch.zeros(2, dtype=torch.float64, device="cuda")
        if rank == 0:
            one = Engine(local)
            one.load_model(m)
            for it in range(2):
                one.lucy_begin()
                one.lucy_photons(it * n, n, it + 1)
                one.lucy_finish()
            ref = one.get_specific_energy()
            one.close()
            verdict[0] = float(np.max(np.abs(got / ref - 1.0)))
            verdict[1] = abs(float(got.sum() / ref.sum()) - 1.0)
        dist.broadcast(verdict, src=0)
        sync_all()
        cell, total = float(verdict[0].item()), float(verdict[1].item())
        if not (cell <= tol_cell and total <= tol_total):
            problems.append("%s engine: max relative difference %.3e per cell, %.3e in total" % (engine, cell, total))
    if saved is None:
        os.environ.pop("HYPERION_B200_ENGINE", None)
    else:
        os.environ["HYPERION_B200_ENGINE"] = saved
    return "ok" if not problems else "; ".join(problems)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from hyperion_b200.capi import Engine
    from hyperion_b200.multigpu import ShardedLucy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries the JSON line only: NCCL prints its version banner (and NCCL_DEBUG output) to fd 1
        # when the communicator is created, so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            if rank == 0:
                ge.build()
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    elif rank == 0:
        ge.build()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(model):
        eng = Engine(local)
        eng.load_model(model)
        stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

        def all_reduce(buf):
            with torch.cuda.stream(stream):
                dist.all_reduce(buf)
            stream.synchronize()

        return eng, ShardedLucy(eng, rank, world, all_reduce if world > 1 else None), stream

    if a.workload:
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        row = run_config(a.workload, make, Engine, local, rank, world, sync_all)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            key = "" if "value" in row else "final_"
            line = {"metric": "photon_packets_per_sec", "value": row[key + "value"], "unit": "packets/s", "n_gpus": world,
                    "steps": row.get("lucy_iterations", 1), "warmup": 1, "ms_per_step": row[key + "ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": row["workload"], "photons_per_gpu_per_step": row[key + "photons_per_gpu_per_step"]},
                    "roofline": {"bound": "hbm", "achieved": row[key + "roofline_achieved"], "peak": peaks()[0], "unit": "GB/s",
                                 "frac": row[key + "roofline_frac"], "traffic": None,
                                 "bytes_per_crossing": row[key + "bytes_per_crossing"]},
                    "details": row, "clocks": clocks}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    model = build_model(a)
    P = int(a.photons)
    eng, drv, stream = make(model)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, stats, e2e = measure(eng, drv, stream, model, P, world, rank, a, sync_all, e2e=not a.no_e2e)
    clocks = sampler.stop() if rank == 0 else None
    roof, cross, nabs, nscat = roofline_of(stats, eng.n_dust, world, a.steps)
    roof["traffic"], roof["traffic_source"] = measured_traffic(workload_name(a), a.photons)
    gpu_launches = int(sum(s.n_launches for s in stats))
    n_dust, n_cells = eng.n_dust, eng.n_cells
    eng.close()

    # the optically thin and the moderate (tau = 5) variants of the same grid
    def variant(tau):
        a2 = argparse.Namespace(**vars(a))
        a2.tau = tau
        m2 = build_model(a2)
        eng2, drv2, stream2 = make(m2)
        ms2, st2, _ = measure(eng2, drv2, stream2, m2, P, world, rank, a2, sync_all, e2e=False)
        r2, c2, _, _ = roofline_of(st2, eng2.n_dust, world, a.steps)
        eng2.close()
        return {"workload": workload_name(a2), "value": P * world * a.steps / (ms2 * 1e-3), "unit": "packets/s",
                "ms_per_step": ms2 / a.steps, "roofline_achieved": r2["achieved"], "roofline_frac": r2["frac"],
                "roofline_achieved_whole_photon_loop": r2["achieved_whole_photon_loop"],
                "crossings_per_packet": c2 / (P * a.steps)}

    thin = variant(a.thin_tau) if not a.no_thin else None
    moderate = variant(a.moderate_tau) if not a.no_moderate else None

    # the imaging iteration on the same grid (do_final + peel-off): every
    # emission and interaction is peeled towards 4 observers; unit = cell crossings at 8 B each
    imaging = None
    if not a.no_imaging:
        from hyperion_b200.flatmodel import FlatPeeledGroup
        m3 = build_model(a)
        half = float(m3.w1[-1])
        m3.peeled = [FlatPeeledGroup(theta=[30., 60., 90., 140.], phi=[10., 80., 200., 300.],
                                     wavelengths=(50, 0.1, 1000.), image=(256, 256, -half, half, -half, half),
                                     sed=(1, 2 * half, 2 * half), stokes=True)]
        eng3 = Engine(local)
        eng3.load_model(m3)
        Pi = int(a.imaging_photons)
        st3 = None
        for rep in range(2):       # first pass is the warm-up
            sync_all()
            t0 = time.time()
            eng3.final_begin()
            eng3.final_photons((rank + world * rep) * Pi, Pi, False)
            st3 = eng3.final_finish()
            sync_all()
            wall = time.time() - t0
        # crossings actually marched: peel-offs served by the point-source column cache read nothing
        cr = st3.n_crossings + st3.n_peel_crossings - st3.n_peel_cached
        peak, _ = peaks()
        imaging = {"workload": "imaging_iteration_" + workload_name(a) + "_4_views_256x256x50",
                   "value": Pi / (st3.kernel_ms * 1e-3), "unit": "packets/s per GPU", "ms_per_step": st3.kernel_ms,
                   "wall_ms": wall * 1e3, "peeloffs_per_packet": st3.n_peeloffs / Pi,
                   "crossings_per_packet": cr / Pi, "peel_crossings_cached_per_packet": st3.n_peel_cached / Pi,
                   "bytes_per_crossing": 8,
                   "roofline_achieved": 8.0 * cr / (st3.kernel_ms * 1e-3) / 1e9,
                   "roofline_frac": 8.0 * cr / (st3.kernel_ms * 1e-3) / 1e9 / peak,
                   "gpu_launches": int(st3.n_launches)}
        eng3.close()

    configs = []
    if not a.no_configs:
        for name in ("c1", "c3", "c4", "c5"):
            configs.append(run_config(name, make, Engine, local, rank, world, sync_all))

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ranks, t_setup = cpu_contexts(model, cores)
        n = int(a.cpu_photons) if a.cpu_photons else 0
        if n == 0:
            rate, _, _ = cpu_run(ranks, 10000 * cores)
            n = int(max(10000 * cores, min(rate * 15.0, 5e7)))
        rate, dt, cr = cpu_run(ranks, n)
        del ranks
        cpu = {"value": rate, "unit": "packets/s", "cores": cores, "kind": "port",
               "sample": "%d packets of the same model in %.1f s on %d threads (rank r seeded seed+r, grids summed)"
                         % (n, dt, cores),
               "compile_flags": oracle_flags(),
               "crossings_per_packet": cr / n}

    # N > 1: the N-rank reduced grid must equal a 1-rank replay of the same packet ids
    # (src/mpi/mpi_routines.f90:272-323 semantics: reduce, scale, broadcast)
    mg_check = None
    if world > 1:
        mg_check = multi_gpu_check(make, Engine, local, rank, world, sync_all)

    if rank == 0:
        value = P * world * a.steps / (total_ms * 1e-3)
        line = {
            "metric": "photon_packets_per_sec", "value": value, "unit": "packets/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "grid": [a.grid] * 3, "n_dust": n_dust,
                       "photons_per_gpu_per_step": P, "tau_centre_to_face": a.tau,
                       "dust": "9-point realistic table, isotropic, LTE emissivities 1200 states",
                       "cache": "cell records 16 B x n_cells = %.0f MB > 126 MB L2 (no flush needed)"
                                % (16e-6 * n_cells * n_dust)
                       if 16 * n_cells * n_dust > 126e6 else "working set fits L2"},
            "roofline": roof,
            "crossings_per_packet": cross / (P * a.steps),
            "absorptions_per_packet": nabs / (P * a.steps),
            "scatterings_per_packet": nscat / (P * a.steps),
            "rounds_per_step": sum(s.n_rounds for s in stats) / a.steps,
            "gpu_launches": gpu_launches,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if mg_check is not None:
            line["multi_gpu_check"] = mg_check
        others = [w for w in (thin, moderate, imaging) if w] + configs
        if others:
            line["other_workloads"] = others
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
