"""Drop-in replacement of the reference back-end program: ``hyperion_car [-f] input output``.

Mirrors ``program main`` (``src/main/main.f90:74-345``): open the ``.rtin``, copy or link
``/Input``, run the Lucy iterations with the convergence test and per-iteration grid output
(``output_grid``, ``src/grid/grid_generic.f90:29-130``), write the ``.rtout`` attributes the
Python front end expects (``hyperion/model/helpers.py:10``, ``hyperion/model/model_output.py``)
and ``date_ended`` only on success (the launcher treats its absence as failure,
``scripts/hyperion:94-104``).

The photon loop itself runs on the GPU through the C ABI (``hyperion_b200.capi``); there is no
CPU path.  Started as N ranks (``mpirun -n N hyperion_car_mpi``, ``srun``, ``torchrun``,
``bin/hyperion_mpirun`` or ``HYPERION_B200_NGPU=N``: ``hyperion_b200.launch``) every rank drives one
GPU, the packets are sharded by id and the deposit grid is all-reduced over NCCL once per iteration
(``hyperion_b200.multigpu``); rank 0 owns the files, as in the reference (``src/mpi/mpi_io.f90:213-242``).
"""
from __future__ import annotations

import datetime
import os
import sys
import time

import numpy as np

from . import __version__, launch
from .io import h5min, h5write
from .multigpu import ShardedLucy, shard
from .rtin import ModelError, read_rtin


class PeerFailure(ModelError):
    """Raised on the ranks that did not fail themselves when another rank reported an error."""


def wrap_error_text(text, width=61):
    """The line breaking of ``error()`` (``fortranlib/src/lib_messages.f90:141-171``): cut at the last blank
    within ``width`` characters, the blank stays at the end of the line."""
    lines, imin, n = [], 1, len(text)          # 1-based like the Fortran
    while True:
        if imin + width > n:
            imax = n
        else:
            imax = imin + width
            for j in range(width, 0, -1):
                if text[imin + j - 1] == " ":
                    imax = imin + j
                    break
        lines.append(text[imin - 1:imax])
        if imax >= n:
            return lines
        imin = imax + 1


def boxed_error(where, text, stream=None):
    """``error()`` of ``fortranlib/src/lib_messages.f90:126-179``, byte for byte (list-directed writes start
    with a blank): the reference's tests search the log for the wrapped message."""
    stream = stream or sys.stderr
    now = datetime.datetime.now().strftime("%d %B %Y at %H:%M:%S")
    out = [" " + "-" * 72]
    for k, line in enumerate(wrap_error_text(text)):
        out.append((" ERROR   : " if k == 0 else "           ") + line)
    out += [" WHERE   : " + where, " " + "-" * 72, "", "  *** Execution aborted on " + now + " ***", ""]
    stream.write("\n".join(out) + "\n")
    stream.flush()


class ConvergenceCheck:
    """``specific_energy_converged`` (``src/grid/grid_physics_3d.f90:630-689``)."""

    def __init__(self, absolute, relative, percentile, log=lambda s: None):
        self.absolute, self.relative, self.percentile = absolute, relative, percentile
        self.prev = None
        self.value_prev = None
        self.log = log

    @staticmethod
    def quantile(x, percent):
        """``quantile_dp`` (``fortranlib/src/lib_statistics.f90:102-125``)."""
        xs = np.sort(x)
        n = len(xs)
        if percent >= 100.0:
            ipos = n
        elif percent <= 0.0:
            ipos = 1
        else:
            ipos = int(np.floor(percent / 100.0 * (n - 1) + 0.5)) + 1   # Fortran nint for positive values
        return xs[ipos - 1]

    def __call__(self, se):
        se = np.asarray(se, dtype=np.float64).ravel()
        self.log(" [specific_energy_converged] checking convergence")
        if self.prev is None:
            self.prev = se.copy()
            return False
        prev = self.prev
        if np.array_equal(prev, se):
            value = 0.0
        elif np.all((prev == se) | (prev == 0) | (se == 0)):
            self.log(" [specific_energy_converged] could not check for convergence, as the only cells that "
                     "changed had zero value before or after")
            return False
        else:
            mask = (prev > 0) & (se > 0) & (prev != se)
            a, b = prev[mask], se[mask]
            value = self.quantile(np.maximum(a / b, b / a), self.percentile)
        self.log("     -> Percentile: %7.2f" % self.percentile)
        self.log("     -> Value @ Percentile: %10.3E" % value)
        converged = False
        if self.value_prev is not None:
            if value == 0.0:
                self.log("     -> Exact convergence")
                converged = True
            else:
                ratio = max(self.value_prev / value, value / self.value_prev) if self.value_prev > 0 else np.inf
                self.log("     -> Difference from previous iteration: %10.2f" % ratio)
                converged = bool(value < self.absolute and abs(ratio) < self.relative)
        self.prev = se.copy()
        self.value_prev = value
        return converged


def _copy_tree(src, dst):
    """Copy an h5min group into an h5write group (``mp_copy_group``, ``main.f90:145-147``)."""
    for k, v in src.attrs.items():
        dst.attrs[k] = v
    for name in src.keys():
        link = src.get_link(name)
        if isinstance(link, h5min.ExternalLink):
            dst[name] = h5write.ExternalLink(link.filename, link.path)
            continue
        if isinstance(link, h5min.SoftLink):
            dst[name] = h5write.SoftLink(link.path)
            continue
        child = src[name]
        if isinstance(child, h5min.Dataset):
            ds = dst.create_dataset(name, child.read())
            for k, v in child.attrs.items():
                ds.attrs[k] = v
        else:
            _copy_tree(child, dst.create_group(name))


def write_peeled_output(g, eng, ig, p, n_sources, n_dust, frequencies=None):
    """``image_write`` + ``peeled_images_write`` (``src/images/image_type.f90:608-788``,
    ``src/images/images_peeled.f90:384-408``): datasets ``seds`` / ``images`` (+ ``_unc``) with the
    attributes ``ModelOutput.get_sed`` / ``get_image`` read."""
    c_cgs = 2.99792458e10
    micron = float(np.float32(1.e-4))          # single-precision literal in image_type.f90:262-263
    n_wav, wav_min, wav_max = p.wavelengths
    nu_min, nu_max = c_cgs / (wav_max * micron), c_cgs / (wav_min * micron)
    exact_nu = p.inu_min > 0
    dt = np.float32 if p.io_bytes == 4 else np.float64

    def origin_attrs(d):
        d.attrs["track_origin"] = p.track_origin
        if p.track_origin == "detailed":
            d.attrs["n_sources"] = np.int32(n_sources)
            d.attrs["n_dust"] = np.int32(n_dust)
        elif p.track_origin == "scatterings":
            d.attrs["track_n_scat"] = np.int32(p.track_n_scat)

    if p.sed is not None:
        res = eng.sed(ig, p.uncertainties)
        val, unc = res if p.uncertainties else (res, None)
        d = g.create_dataset("seds", val.astype(dt))
        if unc is not None:
            g.create_dataset("seds_unc", unc.astype(dt))
        if not p.filters and not exact_nu:
            d.attrs["numin"], d.attrs["numax"] = float(nu_min), float(nu_max)
        d.attrs["apmin"], d.attrs["apmax"] = float(p.sed[1]), float(p.sed[2])
        origin_attrs(d)
    if p.image is not None:
        res = eng.image(ig, p.uncertainties)
        val, unc = res if p.uncertainties else (res, None)
        d = g.create_dataset("images", val.astype(dt))
        if unc is not None:
            g.create_dataset("images_unc", unc.astype(dt))
        if not p.filters and not exact_nu:
            d.attrs["numin"], d.attrs["numax"] = float(nu_min), float(nu_max)
        d.attrs["xmin"], d.attrs["xmax"] = float(p.image[2]), float(p.image[3])
        d.attrs["ymin"], d.attrs["ymax"] = float(p.image[4]), float(p.image[5])
        origin_attrs(d)
    if p.filters:
        # image_type.f90:775-779
        g.attrs["use_filters"] = "yes"
        g.attrs["n_filt"] = np.int32(len(p.filters))
        g.create_dataset("filt_nu0", np.array([f[2] for f in p.filters], dtype=np.float64))
    if exact_nu:
        # image_type.f90:781-784
        from .rtin_write import _table
        g.create_dataset("frequencies", _table([("nu", np.asarray(frequencies[p.inu_min - 1:p.inu_max], dtype=np.float64))]))
    if p.binned:
        return          # binned_images_write (images_binned.f90:85-89) writes the cubes only
    g.attrs["inside_observer"] = "yes" if p.inside_observer else "no"
    g.attrs["d_min"] = float(p.d_min)
    g.attrs["d_max"] = float(p.d_max)


def run(input_file, output_file, overwrite=False, device=None, log=None):
    """Run the model in ``input_file`` and write ``output_file``.  Returns 0 on success; raises
    ModelError / HyperionError for the conditions the reference reports through ``error()``."""
    from .capi import Engine

    rank, world, local, scheme = launch.layout()
    if device is not None:
        local = device
    main = rank == 0
    if log is None:
        def log(s):
            if main:
                print(s, flush=True)

    started = datetime.datetime.now().strftime("%d %B %Y at %H:%M:%S")
    log(" " + "-" * 60)
    log(" hyperion_b200 v%s (CUDA photon engine behind the Hyperion file interface)" % __version__)
    log(" Started on %s" % started)
    log(" Input:  %s" % input_file)
    log(" Output: %s" % output_file)
    log(" " + "-" * 60)
    t_start = time.time()

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        n_dev = torch.cuda.device_count()
        if n_dev < 1:
            raise ModelError("no CUDA device is visible to rank %d" % rank)
        local = local % n_dev
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            if scheme != "RANK" or "MASTER_ADDR" not in os.environ:
                # started by mpirun / srun: the rendezvous torchrun would have provided
                addr, port = launch.rendezvous(output_file)
                os.environ.setdefault("MASTER_ADDR", addr)
                os.environ.setdefault("MASTER_PORT", str(port))
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    class _Ranks:
        """Collectives of the run with the reference's fate sharing: ``error()`` on one rank stops the
        job (``mpi_abort``); here every collective is preceded by an agreement on an error flag, so a
        rank that failed never leaves its peers waiting in an all-reduce."""

        def __init__(self):
            self.stream = None

        def agree(self, failure):
            """All ranks call this with their exception or None; raises on every rank if any failed."""
            if world == 1:
                if failure is not None:
                    raise failure
                return
            flag = torch.tensor([1 if failure is not None else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag.item()):
                if failure is not None:
                    raise failure
                raise PeerFailure("another rank reported an error; stopping (see its message)")

        def guarded(self, fn, *a, **k):
            from .capi import HyperionError
            try:
                res = fn(*a, **k)
                failure = None
            except (HyperionError, ModelError, h5min.H5Error) as e:
                res, failure = None, e
            self.agree(failure)
            return res

        def all_reduce(self, buf):
            with torch.cuda.stream(self.stream):
                dist.all_reduce(buf)
            self.stream.synchronize()

        def sum_ints(self, values):
            """Sum a few host integers over the ranks (``mp_sync`` of the killed-photon counters,
            ``src/main/main.f90:317-323``)."""
            if world == 1:
                return [int(v) for v in values]
            t = torch.tensor([int(v) for v in values], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            return [int(v) for v in t.tolist()]

    ranks = _Ranks()

    def open_model():
        if not os.path.exists(input_file):
            raise ModelError("File does not exist: %s" % input_file)
        if main and os.path.exists(output_file):
            if not overwrite:
                raise ModelError("File exists: %s (use -f to overwrite)" % output_file)
            os.remove(output_file)
        model, rs, fin = read_rtin(input_file)
        if rs.monochromatic and model.binned is not None:
            raise ModelError("Binned images cannot be computed in monochromatic mode")   # hyperion/model/model.py:115
        if rs.pda and not getattr(model, "no_dust", False):
            # setup_rt.f90:296-302
            if any(getattr(d, "version", 2) == 1 for d in model.dust):
                raise ModelError("version 1 dust files can no longer be used when PDA is computed due to a bug - to fix "
                                 "this, re-generate the dust file using the latest version of Hyperion")
            if model.grid_type in ("oct", "amr", "vor"):
                raise ModelError("PDA is not available for this grid type")      # grid_pda_disabled.f90
            model.conf.use_pda = True
        # the n_photons array exists with the PDA or when it is to be written (grid_physics_3d.f90:308-317)
        model.conf.count_photons = rs.output_n_photons != "none"
        if rs.specific_energy_type == "additional":
            # setup_initial (src/main/setup_rt.f90:191-194)
            if rs.n_initial_iter == 0:
                raise ModelError("Cannot use specific_energy_type='additional' if the number of specific energy iterations is 0")
            model.conf.specific_energy_additional = True
        return model, rs, fin

    model, rs, fin = ranks.guarded(open_model)
    log(" [main] using random seed = %d" % model.conf.seed)
    eng = ranks.guarded(lambda: Engine(local))
    ranks.guarded(eng.load_model, model)
    all_reduce = None
    if world > 1:
        stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))
        ranks.stream = stream
        all_reduce = ranks.all_reduce

    class _GuardedEngine:
        """The engine as ShardedLucy sees it: every phase that can fail agrees on the outcome before the
        collective that follows it."""

        def lucy_begin(self):
            ranks.guarded(eng.lucy_begin)

        def lucy_photons(self, first, count, iteration):
            ranks.guarded(eng.lucy_photons, first, count, iteration)

        def reduction_buffer(self):
            return eng.reduction_buffer()

        def lucy_finish(self):
            return ranks.guarded(eng.lucy_finish)

    drv = ShardedLucy(_GuardedEngine() if world > 1 else eng, rank, world, all_reduce)

    out = h5write.File()
    out.attrs["date_started"] = started
    out.attrs["fortran_version"] = "hyperion_b200 " + __version__
    if rs.copy_input:
        _copy_tree(fin, out.create_group("Input"))
    else:
        out["Input"] = h5write.ExternalLink(input_file, "/")

    check = ConvergenceCheck(rs.convergence_absolute, rs.convergence_relative, rs.convergence_percentile, log) \
        if rs.check_convergence else None
    io_dtype = np.float32 if rs.physics_io_bytes == 4 else np.float64
    density0 = model.density.copy() if rs.output_density_diff != "none" else None
    converged = False
    n_done = rs.n_initial_iter
    for it in range(1, rs.n_initial_iter + 1):
        log(" [main] starting Lucy iteration %d" % it)
        st = drv.iteration(rs.n_initial_photons, it)
        log(" [main] exiting Lucy iteration")
        se = None
        if check is not None:
            se = eng.get_specific_energy()
            converged = check(se)
            if converged:
                log("      ------ Specific energy calculation converged -----")
        g = out.create_group("iteration_%05d" % it)
        n_iter = it if (check is not None and converged) else rs.n_initial_iter

        def wanted(mode):
            return mode == "all" or (mode == "last" and it == n_iter)

        log(" [output_grid] outputting grid arrays for iteration")
        def put(name, arr, dtype=None):
            """output_grid (src/grid/grid_generic.f90:29-130): one dataset per iteration, or, for AMR
            grids, one per level / grid (src/grid/grid_io_amr.f90)."""
            if model.grid_type == "amr":
                for il, ig, sl, shp in model.amr_slices():
                    path = "level_%05d/grid_%05d" % (il + 1, ig + 1)
                    gg = g.require_group(path)
                    a = arr[:, sl].reshape((-1,) + shp)
                    gg.create_dataset(name, (a[0] if dtype is not None else a).astype(dtype or io_dtype))
            else:
                d = g.create_dataset(name, (arr[0] if dtype is not None else arr).astype(dtype or io_dtype))
                d.attrs["geometry"] = rs.geometry_id

        if wanted(rs.output_n_photons):
            # output_grid (grid_generic.f90:40-46): one value per cell, no dust dimension
            put("n_photons", eng.get_n_photons()[None], dtype=np.int64)
        if wanted(rs.output_specific_energy):
            if se is None:
                se = eng.get_specific_energy()
            put("specific_energy", se)
        if wanted(rs.output_specific_energy_spectrum):
            # output_grid (grid_generic.f90:68-88): the bin edges (1-D) and the spectrum [n_bins, n_dust, cells]
            g.create_dataset("specific_energy_spectrum_bin_edges", np.asarray(model.spectrum_bin_edges, dtype=np.float64))
            se_nu = eng.get_specific_energy_spectrum()
            if model.grid_type == "amr":
                for il, ig, sl, shp in model.amr_slices():
                    gg = g.require_group("level_%05d/grid_%05d" % (il + 1, ig + 1))
                    gg.create_dataset("specific_energy_spectrum",
                                      se_nu[:, :, sl].reshape(se_nu.shape[:2] + shp).astype(io_dtype))
            else:
                d = g.create_dataset("specific_energy_spectrum", se_nu.astype(io_dtype))
                d.attrs["geometry"] = rs.geometry_id
        if wanted(rs.output_density):
            put("density", eng.get_density())
        if wanted(rs.output_density_diff):
            put("density_diff", eng.get_density() - density0)
        g.attrs["killed_photons_geo"] = np.int64(st.killed_geo)
        g.attrs["killed_photons_int"] = np.int64(st.killed_int)
        if check is not None and converged:
            n_done = it
            break

    out.attrs["converged"] = "yes" if converged else "no"
    out.attrs["iterations"] = np.int32(n_done)

    # FINAL ITERATION (main.f90:253-290): imaging packets with peel-off; with raytracing on, only
    # scattered light is peeled here (iter_final.f90:119-121)
    log(" [main] starting final iteration")
    killed_final = (0, 0)
    make_peeled = len(model.peeled) > 0
    make_binned = model.binned is not None
    if make_binned:
        log(" [binned_images] setting up %d binned images " % (model.binned.n_theta * model.binned.n_phi))
    if make_peeled:
        log(" [peeled_images] setting up %d peeled image groups " % len(model.peeled))
    mono_counts = (0, 0)
    if rs.monochromatic:
        # do_final_mono (iter_final_mono.f90:58-229): n_last_photons_sources source packets and n_last_photons_dust
        # thermal packets PER FREQUENCY, every packet already scaled, so the cubes of the ranks simply add up
        mono_counts = (rs.n_last_photons_sources if model.sources else 0, rs.n_last_photons_dust)
    if sum(mono_counts) > 0:
        fs, cs = shard(mono_counts[0], rank, world)
        fd, cd = shard(mono_counts[1], rank, world)
        ranks.guarded(eng.final_begin)
        for inu, nu in enumerate(model.frequencies, start=1):
            if mono_counts[0] > 0:
                log(" [mono] computing source photons for nu =%11.4E Hz" % nu)
            if mono_counts[1] > 0:
                log(" [mono] computing dust photons for nu =%11.4E Hz" % nu)
            ranks.guarded(eng.final_mono_photons, inu, fs, cs, mono_counts[0], fd, cd, mono_counts[1], rs.raytracing)
        st = ranks.guarded(eng.final_finish)
        killed_final = tuple(ranks.sum_ints((st.killed_geo, st.killed_int)))
    if rs.n_last_photons > 0:
        first, count = shard(rs.n_last_photons, rank, world)
        ranks.guarded(eng.final_begin)
        ranks.guarded(eng.final_photons, first, count, rs.raytracing)
    log(" [main] exiting final iteration")
    n_ray = (0, 0)
    if rs.raytracing:
        n_ray = (rs.n_ray_photons_sources if model.sources else 0, rs.n_ray_photons_dust)
    if rs.n_last_photons > 0:
        # scale by energy_total / energy_current over ALL ranks (iter_final.f90:136-143): the emitted
        # energy travels with the cubes, so reduce first, then scale
        if world > 1:
            all_reduce(eng.image_buffer())
        st = ranks.guarded(eng.final_finish)
        killed_final = (st.killed_geo, st.killed_int)
        if world > 1:
            # every rank applied the scale to the reduced cubes; keep one copy for the final sum below
            if rank != 0:
                with torch.cuda.stream(stream):
                    eng.image_buffer().zero_()
                stream.synchronize()
    out.attrs["killed_photons_geo_final"] = np.int64(killed_final[0])
    out.attrs["killed_photons_int_final"] = np.int64(killed_final[1])
    killed_ray = (0, 0)
    if rs.raytracing:
        if any(p.filters for p in model.peeled):
            raise ModelError("filter convolution cannot be used with raytracing")     # image_type.f90:541
        log(" [main] starting raytracing iteration")
        fs, cs = shard(n_ray[0], rank, world)
        fd, cd = shard(n_ray[1], rank, world)
        st = ranks.guarded(eng.raytracing_photons, cs, cd, first_source_id=fs, n_total_sources=n_ray[0],
                           first_dust_id=fd, n_total_dust=n_ray[1])
        # every rank counted its own share (main.f90:317-323 syncs the counters)
        killed_ray = tuple(ranks.sum_ints((st.killed_geo, st.killed_int)))
        log(" [main] exiting raytracing iteration")
    out.attrs["killed_photons_geo_raytracing"] = np.int64(killed_ray[0])
    out.attrs["killed_photons_int_raytracing"] = np.int64(killed_ray[1])
    # mp_collect_images (mpi_routines.f90:363-471)
    if (make_peeled or make_binned) and world > 1:
        all_reduce(eng.image_buffer())
    if make_binned:
        # main.f90:263,326: the cubes go straight into /Binned
        write_peeled_output(out.create_group("Binned"), eng, len(model.peeled), model.binned,
                            len(model.sources), 0 if getattr(model, "no_dust", False) else len(model.dust))
    if make_peeled:
        gp = out.create_group("Peeled")
        for ig, p in enumerate(model.peeled):
            write_peeled_output(gp.create_group("group_%05d" % (ig + 1)), eng, ig, p,
                                len(model.sources), 0 if getattr(model, "no_dust", False) else len(model.dust),
                                model.frequencies)

    eng.close()
    out.attrs["cpu_time"] = float(time.time() - t_start)
    ended = datetime.datetime.now().strftime("%d %B %Y at %H:%M:%S")
    out.attrs["date_ended"] = ended
    if main:
        out.write(output_file)
    log(" " + "-" * 60)
    log(" Total time elapsed: %16.2f" % (time.time() - t_start))
    log(" Ended on %s" % ended)
    log(" " + "-" * 60)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main(argv=None):
    """``hyperion_car [-f] input_file output_file`` (``src/main/main.f90:74-106``)."""
    argv = list(sys.argv[1:] if argv is None else argv)
    overwrite = False
    if argv and argv[0] == "-f":
        overwrite = True
        argv = argv[1:]
    if len(argv) != 2:
        sys.stderr.write("Usage: hyperion_car|hyperion_sph [-f] input_file output_file\n")
        return 2
    # a serial start with HYPERION_B200_NGPU=N fans out to N ranks, one per GPU (the reference's -m N)
    if launch.layout()[3] is None and launch.wanted_gpus() > 1:
        return launch.spawn(launch.wanted_gpus(), [sys.executable, "-m", "hyperion_b200"] + list(sys.argv[1:] if argv is None else
                            (["-f"] if overwrite else []) + argv))
    from .capi import HyperionError
    try:
        return run(argv[0], argv[1], overwrite=overwrite)
    except PeerFailure:
        return 1          # the failing rank has printed the message
    except (ModelError, HyperionError, h5min.H5Error) as e:
        boxed_error("main", str(e))
        return 1


if __name__ == "__main__":
    sys.exit(main())
