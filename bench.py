#!/usr/bin/env python
"""bench.py -- particle-substeps/sec of the PIC/FLIP transfer hot path (P2G + PIC/FLIP G2P + RK4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gfs|reference] [--workload NAME]

One "step" = one substep of the hot path over the whole particle set of the workload: cell sort,
classification + u/v/w splat + normalisation + face assembly (P2G), then PIC/FLIP velocity update + RK4
advection + solid test (G2P).  Metric and roofline definitions: SURVEY.md §8(d), DESIGN.md §5.

  value     device-resident throughput (inputs already in HBM), CUDA events on the launch stream
  e2e       the same substep through the host-pointer C-ABI calls, with the host->device copy of particles and
            of the new/saved fields and the device->host read of particles, P2G fields and material inside the
            timed region
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the unmodified reference's CPU stages (oracle/_ref/libgfsref.so; the oracle port if that library is
            absent) on a bounded sample of the same workload, all host threads

Multi-GPU (N > 1, launched by torch.distributed.run): z-slab sharding, see gridfluidsim3d_b200/slabs.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec (P2G+RK4 G2P)"
UNIT = "particle-substeps/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback ("of fallback")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
# synthetic workload on the device (same construction as gridfluidsim3d_b200/synth.py, generated with torch so
# that the 100 M-particle scene takes seconds, not minutes)
# ---------------------------------------------------------------------------------------------------------
def make_scene_device(name, device, seed=12345, k_range=None):
    """Counter-based scene (gridfluidsim3d_b200/synth.py:make_scene_torch): the particle SET depends on (name, seed)
    only -- every rank of a sharded run generates its own layers of the very scene the single-GPU run uses."""
    from gridfluidsim3d_b200 import synth
    return synth.make_scene_torch(name, device, seed=seed, k_range=k_range)


# ---------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (one sample per ~1 ms; the timed
    region of a short run lasts tens of ms), with `nvidia-smi -lms` as the fallback when NVML is not importable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    # NVML clocks-event-reason bits (nvml.h)
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.sm, self.mx, self.reasons, self.stop = [], [], set(), False
        self.nvml = self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)      # honours CUDA_VISIBLE_DEVICES
                try:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for nm, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.nvml is not None:
            self.stop = True
            self.thread.join(timeout=2)
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------
# CPU reference timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------
def host_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    try:
        with open("/proc/meminfo") as f:
            avail_kb = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0]
    except (OSError, IndexError):
        avail_kb = 64 << 20
    return n, avail_kb


class CpuHotpath:
    """The reference's CPU stages on a bounded particle sample against the full-size grid.  Grid-proportional
    cost (measured once with zero particles) is charged pro rata n_sample/N, so the figure estimates the
    throughput of the full workload spread over `cores` threads."""

    def __init__(self, scene_host, interp, per_thread, force_cores=None):
        from oracle import pyoracle
        self.s, self.interp = scene_host, interp
        dims = scene_host["dims"]
        cells = dims[0] * dims[1] * dims[2]
        ncpu, avail_kb = host_threads()
        per_sim_kb = max(1, cells * 90 // 1024)          # ~90 B/cell per private simulator (measured at 128^3)
        self.cores = int(max(1, min(ncpu, 32, avail_kb // 2 // per_sim_kb)))
        if force_cores:
            self.cores = int(force_cores)
        self.kind, self.ref = "port", None
        try:
            self.ref = pyoracle.Reference(build=os.path.isdir("/root/reference/src"))
            self.kind = "reference"
        except (FileNotFoundError, OSError) as e:
            log("reference library unavailable (%s): timing the oracle port instead" % e)
            self.orc = pyoracle.Oracle()
            self.cores = int(force_cores) if force_cores else int(min(ncpu, self.orc.lib.orc_max_threads()))
        N = len(scene_host["pos"])
        self.n = int(min(N, per_thread * self.cores))
        self.N = N
        if self.kind == "reference":
            self.hp = self.ref.hotpath(self.cores, dims, scene_host["dx"], scene_host["new"], scene_host["saved"])
            _, st = self.hp.step(scene_host["pos"][:0], scene_host["vel"][:0], scene_host["dt"], interp)
            self.t_grid = float(st.sum())
        else:
            self.t_grid = 0.0

    def sample_desc(self):
        return ("%d of %d particles (first %d of the shuffled array, %d per thread) against the full %dx%dx%d grid; "
                "grid-proportional cost %.2fs (zero-particle run) charged pro rata" %
                (self.n, self.N, self.n, self.n // self.cores, *self.s["dims"], self.t_grid))

    def step(self):
        """-> (particle-substeps/s estimate for the full workload, seconds spent)"""
        s = self.s
        pos, vel = s["pos"][:self.n], s["vel"][:self.n]
        t0 = time.time()
        if self.kind == "reference":
            t, st = self.hp.step(pos, vel, s["dt"], self.interp)
            t_eff = max(t - self.t_grid, 1e-9) + self.t_grid * self.n / self.N
        else:
            mat = s["material"].copy()
            self.orc.p2g(pos, vel, s["dims"], s["dx"], mat)
            self.orc.g2p_advect(pos, vel, s["new"], s["saved"], s["dims"], s["dx"], s["dt"], mode=self.interp, material=mat)
            t_eff = time.time() - t0
        return self.n / t_eff, time.time() - t0

    def close(self):
        if self.kind == "reference":
            self.hp.close()


def scene_to_host(sc, max_particles=None):
    aos = sc["aos"] if max_particles is None else sc["aos"][:max_particles]
    a = aos.cpu().numpy()
    return dict(dims=sc["dims"], dx=sc["dx"], dt=sc["dt"], material=sc["material"],
                pos=np.ascontiguousarray(a[:, :3]), vel=np.ascontiguousarray(a[:, 3:]),
                new=[t.cpu().numpy() for t in sc["new"]], saved=[t.cpu().numpy() for t in sc["saved"]])


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this arm's config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from gridfluidsim3d_b200 import synth
    interp = 0 if args.interp == "trilinear" else 1
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    per_thread = args.cpu_sample
    dims, dx, _ = synth.CONFIGS[args.workload]
    sc = make_scene_device(args.workload, dev)
    N = sc["aos"].shape[0]
    ncpu, _ = host_threads()
    host = scene_to_host(sc, max_particles=per_thread * min(ncpu, 32))
    host_full_n = N
    del sc
    cpu = CpuHotpath(host, interp, per_thread)
    cpu.N = host_full_n
    for _ in range(args.warmup):
        cpu.step()
    vals, t0 = [], time.time()
    for _ in range(args.steps):
        v, _ = cpu.step()
        vals.append(v)
    wall = time.time() - t0
    value = float(len(vals) / sum(1.0 / v for v in vals))        # harmonic mean = total particles / total time
    G = dims[0] * dims[1] * dims[2]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * host_full_n / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 (fp64 index/interpolation arithmetic)", "data": "synthetic",
        "config": {"workload": args.workload, "grid": list(dims), "dx": dx, "particles": host_full_n, "cells": G,
                   "interp": args.interp, "rk_order": 4, "note": "CPU, OpenCL disabled; ms_per_step is the full-workload estimate"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.sample_desc()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    cpu.close()
    emit(line)


# ---------------------------------------------------------------------------------------------------------
ALG_BYTES = {   # algorithmic bytes per launch of each hot kernel: (per particle, per cell)
    "gfs::k_p2g_scatter<0>": (24, 13), "gfs::k_p2g_tile<0>": (24, 13), "gfs::k_p2g_scatter<2>": (24, 13), "gfs::k_p2g_tile<2>": (24, 13),
    "gfs::k_p2g_tile2<0>": (24, 13), "gfs::k_p2g_tile2<1>": (24, 13), "gfs::k_g2p_tri<0>": (48, 24), "gfs::k_g2p_tri<1>": (48, 24),
    "gfs::k_g2p_advect<0>": (48, 24), "gfs::k_g2p_advect<1>": (48, 24), "gfs::k_g2p_advect<2>": (48, 24), "gfs::k_g2p_brick<0>": (48, 24), "gfs::k_g2p_brick<1>": (48, 24),
}


def run_gfs(args):
    import torch
    import torch.distributed as dist
    from gridfluidsim3d_b200 import capi, slabs, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        log("warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE" % (args.gpus, world))
    interp = capi.TRILINEAR if args.interp == "trilinear" else capi.TRICUBIC
    dims, dx, _ = synth.CONFIGS[args.workload]
    G = dims[0] * dims[1] * dims[2]
    hbm_gbs, peak_src = measured_peaks()
    if world > 1 and args.cuts == "weighted":
        # cuts that balance particles per slab (8 per seeded cell); every rank computes the same cuts from the scene
        layer_counts = synth.fluid_cells(synth.CONFIGS[args.workload][2], dims, synth.border_material(dims)).reshape(dims[2], -1).sum(1)
        ranges = slabs.slab_ranges_weighted(layer_counts, world, min_layers=max(4, capi.slab_halo_cells(capi.TRICUBIC, 0.5 * dx, dx)))
    else:
        ranges = slabs.slab_ranges(dims[2], world)
    owned = ranges[rank]

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    halo = capi.slab_halo_cells(interp, 0.5 * dx, dx)
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream: events and kernels share it
    torch.cuda.set_stream(stream)

    def build(owned, final):
        """Scene (this rank's layers), context, resident upload and -- for N > 1 -- the slab driver and its transport."""
        t_gen = time.time()
        sc = make_scene_device(args.workload, dev, seed=12345, k_range=owned if world > 1 else None)
        torch.cuda.synchronize()
        n_local = sc["aos"].shape[0]
        N = allsum(n_local)
        if rank == 0:
            log("scene %s: %d particles (%d on rank 0), %d cells, generated in %.1fs" % (args.workload, N, n_local, G, time.time() - t_gen))
        ctx = capi.Context(local, stream=stream.cuda_stream)
        if rank == 0 and final:
            log(ctx.device_info())
        ctx.domain_init(dims, dx)
        ctx.set_material(sc["material"])
        # pinned host copies: the e2e leg's inputs, and the source of the resident upload
        aos_host = torch.empty(sc["aos"].shape, dtype=torch.float32, pin_memory=True)
        aos_host.copy_(sc["aos"])
        new_host = [torch.empty(t.shape, dtype=torch.float32, pin_memory=True).copy_(t) for t in sc["new"]]
        saved_host = [torch.empty(t.shape, dtype=torch.float32, pin_memory=True).copy_(t) for t in sc["saved"]]
        torch.cuda.synchronize()
        ctx.set_particles_aos(aos_host.numpy())
        ctx.set_field(capi.FIELD_NEW, *[t.numpy() for t in new_host])
        ctx.set_field(capi.FIELD_SAVED, *[t.numpy() for t in saved_host])
        host_small = scene_to_host(sc, max_particles=args.cpu_sample * 32) if (final and world == 1 and not args.no_cpu_baseline) else None
        dt = sc["dt"]
        del sc
        torch.cuda.empty_cache()
        drv = transport = make_driver = None
        if world > 1:
            def make_driver(ip):
                return slabs.SlabDriver(slabs.CudaSlabBackend(ctx, dims, owned, ip, migrate_cap=max(4096, n_local // 8), shared_stream=True), rank, world,
                                        halo=capi.slab_halo_cells(ip, 0.5 * dx, dx))
            drv = make_driver(interp)
            if args.transport == "peer":      # neighbours write into each other's HBM over NVLink (gfs_comm_*), no NCCL in the data path
                other_ip = capi.TRICUBIC if interp == capi.TRILINEAR else capi.TRILINEAR
                transport = slabs.PeerTransport(drv, particle_cap=max(4096, n_local // 8),
                                                layer_bytes=slabs.PeerTransport.layer_bytes(make_driver(other_ip)))
            else:
                transport = slabs.DistTransport()
        return dict(ctx=ctx, n_local=n_local, N=N, aos_host=aos_host, new_host=new_host, saved_host=saved_host, dt=dt,
                    host_small=host_small, drv=drv, transport=transport, make_driver=make_driver)

    balance = None
    for cal_pass in range(args.balance_passes if (world > 1 and args.cuts == "weighted" and not args.no_balance) else 0):
        # One calibration pass: particle-weighted cuts equalise the particle COUNT, but the cost per particle is not the same
        # in every slab (partially filled bricks at the free surface, one exchange partner instead of two at the ends), and a
        # rank that finishes early only spins on its neighbours' flags.  Measure every rank's busy time (its kernels minus the
        # flag waits) over a few substeps, turn it into a cost per particle of its layers, and cut again.
        B = build(owned, final=False)
        for _ in range(3):
            slabs.substep(B["drv"], B["transport"], B["dt"])
        B["ctx"].profile_enable(True); B["ctx"].profile_read(reset=True)
        for _ in range(4):
            slabs.substep(B["drv"], B["transport"], B["dt"])
        prof = B["ctx"].profile_read(reset=True)
        B["ctx"].profile_enable(False)
        waits = ("k_gather_counts", "k_copy_batch_wait", "k_allmax")
        busy = sum(ms for k, (ms, cnt) in prof.items() if not any(w in k for w in waits)) / 4.0
        wait = sum(ms for k, (ms, cnt) in prof.items() if any(w in k for w in waits)) / 4.0
        rows = [None] * world
        dist.all_gather_object(rows, (busy, wait, B["n_local"]))
        cost = [b_ / max(1, n_) for (b_, w_, n_) in rows]
        weights = np.asarray(layer_counts, np.float64).copy()      # particles per layer x measured cost per particle of the slab it is in
        for r, (k0, k1) in enumerate(ranges):
            weights[k0:k1] *= cost[r] / (sum(cost) / world)
        old_ranges = ranges
        ranges = slabs.slab_ranges_weighted(weights, world, min_layers=max(4, capi.slab_halo_cells(capi.TRICUBIC, 0.5 * dx, dx)))
        owned = ranges[rank]
        balance = {"pass": cal_pass, "calibration_cuts": [list(r) for r in old_ranges], "busy_ms": [round(r_[0], 4) for r_ in rows],
                   "wait_ms": [round(r_[1], 4) for r_ in rows], "particles": [r_[2] for r_ in rows],
                   "earlier_passes": ([balance] if balance else [])}
        dist.barrier()
        B["ctx"].close()
        del B
        torch.cuda.empty_cache()
        dist.barrier()

    B = build(owned, final=True)
    ctx, n_local, N, aos_host, new_host, saved_host, dt = B["ctx"], B["n_local"], B["N"], B["aos_host"], B["new_host"], B["saved_host"], B["dt"]
    host_small, drv, transport, make_driver = B["host_small"], B["drv"], B["transport"], B["make_driver"]
    if world > 1:
        def substep():
            slabs.substep(drv, transport, dt)
    else:
        def substep():
            ctx.substep(dt, order=4, interp=interp, arith=capi.FAST)

    # ---- value: device-resident substeps --------------------------------------------------------------
    for _ in range(args.warmup):
        substep()
    barrier()
    st0 = ctx.stats()
    launches0 = st0["kernel_launches"]
    bytes0 = transport.bytes_sent if transport else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            substep()
        e1.record(stream)
        barrier()
    ms = allmax(e0.elapsed_time(e1))
    st1 = ctx.stats()
    launches_timed = st1["kernel_launches"] - launches0
    graph_replays = st1["graph_replays"] - st0["graph_replays"]
    comm_bytes = (transport.bytes_sent - bytes0) / args.steps if transport else 0
    # per-kernel CUDA-event times: a second pass of the same K steps with the library's profiling on (event pairs around
    # every launch; the single-GPU substep is launched kernel by kernel instead of as a replayed graph while it is on)
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        substep()
    p1.record(stream)
    barrier()
    prof_ms = p0.elapsed_time(p1)
    prof = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    # ---- verification (outside every timed region): order- and distribution-independent hashes of the state after
    # warmup + 2 x steps substeps.  The sharded result is designed to be bit-identical to the single-GPU one, so these
    # five numbers must be the same at every GPU count for the same --steps / --warmup.
    local_hash = ctx.state_hash()
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local_hash)
    else:
        gathered = [local_hash]
    hashes = [sum(h[i] for h in gathered) & ((1 << 64) - 1) for i in range(5)]
    verify = {"after_substeps": args.warmup + 2 * args.steps, "seed": 12345,
              "hash_material": "%016x" % hashes[0], "hash_p2g_u": "%016x" % hashes[1], "hash_p2g_v": "%016x" % hashes[2],
              "hash_p2g_w": "%016x" % hashes[3], "hash_particles": "%016x" % hashes[4],
              "note": "gfs_state_hash summed over ranks mod 2^64; identical for every --gpus N by construction of the path"}
    st = ctx.stats()
    launches = launches_timed
    n_now = allsum(ctx.num_particles)
    n_max = allmax(ctx.num_particles)
    ms_per_step = ms / args.steps
    value = N / (ms_per_step * 1e-3)

    # dominant kernel (on this rank) and its roofline: algorithmic bytes of the launch / CUDA-event time of the launch
    ranked = sorted(prof.items(), key=lambda kv: -kv[1][0])
    kernels = {k: {"ms_per_launch": v[0] / max(1, v[1]), "launches_per_step": v[1] / args.steps,
                   "share_of_step": v[0] / prof_ms} for k, v in ranked}
    top = next((k for k, _ in ranked if k in ALG_BYTES), ranked[0][0])
    pb, cb = ALG_BYTES.get(top, (0, 0))
    top_ms = prof[top][0] / max(1, prof[top][1])
    G_local = G * (owned[1] - owned[0]) // dims[2]
    alg_bytes = pb * ctx.num_particles + cb * G_local
    achieved = alg_bytes / (top_ms * 1e-3) / 1e9
    step_alg_bytes = 72 * N + 37 * G
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                "frac": achieved / hbm_gbs, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": top_ms,
                "substep_algorithmic_bytes": step_alg_bytes,
                "substep_frac": step_alg_bytes / (ms_per_step * 1e-3) / 1e9 / (hbm_gbs * world)}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the ncu capture
    if os.path.exists(traffic_file) and world == 1:
        with open(traffic_file) as f:
            import re
            m = re.search(r"(\w+)\s*<\s*(\d+)", top)          # "gfs::k_g2p_brick<0>" -> "k_g2p_brick<0>", as profiles/summarize.py keys it
            roofline["traffic"] = json.load(f).get(args.workload, {}).get("%s<%s>" % (m.group(1), m.group(2)) if m else top.split("::")[-1])
    # The HBM fraction above is the contract's number; what the kernel is actually bound by comes from its ncu capture
    # (profiles/bounds.json, written by profiles/summarize.py): both hot kernels of this path sit on the shared-memory pipe.
    bounds_file = os.path.join(ROOT, "profiles", "bounds.json")
    if os.path.exists(bounds_file):
        import re
        with open(bounds_file) as f:
            allb = json.load(f).get(args.workload, {})

        def bkey(name):
            m = re.search(r"(\w+)\s*<\s*(\d+)", name)
            return "%s<%s>" % (m.group(1), m.group(2)) if m else name.split("::")[-1]
        roofline["ncu_pipe_utilisation"] = {k: allb[bkey(k)] for k, _ in ranked[:3] if bkey(k) in allb} or None
        if bkey(top) in allb and allb[bkey(top)].get("lsu_wavefronts_pct_of_peak"):
            roofline["limiting_pipe"] = {"pipe": "shared-memory LSU wavefronts (1 per clock per SM)",
                                         "frac": allb[bkey(top)]["lsu_wavefronts_pct_of_peak"] / 100.0,
                                         "source": allb[bkey(top)]["source"]}

    # ---- e2e: the same substep through host buffers ------------------------------------------------------
    aos_out = torch.empty((int(n_max) + 1024, 6), dtype=torch.float32, pin_memory=True)
    p2g_pinned = [torch.empty(t.numel(), dtype=torch.float32, pin_memory=True) for t in new_host]     # every host buffer of the
    p2g_out = [t.numpy() for t in p2g_pinned]                                                         # e2e leg is pinned
    mat_pinned = torch.empty(G, dtype=torch.uint8, pin_memory=True)
    nu, nv, nw = [t.numel() for t in new_host]
    if world == 1:
        up_lo, up_n, dn_lo, dn_n = 0, dims[2], 0, dims[2]
    else:           # a z-slab rank moves its owned layers + the G2P halo up, and what it owns down
        up_lo, up_hi = max(0, owned[0] - halo), min(dims[2], owned[1] + halo)
        up_n, dn_lo, dn_n = up_hi - up_lo, owned[0], owned[1] - owned[0]
    per_layer = (dims[0] + 1) * dims[1] + dims[0] * (dims[1] + 1) + dims[0] * dims[1]
    h2d = n_local * 24 + 2 * 4 * (per_layer * up_n + dims[0] * dims[1])
    d2h = n_local * 24 + 4 * (per_layer * dn_n + dims[0] * dims[1]) + dims[0] * dims[1] * dn_n
    new_np, saved_np = [t.numpy() for t in new_host], [t.numpy() for t in saved_host]

    e2e_parts = {}

    def timed(name, fn, *a):
        t = time.perf_counter()
        fn(*a)                                  # every one of these C-ABI calls returns after its own stream synchronisation
        e2e_parts[name] = e2e_parts.get(name, 0.0) + time.perf_counter() - t

    def e2e_step():
        timed("set_particles", ctx.set_particles_aos, aos_host.numpy())                                  # H2D 24 B/particle
        timed("set_fields", ctx.set_field_layers, capi.FIELD_NEW, *new_np, up_lo, up_n)                  # H2D post-pressure field (owned layers + halo)
        timed("set_fields", ctx.set_field_layers, capi.FIELD_SAVED, *saved_np, up_lo, up_n)              # H2D saved field
        timed("substep", substep)
        timed("get_particles", ctx.get_particles_aos, aos_out.numpy().reshape(-1)[: ctx.num_particles * 6])   # D2H particles
        timed("get_fields", ctx.get_field_layers, capi.FIELD_P2G, p2g_out, dn_lo, dn_n)                  # D2H P2G u,v,w (owned layers)
        timed("get_material", ctx.get_material_layers, mat_pinned.numpy(), dn_lo, dn_n)                  # D2H material

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step()
    barrier()
    e2e_parts.clear()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = allmax((time.perf_counter() - t0) / e2e_steps)
    e2e = {"value": N / e2e_s, "unit": UNIT, "h2d_bytes_per_step": allsum(h2d), "d2h_bytes_per_step": allsum(d2h),
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "breakdown_ms_rank0": {k: round(v / e2e_steps * 1e3, 3) for k, v in e2e_parts.items()},
           "api": "gfs_set_particles + gfs_set_field_layers x2 + substep + gfs_get_particles + gfs_get_field_layers + gfs_get_material_layers"
                  + (" (per rank: owned layers + halo up, owned layers down)" if world > 1 else " (all layers)")}

    # ---- the other interpolation, short --------------------------------------------------------------------
    other = capi.TRICUBIC if interp == capi.TRILINEAR else capi.TRILINEAR
    other_name = "tricubic" if interp == capi.TRILINEAR else "trilinear"
    ctx.set_particles_aos(aos_host.numpy())
    if world > 1:
        drv2 = make_driver(other)

        def substep2():
            slabs.substep(drv2, transport, dt)
    else:
        def substep2():
            ctx.substep(dt, order=4, interp=other, arith=capi.FAST)
    for _ in range(2):
        substep2()
    barrier()
    vs = max(3, min(args.steps, 5))
    e0.record(stream)
    for _ in range(vs):
        substep2()
    e1.record(stream)
    barrier()
    oms = allmax(e0.elapsed_time(e1)) / vs
    variants = {other_name: {"value": N / (oms * 1e-3), "ms_per_step": oms,
                             "substep_frac": step_alg_bytes / (oms * 1e-3) / 1e9 / (hbm_gbs * world)}}
    # The tricubic G2P kernel is bound by shared-memory bandwidth, not HBM: 5 samples x 3 components x 64 taps
    # x 4 B = 3840 B of shared-memory reads per particle against 128 B/clk/SM.  Two profiled steps give its kernel time.
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    for _ in range(2):
        (substep2 if other == capi.TRICUBIC else substep)()
    barrier()
    pc = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    cubic = [v for k, v in pc.items() if "k_g2p_brick<1>" in k]
    if cubic and cubic[0][1] > 0:
        k_ms = cubic[0][0] / cubic[0][1]
        sm_count, clk_ghz = 148, (clocks.summary()["sm_mhz"] or 1965.0) / 1e3
        smem_tbs = sm_count * 128 * clk_ghz / 1e3
        smem_bytes = 3840.0 * ctx.num_particles
        tc = variants["tricubic"] if "tricubic" in variants else variants.setdefault("tricubic_kernel", {})
        tc["g2p_kernel_ms"] = k_ms
        tc["smem_bound"] = {"bytes_per_particle": 3840, "peak_tb_s": smem_tbs, "achieved_tb_s": smem_bytes / (k_ms * 1e-3) / 1e12,
                            "frac": smem_bytes / (k_ms * 1e-3) / 1e12 / smem_tbs,
                            "note": "k_g2p_brick<1> on rank 0; peak = 148 SMs x 128 B/clk x SM clock; ncu: 87 % of LSU wavefront peak (profiles/r01_k_g2p_brick_tricubic.md)"}

    # ---- "a second run at the reference's real dt = 1/30 for realism" (single GPU).  At this workload's dx the
    # fastest particles then move ~1.07 cells per substep: past the one-cell margin of the staged tiles, so they take the
    # slow list (k_g2p_slow) -- the number says what the margin costs outside the CFL 0.5 regime the headline uses.
    if world == 1:
        try:
            ctx.set_particles_aos(aos_host.numpy())
            dt_real = 1.0 / 30.0
            for _ in range(2):
                ctx.substep(dt_real, order=4, interp=interp, arith=capi.FAST)
            torch.cuda.synchronize(dev)
            ctx.profile_enable(True); ctx.profile_read(reset=True)
            e0.record(stream)
            for _ in range(3):
                ctx.substep(dt_real, order=4, interp=interp, arith=capi.FAST)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            pr_ = ctx.profile_read(reset=True); ctx.profile_enable(False)
            rms = e0.elapsed_time(e1) / 3
            slow = [v for k, v in pr_.items() if "k_g2p_slow" in k]
            variants["dt_1_30"] = {"value": ctx.num_particles / (rms * 1e-3), "ms_per_step": rms, "dt": dt_real,
                                   "max_displacement_cells": float(dt_real / dx),
                                   "k_g2p_slow_ms": (slow[0][0] / max(1, slow[0][1])) if slow else None,
                                   "note": "profiled steps (event pairs per launch add a few %); particles advected out of the grid leave the count"}
        except Exception as e:
            variants["dt_1_30"] = {"unavailable": repr(e)[:200]}

    # ---- cpu baseline (bounded sample, rank 0, N = 1 only) ---------------------------------------------
    cpu_baseline = None
    if host_small is not None:
        try:
            cpu = CpuHotpath(host_small, interp, args.cpu_sample)
            cpu.N = N
            v, spent = cpu.step()
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.sample_desc(),
                            "seconds": spent}
            cpu.close()
            if cpu.kind == "reference":
                # the reference is single-threaded -- its faithful number is one core's
                one = CpuHotpath(host_small, interp, min(args.cpu_sample, 20000), force_cores=1)
                one.N = N
                v1, spent1 = one.step()
                cpu_baseline["one_core"] = {"value": v1, "cores": 1, "sample": one.sample_desc(), "seconds": spent1}
                one.close()
        except Exception as e:          # the baseline leg must never take the GPU numbers down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (fp64 index arithmetic, 64-bit fixed-point P2G accumulation)", "data": "synthetic",
        "config": {"workload": args.workload, "grid": list(dims), "dx": dx, "particles": N, "cells": G,
                   "interp": args.interp, "rk_order": 4, "arith": "fast", "cfl": 0.5,
                   "parallelism": "z-slabs x%d (%s cuts: %s), halo %d layers" % (world, args.cuts, ranges, halo) if world > 1 else "single GPU",
                   "l2": "inputs (%.2f GB particles + %.2f GB fields) exceed the 126 MB L2; no flush needed"
                         % (N * 24 / 1e9, 8 * (nu + nv + nw) / 1e9)},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks.summary(), "kernels": kernels, "variants": variants,
        "stats": {k: int(v) for k, v in st.items()}, "verify": verify,
        "launch": {"graph_replays_in_timed_region": int(graph_replays), "kernels_per_step": launches_timed / args.steps,
                   "note": "gfs_substep replays a captured CUDA graph in steady state (single GPU); gpu_launches counts the kernels inside"},
    }
    # ---- sub-metrics of SURVEY 8(d): P2G only, G2P + RK4 only (from the per-kernel CUDA-event times of the timed steps),
    # and advection only through the host-pointer operator (C5: uniformly random particles, RK4 through the NEW field)
    def per_step(names):
        return sum(kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"] for k in kernels if any(nm in k for nm in names))
    n_rank = ctx.num_particles
    p2g_ms = per_step(["ExclusiveSum", "k_build_index", "k_hist", "k_scatter_sorted", "k_classify", "k_p2g_tile", "k_p2g_scatter", "k_p2g_finalize", "k_assemble"])
    g2p_ms = per_step(["k_g2p_brick", "k_g2p_advect", "k_g2p_tri", "k_g2p_slow", "k_resolve_collisions"])
    sub = {}
    if p2g_ms > 0:
        b = 24 * n_rank + 13 * G_local
        sub["p2g_only"] = {"value": n_rank / (p2g_ms * 1e-3), "unit": UNIT, "ms": p2g_ms, "algorithmic_bytes": b,
                           "hbm_frac": b / (p2g_ms * 1e-3) / 1e9 / hbm_gbs, "includes": "counting-sort scan + index, classify, splat, finalize, assemble (rank 0)"}
    if g2p_ms > 0:
        b = 48 * n_rank + 24 * G_local
        sub["g2p_rk4_only"] = {"value": n_rank / (g2p_ms * 1e-3), "unit": UNIT, "ms": g2p_ms, "algorithmic_bytes": b,
                               "hbm_frac": b / (g2p_ms * 1e-3) / 1e9 / hbm_gbs, "includes": "PIC/FLIP update + RK4 + solid test + binning for the next sort (rank 0)"}
    if world == 1:
        n_adv = 1 << 24
        rng = np.random.default_rng(12345 + 24)
        lo, span = np.float32(1.5 * dx), np.float32((min(dims) - 3) * dx)
        pos_adv = (lo + span * rng.random((n_adv, 3), dtype=np.float32)).astype(np.float32)
        ctx.profile_enable(True)
        ctx.profile_read(reset=True)
        t0 = time.perf_counter()
        ctx.advect(pos_adv, *[t.numpy() for t in new_host], dims, dx, dt, order=4, interp=interp, arith=capi.FAST)
        wall = time.perf_counter() - t0
        pa = ctx.profile_read(reset=True)
        ctx.profile_enable(False)
        k_ms = sum(v[0] for k, v in pa.items() if "k_advect" in k)
        b = 24 * n_adv + 12 * G
        sub["advect_only"] = {"particles": n_adv, "kernel_ms": k_ms, "value": n_adv / (k_ms * 1e-3) if k_ms > 0 else None, "unit": "particles/s",
                              "algorithmic_bytes": b, "hbm_frac": b / (k_ms * 1e-3) / 1e9 / hbm_gbs if k_ms > 0 else None,
                              "through_host_value": n_adv / wall, "api": "gfs_advect (host pointers, unsorted random positions, global loads)"}
    if world == 1 and not args.no_pressure:
        # stages 6-8 on the field the last substep's P2G left on the device (this workload's grid and
        # fluid cells): constant gravity, the reference's MICCG(0) with its own tolerance / iteration limit, the update.
        try:
            ctx.profile_enable(True)
            ctx.profile_read(reset=True)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.apply_body_force(capi.FIELD_P2G, (0.0, -9.8, 0.0), dt)
            with torch.cuda.stream(stream):
                ev0.record(stream)
            iters, resid = ctx.pressure_solve(capi.FIELD_P2G, dt)
            with torch.cuda.stream(stream):
                ev1.record(stream)
            ctx.apply_pressure(capi.FIELD_P2G, capi.FIELD_NEW, dt)
            torch.cuda.synchronize(dev)
            pk = ctx.profile_read(reset=True)
            ctx.profile_enable(False)
            fluid = int(st["fluid_cells"])
            solve_ms = ev0.elapsed_time(ev1)
            n_it = max(1, iters)
            kms = {k.replace("gfs::", ""): {"ms_per_launch": v[0] / max(1, v[1]), "launches": int(v[1])} for k, v in pk.items() if "press" in k or "body_force" in k}
            am = [v for k, v in kms.items() if "apply_matrix" in k]
            sub["pressure_solve"] = {
                "cells": G, "fluid_cells": fluid, "iterations": iters, "residual_max": resid, "tolerance": 1e-6, "max_iterations": 200,
                "solve_ms": solve_ms, "ms_per_iteration": solve_ms / n_it, "kernels": kms,
                "apply_matrix": {"algorithmic_bytes": 17 * G, "hbm_frac": (17 * G / (am[0]["ms_per_launch"] * 1e-3) / 1e9 / hbm_gbs) if am and am[0]["ms_per_launch"] > 0 else None,
                                 "note": "dense 7-point SpMV over all cells: search vector read (8 B) + flags (1 B) + result written (8 B) per cell"},
                "note": "PressureSolver::solve restated operation for operation (fp64, MIC(0) sweeps as tile wavefronts); "
                        "the reference's CPU time for the same stage is in submetrics.dropin.*.stage_seconds['Update Pressure Grid'] at 64^3"}
        except Exception as e:
            sub["pressure_solve"] = {"unavailable": repr(e)[:200]}
    if world == 1 and not args.no_sweep:
        ctx.close()                       # the sweep's 1 B-particle point needs the memory
        ctx = None
        torch.cuda.empty_cache()
        free_b, _total = torch.cuda.mem_get_info(dev)
        sub["advect_sweep"] = advect_sweep(local, dev, stream, hbm_gbs, [1 << 20, 1 << 22, 1 << 24, 1 << 26, 1 << 28, 1000000000], free_b - (6 << 30))
    if world == 1 and not args.no_cpu_baseline and not args.no_dropin:
        sub["dropin"] = dropin_submetric()
    line["submetrics"] = sub
    if world > 1:
        line["multi_gpu_balance"] = balance
        line["multi_gpu"] = {"transport": "peer memory (CUDA IPC, NVLink) written by gfs kernels" if args.transport == "peer"
                             else "torch.distributed batch_isend_irecv (NCCL)", "comm_bytes_per_step_rank0": comm_bytes, "particles_max_over_ranks": int(n_max),
                             "particles_mean": n_now / world, "particles_after": n_now}
    if ctx is not None:
        ctx.close()
    if world > 1 and not args.no_peer_check:
        # the CUDA-IPC peer transport against the torch.distributed (NCCL) one on a small sharded scene, across these very
        # processes: owned P2G layers, material and particle rows bit for bit after every substep (tests/peer_check.py)
        try:
            from tests import peer_check
            torch.cuda.set_stream(torch.cuda.default_stream(dev))
            ok, n_chk = peer_check.check(rank, world, local, dev, "small32" if world <= 4 else "tall64")
            line["verify"]["peer_vs_nccl"] = {"ok": bool(ok), "scene_particles": n_chk, "ranks": world}
        except Exception as e:          # never lose the measurement to the check
            line["verify"]["peer_vs_nccl"] = {"ok": False, "error": repr(e)[:300]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def advect_sweep(local, dev, stream, hbm_gbs, sizes, budget_bytes):
    """BASELINE configs[4]: RK4-only advection of N uniformly random particles through the 256^3 vortex field on the sorted
    brick path (gfs_advect_substep: index sort + trilinear brick kernel, positions only).  Algorithmic bytes 24 N + 12 G
    (SURVEY 8d).  Particles are generated straight into the context's device arrays."""
    import torch
    from gridfluidsim3d_b200 import capi, synth
    dims, dx, _ = synth.CONFIGS["splash256"]
    G = dims[0] * dims[1] * dims[2]
    ext = (dims[0] * dx, dims[1] * dx, dims[2] * dx)
    ctx = capi.Context(local, stream=stream.cuda_stream)
    ctx.domain_init(dims, dx)
    ctx.set_material(synth.border_material(dims))
    for comp_slot in (capi.FIELD_NEW,):
        fields = []
        for comp, (ni, nj, nk) in enumerate(synth.face_dims(dims)):
            i = torch.arange(ni, device=dev, dtype=torch.float32)[None, None, :]
            j = torch.arange(nj, device=dev, dtype=torch.float32)[None, :, None]
            k = torch.arange(nk, device=dev, dtype=torch.float32)[:, None, None]
            x = (i + (0.0 if comp == 0 else 0.5)) * dx
            y = (j + (0.0 if comp == 1 else 0.5)) * dx
            z = (k + (0.0 if comp == 2 else 0.5)) * dx
            fields.append(synth.vortex_t(x, y, z, ext)[comp].expand(nk, nj, ni).contiguous().reshape(-1).cpu().numpy())
        ctx.set_field(comp_slot, *fields)
    dt = synth.cfl_dt(dx)

    class _Arr:
        pass

    def dev_f32(ptr, n):
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(a, device=dev)
    rows = []
    for n in sizes:
        if n * 96 > budget_bytes:            # 80 B of resident arrays per particle + slack
            rows.append({"particles": n, "skipped": "needs %.0f GB" % (n * 96 / 1e9)})
            continue
        try:
            ctx.resize_particles(n)
            gen = torch.Generator(device=dev)
            gen.manual_seed(12345 + int(np.log2(n)))
            lo, span = 1.5 * dx, (min(dims) - 3) * dx
            chunk = 1 << 26
            for a in range(6):
                t = dev_f32(ctx.device_ptr(10 + a), n)
                for c0 in range(0, n, chunk):
                    c1 = min(n, c0 + chunk)
                    if a < 3:
                        t[c0:c1] = lo + span * torch.rand(c1 - c0, generator=gen, device=dev, dtype=torch.float32)
                    else:
                        t[c0:c1] = 0.0
            torch.cuda.synchronize()
            for _ in range(2):
                ctx.advect_substep(dt, order=4)
            ctx.sync()
            steps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                ctx.advect_substep(dt, order=4)
            e1.record(stream)
            ctx.sync()
            ms = e0.elapsed_time(e1) / steps
            b = 24 * n + 12 * G
            rows.append({"particles": n, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "particles/s", "algorithmic_bytes": b,
                         "hbm_frac": b / (ms * 1e-3) / 1e9 / hbm_gbs, "stats_particles": int(ctx.num_particles)})
        except Exception as e:
            rows.append({"particles": n, "error": repr(e)[:200]})
            break
    ctx.close()
    return {"field": "256^3 analytic vortex, CFL 0.5", "operator": "gfs_advect_substep (index sort + k_g2p_tri<advect only>, RK4, trilinear)",
            "bytes": "24 N + 12 G", "points": rows}


def dropin_submetric(n=64, frames=1):
    """FluidSimulation::update() of the UNMODIFIED reference simulator (Hello World scene of README.md:113-117 at 64^3,
    BASELINE configs[0]) three ways: its own CPU accelerator classes, the CUDA drop-in classes (host-pointer C-ABI calls
    per stage), and the device-resident stages of dropin/fluidsimulation_resident.cpp.  The libraries are the reference
    sources compiled where they lie (oracle/Makefile: ref, dropin, resident); particle-substeps/s counts every substep
    the simulator's own log reports.  Meshing, level set and output stay reference CPU code in all three (the resident
    build also runs body forces, the pressure solve, the pressure update and both extrapolations on the device), so this is
    the speed-up a user of the reference sees, not the kernel speed-up."""
    import ctypes
    from oracle import pyoracle
    from oracle.pyoracle import RefSim
    out = {"scene": "%d^3 sphere drop, %d frame(s) of 1/30 s" % (n, frames), "unit": UNIT}
    libs = (("cpu_classes", pyoracle.REF_SO, False), ("cuda_classes", pyoracle.DROPIN_SO, True), ("cuda_resident", pyoracle.RESIDENT_SO, True))
    for name, path, cuda in libs:
        if not os.path.exists(path):
            out[name] = {"unavailable": "%s not built" % os.path.basename(path)}
            continue
        try:
            lib = pyoracle.Reference(build=False, path=path)
            ctypes.CDLL(None).srand(1)
            sim = lib.sim((n, n, n), 8.0 / n)
            sim.add_fluid_sphere((4.0, 4.0, 4.0), 6.0)
            sim.add_body_force((0.0, -25.0, 0.0))
            if cuda:
                sim.set_accel(True, True)
            log_path = sim.log_path()
            if os.path.exists(log_path):
                os.remove(log_path)
            sim.initialize()
            if cuda:                      # context creation, first allocations and the first kernel loads are not the steady state
                sim.update(1.0 / 30.0)
                if os.path.exists(log_path):
                    os.remove(log_path)
            t0 = time.perf_counter()
            for _ in range(frames):
                sim.update(1.0 / 30.0)
            wall = time.perf_counter() - t0
            stages, substeps = RefSim.stage_times(log_path)
            npart = sim.n
            sim.close()
            hot = sum(stages.get(k, 0.0) for k in ("Update Fluid Cells", "Advect Velocity Field", "Update PIC/FLIP Velocities", "Advance Marker Particles"))
            grid = sum(stages.get(k, 0.0) for k in ("Apply Body Forces", "Update Pressure Grid", "Apply Pressure", "Extrapolate Fluid Velocities"))
            out[name] = {"value": npart * max(1, substeps) / wall, "seconds": wall, "substeps": substeps, "particles": npart,
                         "hot_path_stage_seconds": hot, "hot_path_value": npart * max(1, substeps) / hot if hot > 0 else None,
                         "grid_stage_seconds": grid,
                         "stage_seconds": {k: round(v, 4) for k, v in stages.items()}}
        except Exception as e:
            out[name] = {"unavailable": repr(e)[:200]}
    try:
        out["speedup_whole_update"] = out["cuda_resident"]["value"] / out["cpu_classes"]["value"]
        out["speedup_hot_path_stages"] = out["cuda_resident"]["hot_path_value"] / out["cpu_classes"]["hot_path_value"]
        out["speedup_grid_stages_6_to_9"] = out["cpu_classes"]["grid_stage_seconds"] / max(1e-9, out["cuda_resident"]["grid_stage_seconds"])
    except (KeyError, TypeError, ZeroDivisionError):
        pass
    return out


JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else any library prints to fd 1 during the
    run (NCCL's version banner, for one) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(JSON_FD if JSON_FD is not None else 1, data)


def main():
    global JSON_FD
    sys.stdout.flush()
    JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gfs", choices=["gfs", "reference"])
    ap.add_argument("--workload", default="splash256")
    ap.add_argument("--interp", default="trilinear", choices=["trilinear", "tricubic"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cuts", default="weighted", choices=["weighted", "uniform"],
                    help="N>1 slab cuts: weighted = balance particles per slab, uniform = equal layer counts")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N>1 neighbour exchange: peer = CUDA-IPC peer memory written by our kernels; nccl = torch.distributed P2P batches")
    ap.add_argument("--cpu-sample", type=int, default=40000, help="CPU-baseline particles per host thread")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the advection-only sweep 1 M .. 1 B particles (BASELINE configs[4])")
    ap.add_argument("--balance-passes", type=int, default=1, help="N>1: timing-calibration passes for the slab cuts")
    ap.add_argument("--no-balance", action="store_true", help="N>1: keep the particle-count-weighted cuts (no timing calibration pass)")
    ap.add_argument("--no-pressure", action="store_true", help="skip the pressure-solve sub-metric (stages 6-8 on the workload's grid)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the FluidSimulation::update drop-in sub-metric (64^3, CPU vs CUDA classes)")
    ap.add_argument("--no-peer-check", action="store_true", help="N>1: skip the peer-memory vs NCCL transport cross-check after the timed runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gfs" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gfs(args)


if __name__ == "__main__":
    main()
