#!/usr/bin/env python
"""bench.py -- headline benchmark of the FDTD hot path (BASELINE.json: Gcell-updates/s in double).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full leapfrog time step (source, H update, source, E update: the loop body at
main.c:770-779 of the reference) over the whole grid.  Workload = BASELINE.json configs[2],
the 1024^3 cavity in computation mode, per GPU; with N GPUs the cavity is 1024 x 1024 x 1024*N
cells, z-slab decomposed (weak scaling, one process per GPU under torchrun, NCCL halos).

Prints ONE JSON line (rank 0).  `value` is timed on the device with CUDA events recorded on
the stream the kernels are launched on (fdtd_run_timed), state resident in HBM, max over ranks.
`e2e` is the same metric through the C ABI with pinned HOST buffers: fdtd_upload_slab, K steps,
fdtd_download_slab, all inside the timed region.  `roofline` is for the slower of the two update
kernels.  `cpu_baseline` / `--impl reference` time the reference's own sequential CPU code
(oracle/_ref, the unmodified main.c) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_XY = 1024                # cells along x and y
NZ_PER_GPU = 1024          # cells along z per GPU
DX, DT = 0.001, 6e-13      # the reference's stock steps (params.txt:4-5)
BYTES_PER_CELL_HALF_STEP = 72.0   # 144 B per cell-update = 2 half-steps x 9 doubles
CPU_SAMPLE_NZ = 32         # bounded CPU sample: the same 1024 x 1024 cross-section, 32 cell planes


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def mark(self):
        """rows between the first two marks are the timed region"""
        self.marks.append(len(self.rows))

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=5)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = self.rows
        if len(self.marks) >= 2 and self.marks[1] - self.marks[0] >= 3:
            rows = self.rows[self.marks[0]:self.marks[1]]     # the timed region alone
        elif self.marks:
            rows = self.rows[max(0, self.marks[0] - 5):]      # short region: the warm-up right before it counts too
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_mem_available_bytes():
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except OSError:
        pass
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            txt = open(path).read().strip()
            if txt.isdigit():
                lim = int(txt)
                used = 0
                try:
                    used = int(open(os.path.join(os.path.dirname(path), "memory.current")).read())
                except OSError:
                    pass
                avail = min(avail, lim - used) if avail is not None else lim - used
        except OSError:
            pass
    return avail


def cpu_reference_rate(steps, warmup, nz=CPU_SAMPLE_NZ):
    """The reference's sequential CPU path on a bounded sample: the workload's 1024 x 1024
    cross-section, `nz` cell planes, computation mode; Gcell-updates/s on one core."""
    import oracle as O
    chk = O.reference() or O.restatement()
    p = O.make_params(N_XY * DX, N_XY * DX, nz * DX, DX, DT, 1e-9, 1 << 30, 1)
    assert p.dims() == (N_XY, N_XY, nz), p.dims()
    f = O.alloc_fields(*p.dims())
    t = chk.run(p, f, warmup, 0.0)
    t0 = time.perf_counter()
    chk.run(p, f, steps, t)
    dt = time.perf_counter() - t0
    cells = N_XY * N_XY * nz
    return {"value": cells * steps / dt / 1e9, "unit": "Gcell-updates/s", "cores": 1,
            "kind": chk.kind,
            "sample": f"{N_XY}x{N_XY}x{nz} cells (the workload's cross-section, {nz} of its z-planes), "
                      f"{steps} steps after {warmup} warm-up, computation mode, "
                      f"{'oracle/_ref = unmodified main.c' if chk.kind == 'reference' else 'oracle/fdtd_oracle.c'}"
                      f" gcc -std=c99 -O3, sequential like the reference",
            "seconds": dt}


def workload_config(n_gpus, nz_total=None):
    nz_total = NZ_PER_GPU * n_gpus if nz_total is None else nz_total
    which = ("BASELINE.json configs[2]: 1024^3 per GPU" if (N_XY, NZ_PER_GPU) == (1024, 1024)
             else "BASELINE.json configs[3]: 2048^3 class cavity in z-slabs" if N_XY == 2048 else "custom grid")
    state_gb = 6 * 8 * N_XY * N_XY * (nz_total / n_gpus) / 1e9
    return {"workload": f"{N_XY}x{N_XY}x{nz_total} PEC cavity, computation mode (waveguide source on), "
                        f"dx=1mm dt=0.6ps, double precision ({which})",
            "cells": N_XY * N_XY * nz_total,
            "decomposition": f"{n_gpus} z-slab(s) of {nz_total // n_gpus} planes, one process per GPU",
            "l2": f"state is {state_gb:.1f} GB per GPU, far larger than the 126 MB L2; no flush needed"}


def run_reference_arm(args):
    """The reference's own sequential CPU code (oracle/_ref = the unmodified main.c) on the box's host
    cores, rank 0 only.  At N = 1 it runs the WHOLE 1024^3 workload when the host has the memory and
    the W + K steps fit in about six minutes (estimated from a 2-step probe on a 32-plane sample);
    otherwise, and for N > 1 (a 1024 x 1024 x 1024*N cavity does not fit any host), it runs the
    workload's cross-section with fewer z-planes and config.workload says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    t0 = time.perf_counter()
    nz_full = NZ_PER_GPU * args.gpus
    probe = cpu_reference_rate(2, 1)
    est_step_s = N_XY * N_XY * nz_full / (probe["value"] * 1e9)
    need = 6 * 8 * (N_XY + 1) * (N_XY + 1) * (nz_full + 1)
    avail = host_mem_available_bytes()
    full = (avail is not None and need < 0.8 * avail and est_step_s * (steps + warm) < 360.0
            and os.environ.get("FDTD_BENCH_REF_SAMPLE") != "1")
    nz = nz_full if full else CPU_SAMPLE_NZ
    r = cpu_reference_rate(steps, warm, nz=nz)
    cfg = workload_config(args.gpus)
    if not full:
        cfg["workload"] = (f"SAMPLE of {cfg['workload']}: the same {N_XY}x{N_XY} cross-section with {nz} of its "
                           f"{nz_full} z-planes ({N_XY * N_XY * nz} cells), because "
                           + (f"the whole cavity needs {need / 1e9:.0f} GB of host memory" if avail is None or need >= 0.8 * avail
                              else f"{steps + warm} steps of the whole cavity would take {est_step_s * (steps + warm):.0f} s on one core"))
        cfg["cells_timed"] = N_XY * N_XY * nz
    line = {"impl": "reference", "metric": "cell_updates_per_second", "value": r["value"],
            "unit": "Gcell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (zero fields driven by the waveguide source)",
            "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


WORKLOADS = {
    # name: (nxy, nz per GPU or None, nz total or None, what)
    "cube1024": (1024, 1024, None, "BASELINE.json configs[2], weak scaling: 1024 x 1024 x 1024*N"),
    "cube2048": (2048, 256, None, "BASELINE.json configs[3], weak scaling: 2048 x 2048 x 256*N (2048^3 at N = 8)"),
    "strong2048": (2048, None, 1024, "BASELINE.json configs[3], strong scaling: 2048 x 2048 x 1024 split over N GPUs"),
}


def wire(F, dist, ctx, rank, world, transport):
    """Connect the slabs: peer memory (CUDA IPC + flags) when every rank can, else NCCL."""
    import torch
    if world == 1:
        return "none"
    if transport in ("auto", "peer"):
        ok = 1
        try:
            blob = ctx.peer_export()
        except F.FdtdError as e:
            blob, ok = repr(e), 0
        blobs = [None] * world
        dist.all_gather_object(blobs, blob)
        if all(isinstance(b, bytes) for b in blobs):
            try:
                ctx.peer_connect(blobs)
            except F.FdtdError as e:
                ok, blob = 0, repr(e)
        else:
            ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            return "peer"
        if transport == "peer":
            raise SystemExit(f"--transport peer: peer memory is not available on rank {rank}: {blob}")
        # a rank that did connect cannot be re-wired: start over with a fresh context (caller)
        return None
    box = [F.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0])
    return "nccl"


def main():
    global NZ_PER_GPU, N_XY
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["dumps50"],
                    help="named workloads of BASELINE.json (default cube1024); dumps50 = configs[4]")
    ap.add_argument("--nz", type=int, default=None, help="z cells per GPU (default: the 1024^3 workload)")
    ap.add_argument("--nxy", type=int, default=None, help="cells along x and y (default 1024)")
    ap.add_argument("--nz-total", type=int, default=None,
                    help="fixed total z cells split over the GPUs (strong scaling, e.g. 2048 with --nxy 2048)")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"],
                    help="how halo planes travel for N > 1: peer memory over NVLink (CUDA IPC) or NCCL send/recv")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true")
    ap.add_argument("--serial-e2e", action="store_true", help="e2e as upload, run, download in sequence (round 1)")
    ap.add_argument("--opt", action="append", default=[], help="context option key=value (tuning)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 100
    if args.workload == "dumps50":
        return run_dumps50(args)
    if args.workload:
        nxy, nz, nz_total, _ = WORKLOADS[args.workload]
        args.nxy = args.nxy or nxy
        args.nz = args.nz or nz
        args.nz_total = args.nz_total or nz_total
    N_XY = args.nxy or N_XY
    NZ_PER_GPU = args.nz or NZ_PER_GPU
    if args.impl == "reference":
        return run_reference_arm(args)

    W = max(args.warmup, 3)
    K = args.steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus != 1 and world == 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun (one process per GPU): python -m torch.distributed.run "
                     f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus}")
        args.gpus = world

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import fdtd_b200 as F

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nz_total = NZ_PER_GPU * world
    if args.nz_total is not None:
        nz_total = args.nz_total
        NZ_PER_GPU = nz_total // world
    p = F.make_params(N_XY * DX, N_XY * DX, nz_total * DX, DX, DT, 1e-9, 1 << 30, 1)
    assert p.dims() == (N_XY, N_XY, nz_total), p.dims()

    def make_ctx():
        c = F.Context(p, device=local, rank=rank, nranks=world)
        for kv in args.opt:
            k, v = kv.split("=")
            c.set_option(k, int(v))
        return c

    with ClockSampler(local) as clocks:          # sampled from before the warm-up: short timed regions still get rows
        ctx = make_ctx()
        transport = wire(F, dist, ctx, rank, world, args.transport)
        if transport is None:                    # peer memory unavailable somewhere: everybody falls back to NCCL
            ctx.close()
            ctx = make_ctx()
            transport = wire(F, dist, ctx, rank, world, "nccl")
        cells_local = N_XY * N_XY * (ctx.k1 - ctx.k0)
        cells_total = N_XY * N_XY * nz_total

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ctx.sync()

        # ---- device-resident timing -------------------------------------------------------------
        ctx.fill_test_pattern(20261018)
        t = ctx.run(W, 0.0)
        barrier()
        launches0 = ctx.get_option("launches")
        clocks.mark()
        barrier()
        t, total_ms, h_ms, e_ms = ctx.run_timed(K, t)
        barrier()
        clocks.mark()
    launches = ctx.get_option("launches") - launches0
    ms = torch.tensor([total_ms, h_ms, e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms, h_ms, e_ms = (float(x) for x in ms.cpu())
    value = cells_total * K / (total_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peak, peak_src = peak_hbm()
    kopts = {k: ctx.get_option(k) for k in ("kernel", "strip", "kchunk", "warps_x", "warps_y", "stages")}
    kernel_id = kopts["kernel"]
    rolling = bool(ctx.get_option("rolling"))
    cap = None if rolling else traffic_for(kopts)    # no capture of the in-place (rolling window) form
    steps_per_launch = 1
    if kernel_id == 4:
        # ONE launch advances every cell by TWO time steps (an odd K ends with one single-step sweep)
        steps_per_launch = 2
        n_launch = (K + 1) // 2
        dom_ms = h_ms / n_launch
        dom_name = (f"k_step2_tma<{kopts['warps_y']}> (two time steps per sweep, TMA-staged, {kopts['stages']} stages, "
                    f"{kopts['kchunk']} planes per block)")
        alg_bytes = 2 * 2 * BYTES_PER_CELL_HALF_STEP * cells_local   # 144 B per cell-update, two of them
        min_bytes = 96.0 * cells_local                               # six arrays in, six arrays out, once per sweep
    elif kernel_id >= 2:
        # fused step: ONE launch advances every cell by a full time step (H and E)
        dom_ms = h_ms / K
        dom_name = (f"k_step_fused_tma<{kopts['strip']},{kopts['warps_x']},{kopts['warps_y']}> (H+E in one sweep, TMA-staged, "
                    f"{kopts['stages']} stages)" if kernel_id == 3 else "k_step_fused (H+E in one sweep)")
        alg_bytes = 2 * BYTES_PER_CELL_HALF_STEP * cells_local   # 144 B per cell-update
        min_bytes = 96.0 * cells_local                           # what a fused sweep has to move: 6 reads + 6 writes
    else:
        dom_ms = max(h_ms, e_ms) / K
        dom_name = "k_update_h_march (H half-step)" if h_ms >= e_ms else "k_update_e_march (E half-step)"
        alg_bytes = BYTES_PER_CELL_HALF_STEP * cells_local
        min_bytes = alg_bytes
    traffic = cap["bytes_per_cell"] * cells_local if cap else None
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                "traffic_source": cap["source"] if cap else None,
                "traffic_over_minimum": traffic / min_bytes if traffic else None,
                "dram_gbs_from_traffic": traffic / (dom_ms * 1e-3) / 1e9 if traffic else None,
                "dram_frac_of_peak": traffic / (dom_ms * 1e-3) / 1e9 / peak if traffic else None,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dom_ms,
                "time_steps_per_launch": steps_per_launch,
                "bytes_per_cell_update_basis": 144,
                "minimum_bytes_per_launch": min_bytes,
                "achieved_on_minimum": min_bytes / (dom_ms * 1e-3) / 1e9,
                "frac_on_minimum": min_bytes / (dom_ms * 1e-3) / 1e9 / peak,
                "note": ("achieved/frac use SURVEY.md 8(d)'s 144 B per cell-update (the split H + E half-steps). A sweep "
                         "that fuses H and E -- and here two whole time steps -- moves far fewer bytes per cell-update "
                         "(48 B at best for two steps), so frac exceeds 1 by construction; the HBM utilisation is "
                         "dram_frac_of_peak (bytes ncu saw the kernel move / its duration / peak) and frac_on_minimum "
                         "(the bytes it cannot avoid)" if kernel_id >= 2 else "two launches per step; 72 B per cell and half-step"),
                "step_frac_of_roofline": (2 * BYTES_PER_CELL_HALF_STEP * cells_local / (total_ms / K * 1e-3) / 1e9) / peak}

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    e2e = None
    if not args.no_e2e:
        shapes = ctx.slab_shapes()
        need = sum(8 * a * b * c for a, b, c in shapes.values())
        # every rank pins its own slab on the same host: decide collectively on the node's total
        avail = host_mem_available_bytes()
        av = torch.tensor([float(avail if avail is not None else 1e18)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(av, op=dist.ReduceOp.MIN)
        avail = float(av.cpu())
        if need * world > 0.6 * avail:
            e2e = {"value": None, "unit": "Gcell-updates/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "skipped": f"{world} rank(s) x {need / 1e9:.1f} GB of pinned host buffers do not fit the "
                              f"{avail / 1e9:.0f} GB of host memory available on this node"}
        else:
            host = F.PinnedArrays(p, shapes=shapes, device=local)   # pages from the NUMA node of this rank's GPU
            ctx.fill_test_pattern(7)
            ctx.download_slab(host.arrays)          # synthetic input now lives in HOST memory
            barrier()
            t0 = time.perf_counter()
            if args.serial_e2e:
                ctx.upload_slab(host.arrays)        # host -> HBM
                ctx.run(K, 0.0)                     # K steps
                ctx.download_slab(host.arrays)      # HBM -> host (blocks until complete)
                what = "fdtd_upload_slab + fdtd_run(K) + fdtd_download_slab in sequence"
            else:
                ctx.run_hosted(host.arrays, K, 0.0)  # the same three, pipelined over z-chunks inside the library
                what = ("fdtd_run_hosted: host arrays in, K steps, host arrays out; upload, stepping and download "
                        "overlap chunk by chunk along z where the slab layout allows it")
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.cpu())
            e2e = {"value": cells_total * K / dt / 1e9, "unit": "Gcell-updates/s",
                   "h2d_bytes_per_step": need * world / K, "d2h_bytes_per_step": need * world / K,
                   "seconds": dt, "what": what + "; pinned host arrays in the reference's dense layout, wall clock, "
                                                 "max over ranks",
                   "host_numa": dict(F.host_numa_info(local), policy=os.environ.get("FDTD_B200_HOST_NUMA", "near")),
                   "checksum_after": None}
            if not args.no_selfcheck:
                # the pipelined path must leave the same state in the host arrays as the plain one
                ctx.upload_slab(host.arrays)
                e2e["checksum_after"] = ctx.checksum()
            host.close()

    # ---- self-check at the benchmarked size: default kernel vs the plain one-thread-per-cell operators
    selfcheck = None
    if not args.no_selfcheck:
        default_opts = {k: ctx.get_option(k) for k in ("kernel", "strip", "kchunk", "warps_x", "warps_y", "stages")}
        sums = {}
        for label, opts in (("default", default_opts), ("plain", {"kernel": 0})):
            for k, v in opts.items():
                ctx.set_option(k, v)
            ctx.fill_test_pattern(7)
            ctx.run(K, 0.0)
            mine = ctx.checksum()
            if world > 1:
                allsums = [None] * world
                dist.all_gather_object(allsums, mine)
            else:
                allsums = [mine]
            sums[label] = [sum(s[a] for s in allsums) % (1 << 64) for a in range(6)]
        selfcheck = {"equal": sums["default"] == sums["plain"], "kernels": [default_opts["kernel"], 0], "steps": K,
                     "what": "fdtd_checksum of all six arrays, summed over the slabs, after K steps from "
                             "fdtd_fill_test_pattern(7): benchmarked kernel vs kernel 0 (oracle-pinned operators)",
                     "checksums": [f"{x:016x}" for x in sums["default"]]}
        if e2e and e2e.get("checksum_after") is not None:
            if world > 1:
                allsums = [None] * world
                dist.all_gather_object(allsums, e2e["checksum_after"])
            else:
                allsums = [e2e["checksum_after"]]
            tot = [sum(s[a] for s in allsums) % (1 << 64) for a in range(6)]
            e2e["checksum_after"] = None
            e2e["matches_selfcheck"] = tot == sums["plain"]
        if e2e:
            e2e.pop("checksum_after", None)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_rate(20, 2)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "cell_updates_per_second", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
                "scaling": "strong" if args.nz_total is not None else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (hash-pattern fields in HBM, waveguide source on)",
                "config": dict(workload_config(world, nz_total), kernel=kopts, transport=transport,
                               fallback=bool(ctx.get_option("fallback")), rolling_window=bool(ctx.get_option("rolling"))),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "selfcheck": selfcheck,
                "gpu_launches": launches, "clocks": clocks.summary(), "hbm_bytes_per_gpu": ctx.info()["hbm_bytes"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()      # nobody unmaps a neighbour's memory while it may still be in use
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_dumps50(args):
    """BASELINE.json configs[4]: a run with field dumps every 50 steps through fdtd_propagate, against
    the same run without dumps (one GPU).  The sink only counts what it is handed (the pinned buffer is
    not copied again), so what is timed is the library's dump pipeline: aggregation on the compute
    stream, D2H on a side stream through two pinned buffers, hand-over from a writer thread.  The
    pipeline hides a dump behind the following steps as long as PCIe can drain it in that time;
    `pcie_floor_seconds` says how long the dumped bytes take at the D2H rate measured here."""
    import ctypes as C
    import fdtd_b200 as F
    out = {}
    for n in (512, 1024):
        steps = 400 if n == 512 else 150
        sim_t = (steps - 0.5) * DT
        res = {}
        d2h_gbs = None
        for label, rate, dumps in (("without", 1 << 30, False), ("with", 50, True)):
            p = F.make_params(n * DX, n * DX, n * DX, DX, DT, sim_t, rate, 1)
            assert p.dims() == (n, n, n)
            seen = {"bytes": 0, "files": 0}

            def count_variable(user, name, data, count):
                seen["bytes"] += count * 8
                return 0

            def count_file(user, it, dims, k0):
                seen["files"] += 1
                return 0

            sink = F.DumpSink(None, F._BEGIN(count_file), F._VARIABLE(count_variable), F._END(lambda u: 0))
            try:
                with F.Context(p, device=0) as ctx:
                    st, tc = C.c_size_t(), C.c_double()
                    for rep in range(2):          # the first call allocates scratch and pins the buffers
                        seen.update(bytes=0, files=0)
                        t0 = time.perf_counter()
                        F._check(F.lib.fdtd_propagate(ctx._h, C.byref(sink) if dumps else None, C.byref(st), C.byref(tc)))
                        ctx.sync()
                        dt = time.perf_counter() - t0
                    res[label] = {"seconds": dt, "steps": int(st.value), "dump_bytes": seen["bytes"], "dumps": seen["files"],
                                  "gcell_s": n ** 3 * int(st.value) / dt / 1e9}
                    if d2h_gbs is None:
                        host = F.PinnedArrays(p)
                        ctx.download_slab(host.arrays)
                        t0 = time.perf_counter()
                        ctx.download_slab(host.arrays)
                        d2h_gbs = sum(a.nbytes for a in host.arrays.values()) / (time.perf_counter() - t0) / 1e9
                        host.close()
            except F.FdtdError as e:
                res[label] = {"error": str(e)}
        if "seconds" in res.get("with", {}) and "seconds" in res.get("without", {}):
            res["ratio"] = res["with"]["seconds"] / res["without"]["seconds"]
            res["d2h_gbs"] = d2h_gbs
            res["pcie_floor_seconds"] = res["with"]["dump_bytes"] / 1e9 / d2h_gbs
            res["with_over_max_of_compute_and_pcie"] = res["with"]["seconds"] / max(res["without"]["seconds"],
                                                                                  res["pcie_floor_seconds"])
        out[f"{n}^3"] = res
    print(json.dumps({"workload": "dumps50 (BASELINE.json configs[4]): dumps every 50 steps, counting sink",
                      "dump_overlap": out}), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, per cell, from the
# ncu --set full captures under profiles/ (taken at 1024 x 1024 x 256 cells so that ncu's save/restore
# of the state stays small; per-cell traffic does not depend on the plane count).  Keyed by the launch
# configuration (kernel, strip, warps_x, warps_y, stages): a configuration without a capture reports
# traffic = null -- never another kernel's bytes.
_CELLS_NCU = 1024 * 1024 * 256
TRAFFIC_TABLE = {
    "k1_s2_wx2_wy2": {"bytes_per_cell": 20.022561e9 / _CELLS_NCU,
                      "source": "profiles/r01_split_ncu_full_raw.csv (k_update_h_march<2>, 20.02 GB per launch at 1024x1024x256)"},
    "k3_s2_wx4_wy2_st3": {"bytes_per_cell": (14.695572e9 + 12.960834e9) / _CELLS_NCU,
                          "source": "profiles/r01_tma_ncu_full_raw.csv (k_step_fused_tma<2,4,2>, 128x4 tile, 3 stages: 14.70 GB read + "
                                    "12.96 GB written per launch at 1024x1024x256)"},
}
try:
    with open(os.path.join(ROOT, "profiles", "traffic_table.json")) as _fh:   # written from tools/ncu_traffic.py output
        for _e in json.load(_fh):
            TRAFFIC_TABLE[_e["key"]] = {"bytes_per_cell": _e["bytes_per_cell"], "source": _e["source"]}
except (OSError, ValueError, KeyError, TypeError):
    pass


def traffic_key(opts):
    k = opts["kernel"]
    if k == 4:     # the chunk length decides how much of the tile overlap L2 absorbs: part of the key
        return f"k4_wy{opts['warps_y']}_st{opts['stages']}_kc{opts['kchunk']}"
    if k == 3:
        return f"k3_s{opts['strip']}_wx{opts['warps_x']}_wy{opts['warps_y']}_st{opts['stages']}"
    return f"k{k}_s{opts['strip']}_wx{opts['warps_x']}_wy{opts['warps_y']}"


def traffic_for(opts):
    return TRAFFIC_TABLE.get(traffic_key(opts))


if __name__ == "__main__":
    main()
