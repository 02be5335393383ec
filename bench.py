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
        self.index, self.rows, self.proc = index, [], None

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
        for r in self.rows:
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


def workload_config(n_gpus):
    which = ("BASELINE.json configs[2]: 1024^3 per GPU" if (N_XY, NZ_PER_GPU) == (1024, 1024)
             else "BASELINE.json configs[3]: 2048^3 class cavity in z-slabs" if N_XY == 2048 else "custom grid")
    return {"workload": f"{N_XY}x{N_XY}x{NZ_PER_GPU * n_gpus} PEC cavity, computation mode (waveguide source on), "
                        f"dx=1mm dt=0.6ps, double precision ({which})",
            "cells": N_XY * N_XY * NZ_PER_GPU * n_gpus,
            "decomposition": f"{n_gpus} z-slab(s) of {NZ_PER_GPU} planes, one process per GPU",
            "l2": "state is 51.6 GB per GPU (held twice by the fused step), far larger than the 126 MB L2; no flush needed"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warm = max(0, args.warmup)
    t0 = time.perf_counter()
    r = cpu_reference_rate(steps, warm)
    line = {"impl": "reference", "metric": "cell_updates_per_second", "value": r["value"],
            "unit": "Gcell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (zero fields driven by the waveguide source)",
            "config": workload_config(args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def main():
    global NZ_PER_GPU, N_XY
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nz", type=int, default=NZ_PER_GPU, help="z cells per GPU (default: the 1024^3 workload)")
    ap.add_argument("--nxy", type=int, default=N_XY, help="cells along x and y (default 1024)")
    ap.add_argument("--nz-total", type=int, default=None,
                    help="fixed total z cells split over the GPUs (strong scaling, e.g. 2048 with --nxy 2048)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="context option key=value (tuning)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 100
    if args.impl == "reference":
        return run_reference_arm(args)

    NZ_PER_GPU = args.nz
    N_XY = args.nxy
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
    ctx = F.Context(p, device=local, rank=rank, nranks=world)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    if world > 1:
        box = [F.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0])
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
    with ClockSampler(local) as clocks:
        barrier()
        t, total_ms, h_ms, e_ms = ctx.run_timed(K, t)
        barrier()
    launches = ctx.get_option("launches") - launches0
    ms = torch.tensor([total_ms, h_ms, e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms, h_ms, e_ms = (float(x) for x in ms.cpu())
    value = cells_total * K / (total_ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel -----------------------------------------------------
    peak, peak_src = peak_hbm()
    kernel_id = ctx.get_option("kernel")
    if kernel_id >= 2:
        # fused step: ONE launch advances every cell by a full time step (H and E)
        dom_ms = h_ms / K
        dom_name = "k_step_fused_tma (H+E in one sweep, TMA-staged)" if kernel_id == 3 else "k_step_fused (H+E in one sweep)"
        alg_bytes = 2 * BYTES_PER_CELL_HALF_STEP * cells_local   # 144 B per cell-update
        min_bytes = 96.0 * cells_local                           # what a fused sweep has to move: 6 reads + 6 writes
        traffic = TRAFFIC_FUSED_BYTES_PER_CELL * cells_local if TRAFFIC_FUSED_BYTES_PER_CELL else None
        traffic_src = TRAFFIC_FUSED_SOURCE
    else:
        dom_ms = max(h_ms, e_ms) / K
        dom_name = "k_update_h_march (H half-step)" if h_ms >= e_ms else "k_update_e_march (E half-step)"
        alg_bytes = BYTES_PER_CELL_HALF_STEP * cells_local
        min_bytes = alg_bytes
        traffic = TRAFFIC_BYTES_PER_CELL_HALF_STEP * cells_local
        traffic_src = TRAFFIC_SOURCE
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": dom_ms,
                "bytes_per_cell_update_basis": 144,
                "fused_minimum_bytes_per_launch": min_bytes,
                "achieved_on_fused_minimum": min_bytes / (dom_ms * 1e-3) / 1e9,
                "frac_on_fused_minimum": min_bytes / (dom_ms * 1e-3) / 1e9 / peak,
                "note": ("achieved/frac use SURVEY.md 8(d)'s 144 B per cell-update (the split H + E half-steps); "
                         "the fused sweep reads and writes each of the six arrays once per step = 96 B per "
                         "cell-update, so frac can exceed 1: frac_on_fused_minimum is the honest HBM utilisation"
                         if kernel_id >= 2 else "two launches per step; 72 B per cell and half-step"),
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
            host = F.PinnedArrays(p, shapes=shapes)
            ctx.fill_test_pattern(7)
            ctx.download_slab(host.arrays)          # synthetic input now lives in HOST memory
            barrier()
            t0 = time.perf_counter()
            ctx.upload_slab(host.arrays)            # host -> HBM
            tt = ctx.run(K, 0.0)                    # K steps
            ctx.download_slab(host.arrays)          # HBM -> host (blocks until complete)
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.cpu())
            e2e = {"value": cells_total * K / dt / 1e9, "unit": "Gcell-updates/s",
                   "h2d_bytes_per_step": need * world / K, "d2h_bytes_per_step": need * world / K,
                   "seconds": dt, "what": "fdtd_upload_slab + fdtd_run(K) + fdtd_download_slab, pinned host arrays "
                                          "in the reference's dense layout, wall clock, max over ranks"}
            host.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_rate(20, 2)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "cell_updates_per_second", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
                "scaling": "strong" if args.nz_total is not None else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (hash-pattern fields in HBM, waveguide source on)",
                "config": dict(workload_config(world),
                               kernel={k: ctx.get_option(k) for k in ("kernel", "strip", "kchunk", "warps_x", "warps_y", "stages")}),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
                "clocks": clocks.summary(), "hbm_bytes_per_gpu": ctx.info()["hbm_bytes"]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, per cell, from
# the ncu --set full capture under profiles/ (taken at 1024 x 1024 x 256 cells so that ncu's
# save/restore of the state stays small; per-cell traffic does not depend on the plane count)
TRAFFIC_BYTES_PER_CELL_HALF_STEP = 20.022561e9 / (1024 * 1024 * 256)
TRAFFIC_SOURCE = "profiles/r01_split_ncu_full_raw.csv (k_update_h_march<2>, 20.02 GB per launch at 1024x1024x256)"
# same for the fused TMA step (None until captured)
TRAFFIC_FUSED_BYTES_PER_CELL = (14.695572e9 + 12.960834e9) / (1024 * 1024 * 256)
TRAFFIC_FUSED_SOURCE = ("profiles/r01_tma_ncu_full_raw.csv (k_step_fused_tma<2,4,2>, i.e. the 128x4 tile: 14.70 GB read + "
                        "12.96 GB written per launch at 1024x1024x256; the fused sweep's minimum is 96 B x cells = "
                        "25.77 GB).  The default tile became 32x8 (k_step_fused_tma<1,1,8>) at the very end of round 1, "
                        "after this capture; its traffic has not been re-captured")

if __name__ == "__main__":
    main()
