#!/usr/bin/env python
"""Benchmark of the powspec hot path (BASELINE.json metric: P_ell(k) wall-time &
particles/s assigned, 1024^3 TSC interlaced).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one full pass of the replaced span, genr_mesh() + powspec(): mass
assignment of the whole catalogue (both interlaced fields), the r2c FFTs, and
the fused multipole binning, ending with the P_ell(k) arrays on the host.

* value   : particles/s with the catalogue already resident in HBM (generated on
            the device), CUDA-event timed, max over ranks.
* e2e     : the same through the host API with the catalogue in (pinned) HOST
            memory — H2D of the 32-byte particle records inside the timed region,
            P_ell(k) read back.
* roofline: the assignment kernel (the dominant hand-written kernel),
            algorithmic bytes 32 N + F Ntot s   over its CUDA-event
            time, against the measured HBM copy bandwidth.
* cpu_baseline / --impl reference: the UNMODIFIED reference genr_mesh()+powspec()
  (oracle/_ref, OpenMP on all host cores, FFT through the repo's FFTW-API shim)
  on a bounded, scaled-down sample of the same workload.

N > 1: one process per GPU (torchrun), each rank transforms its own independent
catalogue of the same size (weak scaling; no data-path collective: the bins of
different mocks are never combined).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(npart=10 ** 8, box=1000.0, ng=1024, assign="TSC", interlace=True,
               poles=(0, 2, 4), kbin=0.01,
               desc="periodic box, 1e8 uniform particles, 1024^3 mesh, TSC + interlacing, P_0/P_2/P_4"),
    # BASELINE.json configs[0] (C1)
    "c1": dict(npart=10 ** 6, box=1000.0, ng=256, assign="CIC", interlace=False,
               poles=(0, 2), kbin=0.01,
               desc="periodic box, 1e6 uniform particles, 256^3 mesh, CIC, no interlacing, P_0/P_2"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def reference_sample(workload, cores):
    """Scaled-down twin of the workload (same particles per cell, same scheme /
    interlacing / multipoles / dk) sized for ~10-30 s of CPU work."""
    w = WORKLOADS[workload]
    if w["ng"] <= 256:
        return dict(w), "full workload"
    ng = 512 if cores >= 16 else 256
    scale = (ng / w["ng"]) ** 3
    s = dict(w)
    s["ng"] = ng
    s["npart"] = int(round(w["npart"] * scale))
    s["box"] = w["box"] * ng / w["ng"]
    return s, (f"scaled twin: {s['npart']} particles, {ng}^3 mesh, box {s['box']:g} "
               f"(same particles/cell, scheme, interlacing, multipoles, dk as the full workload)")


def run_reference(workload, steps, warmup):
    """The reference's own CPU implementation (oracle/_ref when it was built from
    /root/reference, else the C restatement) on all host cores."""
    from oracle import have_ref, load_oracle
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    kind = "reference" if have_ref() else "port"
    orc = load_oracle("ref" if have_ref() else "port")
    s, sample_desc = reference_sample(workload, cores)
    rng = np.random.default_rng(1)
    cat = np.empty((s["npart"], 4))
    cat[:, :3] = rng.random((s["npart"], 3)) * s["box"]
    cat[:, 3] = 1.0
    kw = dict(ng=s["ng"], assign=s["assign"], interlace=s["interlace"], poles=s["poles"],
              box=s["box"], kbin=s["kbin"])
    times, tm, tp = [], [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = orc.run(cat, **kw)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt); tm.append(r.t_mesh); tp.append(r.t_pk)
    t = float(np.mean(times)) if times else float("nan")
    return dict(value=s["npart"] / t, unit="particles/s", cores=cores, kind=kind, sample=sample_desc,
                backend=orc.backend, s_per_step=t, t_genr_mesh_s=float(np.mean(tm)),
                t_powspec_s=float(np.mean(tp)), npart=s["npart"], ng=s["ng"])


def bench_slab(args, ctx, conf, w, world, rank, local_rank, config, barrier):
    """ONE mesh over all ranks: total problem fixed (strong scaling)."""
    import torch
    import torch.distributed as dist

    from powspec_b200.distributed import GpuSlabEngine, TorchComm, slab_power

    class NoComm:
        size, rank = 1, 0
    comm = TorchComm() if world > 1 else NoComm()
    eng = GpuSlabEngine(ctx, conf, world, rank)
    n_total = w["npart"]
    n_loc = n_total // world
    ptr, _ = ctx.generate_catalog(n_loc, w["box"], kind=args.kind, seed=1 + rank)
    share = torch.empty((n_loc, 4), dtype=torch.float64, device="cuda")
    import ctypes
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(share.data_ptr()), ctypes.c_void_p(ptr),
                                           ctypes.c_size_t(n_loc * 32), 3)
    ctx.free_catalog((ptr, n_loc))

    def step():
        return slab_power(eng, comm, [share], [float(n_loc * world)])

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    from powspec_b200.distributed import PROF
    PROF.t.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pk = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    if rank == 0:
        if PROF.on:
            nrun = args.steps
            print("slab stage profile (ms per step, rank 0):",
                  {k: round(1e3 * v / nrun, 2) for k, v in PROF.t.items()}, file=sys.stderr)
        config = dict(config)
        config["parallelism"] = f"one {w['ng']}^3 mesh x-slab-decomposed over {world} GPU(s)"
        config["npart_total"] = n_loc * world
        print(json.dumps({"metric": "particles_per_second_P_ell_1024_TSC_interlaced",
                          "value": n_loc * world / (ms_step * 1e-3), "unit": "particles/s",
                          "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64" if args.precision == 8 else "f32",
                          "data": "synthetic", "config": config, "mode": "slab",
                          "P0_first_bins": [float(x) for x in pk.pl[0][0][:3]]}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--npart", type=int, default=None, help="override the particle number (debug)")
    ap.add_argument("--ng", type=int, default=None, help="override the mesh size (debug)")
    ap.add_argument("--precision", type=int, default=8, choices=[4, 8])
    ap.add_argument("--kind", type=int, default=0, help="0 uniform, 1 clustered catalogue")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="context option name=value (ablations)")
    ap.add_argument("--mode", default="replica", choices=["replica", "slab"],
                    help="N>1: 'replica' = one independent catalogue per GPU (weak scaling, default); "
                         "'slab' = ONE mesh x-slab-decomposed over the GPUs (strong scaling, NCCL "
                         "all-to-all transpose + halo exchange + allreduce)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = dict(WORKLOADS[args.workload])
    if args.npart:
        w["npart"] = args.npart
    if args.ng:
        w["ng"] = args.ng
    config = {"workload": w["desc"] if not (args.npart or args.ng) else
              f"DEBUG override: {w['npart']} particles, {w['ng']}^3", "npart_per_gpu": w["npart"],
              "ng": w["ng"], "box": w["box"], "assign": w["assign"], "interlace": w["interlace"],
              "poles": list(w["poles"]), "kbin": w["kbin"],
              "catalogue": "uniform" if args.kind == 0 else "clustered",
              "parallelism": f"{world} independent catalogue(s), one per GPU" if world > 1 else "single GPU",
              "l2_policy": "inputs (3.2 GB particles, 17 GB meshes) far larger than the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(args.workload, args.steps, max(args.warmup, 1))
        line = {"impl": "reference", "metric": "particles_per_second_P_ell_1024_TSC_interlaced",
                "value": r["value"], "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "particles/s", "cores": r["cores"],
                                 "kind": r["kind"], "sample": r["sample"], "backend": r["backend"],
                                 "t_genr_mesh_s": r["t_genr_mesh_s"], "t_powspec_s": r["t_powspec_s"]},
                "e2e": {"value": r["value"], "unit": "particles/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    import powspec_b200
    from powspec_b200.api import Cata, Conf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the powspec_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = powspec_b200.Context(local_rank)
    for kv in args.opt:
        name, val = kv.split("=")
        ctx.set_option(name, int(val))
    if args.opt:
        config["options"] = args.opt
    n = w["npart"]
    conf = Conf(ndata=1, issim=True, bsize=(w["box"],) * 3, gsize=w["ng"],
                assign=powspec_b200.powspec_assign_names.index(w["assign"]), intlace=w["interlace"],
                poles=tuple(w["poles"]), kbin=w["kbin"], isauto=(True, False), iscross=False,
                precision=args.precision, device=local_rank)
    cat_dev = ctx.generate_catalog(n, w["box"], kind=args.kind, seed=1 + rank)
    cata_dev = Cata(data=[cat_dev], wdata=[float(n)])

    def step(cata):
        mesh = ctx.genr_mesh(conf, cata)
        return ctx.powspec(conf, cata, mesh)

    if args.mode == "slab":
        return bench_slab(args, ctx, conf, w, world, rank, local_rank, config, barrier)

    def timed(cata, steps, warmup):
        for _ in range(warmup):
            step(cata)
        barrier()
        stage = {}
        launches = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        pk = None
        for _ in range(steps):
            pk = step(cata)
            launches += pk.launches
            for k_, v in pk.timings_ms.items():
                stage[k_] = stage.get(k_, 0.0) + v
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall * 1e3, {k_: v / steps for k_, v in stage.items()}, launches, pk

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, wall_ms, stages, launches, pk = timed(cata_dev, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- end to end: catalogue in pinned host memory, H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, 4), dtype=torch.float64, pin_memory=True)
        ctx.L.psb_copy_to_host(ctx.h, host.data_ptr(), cat_dev[0], n * 32)
        cata_host = Cata(data=[host], wdata=[float(n)])
        ms_e2e, _, stages_e2e, _, pk_e = timed(cata_host, args.steps, max(args.warmup, 3))
        ms_e2e_step = ms_e2e / args.steps
        d2h = (2 + 4 * pk_e.nl) * pk_e.nbin * 8 + 6 * 8 * 148 * 8
        e2e = {"value": world * n / (ms_e2e_step * 1e-3), "unit": "particles/s",
               "ms_per_step": ms_e2e_step, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": d2h,
               "host_memory": "pinned", "stages_ms": stages_e2e}
        # the same from PAGEABLE host memory (what the reference's C host hands over:
        # plain malloc), staged through pinned buffers by the library
        if world == 1:
            pageable = host.numpy().copy()
            cata_pg = Cata(data=[pageable], wdata=[float(n)])
            ms_pg, _, stages_pg, _, _ = timed(cata_pg, max(1, args.steps // 2), 1)
            e2e["pageable_ms_per_step"] = ms_pg / max(1, args.steps // 2)
            e2e["pageable_stages_ms"] = stages_pg
            del pageable
        del host

    # ---- roofline of the dominant hand-written kernel (assignment)
    peak, peak_src = measured_peaks()
    s_real = args.precision
    F = 2 if w["interlace"] else 1
    ntot = w["ng"] ** 3
    ncmplx = w["ng"] ** 2 * (w["ng"] // 2 + 1)
    b_assign = 32 * n + F * ntot * s_real
    b_bin = F * ncmplx * 2 * s_real
    b_fft = F * 3 * (ntot * s_real + ncmplx * 2 * s_real)
    t_assign = stages.get("assign", 0.0) * 1e-3
    ach = b_assign / t_assign / 1e9 if t_assign > 0 else 0.0
    # DRAM traffic of the kernel per launch from the committed ncu --set full capture
    # (profiles/r1_v4_ncu_k_assign_coop.txt: dram__bytes_read 21.01 GB + write 17.55 GB);
    # only valid for the workload it was captured on
    traffic = 38.56e9 if (args.workload == "c2" and not (args.npart or args.ng or args.opt)
                          and args.precision == 8) else None
    roofline = {"kernel": "k_assign_coop<%s,%s,%s> (mass assignment, %d field(s))" % (
                    w["assign"], "double" if args.precision == 8 else "float",
                    "interlaced" if w["interlace"] else "single", F),
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "peak_source": peak_src, "traffic": traffic,
                "limiter": "L2 atomic sector-request rate: 2.7e9 requests (18 rows x 1.5 sectors per particle) "
                           "per launch against ~190e9/s measured by tools/red_probe.cu on B200; not HBM",
                "algorithmic_bytes": b_assign, "launch_ms": stages.get("assign", 0.0),
                "other_stages": {
                    "fft": {"ms": stages.get("fft", 0.0), "algorithmic_bytes": b_fft,
                            "GBps": b_fft / max(stages.get("fft", 1e-9), 1e-9) / 1e6,
                            "frac": b_fft / max(stages.get("fft", 1e-9), 1e-9) / 1e6 / peak,
                            "note": "z pass: cuFFT batched 1-D r2c, y and x passes: k_fft_strided "
                                    "(hand-written; Ng in 512/1024/1536/2048), z + y run group by group "
                                    "over planes that fit the L2 (option fft_l2_mb); else cuFFT 3-D"},
                    "fft_x_pass": {"kernel": "k_fft_strided (x pass)", "launches": F,
                                   "ms": stages.get("fft_strided", 0.0),
                                   "algorithmic_bytes_per_launch": ncmplx * 2 * s_real * 2,
                                   "GBps": F * ncmplx * 4 * s_real / max(stages.get("fft_strided", 1e-9), 1e-9) / 1e6,
                                   "frac": F * ncmplx * 4 * s_real / max(stages.get("fft_strided", 1e-9), 1e-9) / 1e6 / peak,
                                   "note": "reads and writes every complex cell once; skips the columns "
                                           "beyond the last k edge (counted as moved here)"},
                    "bin": {"ms": stages.get("bin", 0.0), "algorithmic_bytes": b_bin,
                            "GBps": b_bin / max(stages.get("bin", 1e-9), 1e-9) / 1e6,
                            "frac": b_bin / max(stages.get("bin", 1e-9), 1e-9) / 1e6 / peak,
                            "cells_per_s": ncmplx / max(stages.get("bin", 1e-9), 1e-9) * 1e3}}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference(args.workload, 1, 1)
            cpu = {"value": r["value"], "unit": "particles/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample"], "backend": r["backend"], "s_per_step": r["s_per_step"],
                   "t_genr_mesh_s": r["t_genr_mesh_s"], "t_powspec_s": r["t_powspec_s"]}
        except Exception as ex:      # the oracle is a checker; never let it break the product's line
            cpu = {"value": None, "unit": "particles/s", "cores": os.cpu_count(), "kind": "unavailable",
                   "sample": f"oracle failed: {ex}"}

    line = {"metric": "particles_per_second_P_ell_1024_TSC_interlaced", "value": value,
            "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 8 else "f32", "data": "synthetic",
            "config": config, "stages_ms": stages, "wall_ms_per_step": wall_ms / args.steps,
            "particles_per_s_assigned": world * n / max(t_assign, 1e-12),
            "mesh_cells_per_s_binned": world * ncmplx / max(stages.get("bin", 1e-9) * 1e-3, 1e-12),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "P0_first_bins": [float(x) for x in pk.pl[0][0][:3]]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
