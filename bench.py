#!/usr/bin/env python
"""Benchmark of the powspec hot path (BASELINE.json metric: P_ell(k) wall-time &
particles/s assigned, 1024^3 TSC interlaced).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one full pass of the replaced span, genr_mesh() + powspec(): mass
assignment of the whole catalogue (both interlaced fields), the r2c FFTs, and
the fused multipole binning, ending with the P_ell(k) arrays on the host.

N = 1
* value   : particles/s with the catalogue already resident in HBM (generated on
            the device), CUDA-event timed on the library's stream.
* e2e     : the same through the host API with the catalogue in (pinned) HOST
            memory — H2D of the 32-byte particle records inside the timed region,
            P_ell(k) read back.  `e2e_cold_ms`: the very first call of the process
            (mesh allocation, cuFFT plans, staging buffers), wall clock.
* roofline: the assignment kernel (the dominant hand-written kernel), algorithmic
            bytes 32 N + F Ntot s over its CUDA-event time, against
            the measured HBM copy bandwidth; `stage` = the same bytes over the whole
            assignment stage (sort + memset + scatter), which is what §8(d) counts.
* cpu_baseline + parity: the UNMODIFIED reference genr_mesh()+powspec() (oracle/_ref,
            OpenMP on all host cores; FFT = the repo's FFTW-API shim, FFTW itself is
            not installed) on the SAME catalogue at the FULL size when the host has
            the memory (else a scaled twin, labelled), and the GPU result compared
            with it: mode counts bit for bit, P_ell(k) relative error.
* clustered: the same step on a clustered catalogue.

N > 1 (one process per GPU, torchrun): ONE mesh, x-slab-decomposed over the N GPUs
(strong scaling; csrc/dist.cu: particle routing, halo ring, distributed FFT — peer
stores fused into the y pass, or NCCL all-to-all — and an allreduce of the bins, all
issued by the library).  The line carries the per-stage times, the bytes through
NVLink and their fraction of the link rate, a parity block against the one-GPU
result of the same catalogue, and `replicas` (N independent catalogues, one per
GPU: the no-exchange figure) as an extra key.

--impl reference: the reference's own CPU code on the host cores, same config.
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

METRIC = "particles_per_second_P_ell_1024_TSC_interlaced"
NVLINK_GBS = 770.0          # measured peer copy per direction (B200_PROFILING.md); 900 nominal

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(npart=10 ** 8, box=1000.0, ng=1024, assign="TSC", interlace=True, ncat=1,
               poles=(0, 2, 4), kbin=0.01,
               desc="periodic box, 1e8 uniform particles, 1024^3 mesh, TSC + interlacing, P_0/P_2/P_4"),
    # BASELINE.json configs[0] (C1)
    "c1": dict(npart=10 ** 6, box=1000.0, ng=256, assign="CIC", interlace=False, ncat=1,
               poles=(0, 2), kbin=0.01,
               desc="periodic box, 1e6 uniform particles, 256^3 mesh, CIC, no interlacing, P_0/P_2"),
    # BASELINE.json configs[3] (C4) and configs[4] (C5): slab-decomposed only (--gpus >= 4 / 8)
    "c4": dict(npart=10 ** 9, box=2000.0, ng=2048, assign="PCS", interlace=True, ncat=2,
               poles=(0, 2, 4), kbin=0.01,
               desc="cross power spectrum of two 1e9-particle boxes, 2048^3 mesh, PCS + interlacing"),
    "c5": dict(npart=8 * 10 ** 9, box=4000.0, ng=4096, assign="TSC", interlace=True, ncat=1,
               poles=(0, 2, 4), kbin=0.01,
               desc="periodic box, 8e9 particles, 4096^3 slab-decomposed mesh, TSC + interlacing"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload, precision):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None when there is no
    capture of this kernel on this workload."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            for e in json.load(f):
                if e["kernel"] == kernel and e["workload"] == workload and e["precision"] == precision:
                    return e
    except Exception:
        pass
    return None


def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) / 1048576.0
    except Exception:
        pass
    return 0.0


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


# ---------------------------------------------------------------------------
# the reference's CPU implementation (the checker: oracle/)
# ---------------------------------------------------------------------------
def reference_sample(workload, cores):
    """What the CPU leg runs: the FULL workload when the host has the memory (config 2:
    38.7 GB of meshes + 2 x 3.2 GB of particles), else a scaled-down
    twin (same particles per cell, scheme, interlacing, multipoles, dk)."""
    w = WORKLOADS[workload]
    need_gb = 3 * w["ng"] ** 3 * 8 * (2 if w["interlace"] else 1) / 1e9 * 0.8 + w["npart"] * 32 * 3 / 1e9 + 4
    if w["ng"] <= 256 or host_mem_available_gb() >= need_gb:
        return dict(w), "full workload"
    ng = 512 if cores >= 16 else 256
    scale = (ng / w["ng"]) ** 3
    s = dict(w)
    s["ng"] = ng
    s["npart"] = int(round(w["npart"] * scale))
    s["box"] = w["box"] * ng / w["ng"]
    return s, (f"scaled twin: {s['npart']} particles, {ng}^3 mesh, box {s['box']:g} "
               f"(same particles/cell, scheme, interlacing, multipoles, dk as the full workload; "
               f"host has {host_mem_available_gb():.0f} GB available, the full workload needs {need_gb:.0f})")


def run_reference(workload, steps, warmup, catalogue=None):
    """genr_mesh() + powspec() of the reference on all host cores.  Timed span = the two
    calls themselves (oracle/ref_driver.c brackets them; the driver's particle copy and
    result extraction are outside).  catalogue: host (N, 4) array to use instead of a
    fresh uniform one (the GPU leg's own particles, for the parity block)."""
    from oracle import have_ref, load_oracle
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    kind = "reference" if have_ref() else "port"
    orc = load_oracle("ref" if have_ref() else "port")
    s, sample_desc = reference_sample(workload, cores)
    full = sample_desc == "full workload"
    if catalogue is not None and full:
        cat = catalogue
    else:
        rng = np.random.default_rng(1)
        cat = np.empty((s["npart"], 4))
        for a in range(3):
            cat[:, a] = rng.random(s["npart"]) * s["box"]
        cat[:, 3] = 1.0
    kw = dict(ng=s["ng"], assign=s["assign"], interlace=s["interlace"], poles=s["poles"],
              box=s["box"], kbin=s["kbin"])
    tm, tp, r = [], [], None
    for it in range(warmup + steps):
        r = orc.run(cat, **kw)
        if it >= warmup:
            tm.append(r.t_mesh); tp.append(r.t_pk)
    t = float(np.mean(tm) + np.mean(tp))
    return dict(value=s["npart"] / t, unit="particles/s", cores=cores, kind=kind, sample=sample_desc,
                backend=f"{orc.backend}; FFT = the repo's FFTW-API shim (oracle/fftw_shim), not FFTW "
                        "(FFTW is not installed in this image)",
                s_per_step=t, t_genr_mesh_s=float(np.mean(tm)), t_powspec_s=float(np.mean(tp)),
                npart=s["npart"], ng=s["ng"], steps=steps, warmup=warmup, full=full, same_catalogue=
                catalogue is not None and full), r


def parity_block(got, want, vs, tol):
    """Mode counts bit for bit; P_ell(k) error per point relative to
    max(|P|, 1e-3 max|P|) of the spectrum (tests/parity.py)."""
    def arr(x):
        return None if x is None else np.asarray(x, dtype=np.float64)
    out = {"vs": vs, "tolerance": tol,
           "nmod_equal": bool(np.array_equal(np.asarray(got.cnt, dtype=np.uint64),
                                             np.asarray(want.cnt, dtype=np.uint64))),
           "nbin": int(got.nbin)}
    worst = 0.0
    for g, w in list(zip(got.pl, want.pl)) + [(got.xpl, want.xpl)]:
        g, w = arr(g), arr(w)
        if g is None or w is None:
            continue
        floor = 1e-3 * float(np.max(np.abs(w)))
        worst = max(worst, float(np.max(np.abs(g - w) / np.maximum(np.abs(w), floor))))
    out["pl_max_rel_err"] = worst
    out["kavg_max_rel_err"] = float(np.max(np.abs(np.asarray(got.km) - np.asarray(want.km)) /
                                          np.maximum(np.abs(np.asarray(want.km)), 1e-300)))
    out["ok"] = bool(out["nmod_equal"] and worst < tol)
    return out


# ---------------------------------------------------------------------------
# N > 1: one mesh over all ranks
# ---------------------------------------------------------------------------
def bench_slab(args, ctx, conf, w, world, rank, local_rank, config, barrier):
    import torch
    import torch.distributed as dist

    import powspec_b200
    from powspec_b200.api import Cata
    from powspec_b200.dist import NcclRank

    ncat = w["ncat"]
    eng = NcclRank.from_torch(ctx)
    for kv in args.opt:
        name, val = kv.split("=")
        if name == "p2p":
            eng.set_option(name, int(val))
    n_total = w["npart"] - w["npart"] % world
    n_loc = n_total // world
    chunk = min(n_loc, args.chunk)
    resident = n_loc * 32 * ncat <= 8e9             # else the share is generated chunk by chunk in the step
    lib_stream = ctx.torch_stream()

    def gen(seed, first, m, out=None):
        t = out if out is not None else torch.empty((m, 4), dtype=torch.float64, device="cuda")
        ctx.generate_into(t[:m], w["box"], kind=args.kind, seed=seed, first_index=first)
        return t[:m]

    shares = [gen(1 + c, rank * n_loc, n_loc) for c in range(ncat)] if resident else None
    scratch = None if resident else torch.empty((chunk, 4), dtype=torch.float64, device="cuda")

    def step():
        eng.begin(conf)
        for c in range(ncat):
            if resident:
                eng.add(c, shares[c])
            else:       # catalogues larger than HBM: generated on the device, chunk by chunk
                done = 0
                while done < n_loc:
                    m = min(chunk, n_loc - done)
                    eng.add(c, gen(1 + c, rank * n_loc + done, m, scratch))
                    done += m
        return eng.finish([float(n_total)] * ncat)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage, dstage, launches, pk = {}, {}, 0, None
        e0.record(lib_stream)
        for _ in range(steps):
            pk = fn()
            launches += pk.launches
            for k_, v in pk.timings_ms.items():
                stage[k_] = stage.get(k_, 0.0) + v / steps
            for k_, v in pk.stages_ms.items():
                dstage[k_] = dstage.get(k_, 0.0) + v / steps
        e1.record(lib_stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, stage, dstage, launches, pk

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, stage, dstage, launches, pk = timed(step, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    # per-stage times: the slowest rank of each stage
    keys = sorted(dstage)
    mx = torch.tensor([dstage[k] for k in keys] + [stage.get("assign", 0.0), stage.get("sort", 0.0)],
                      device="cuda", dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dmax = {k: float(v) for k, v in zip(keys, mx[:len(keys)].tolist())}
    kern_assign_ms, sort_ms = float(mx[-2]), float(mx[-1])

    # ---- end to end: every rank's share in pinned host memory, uploaded inside the step
    e2e = None
    if resident and not args.no_e2e:
        hosts = []
        for c in range(ncat):
            h = torch.empty((n_loc, 4), dtype=torch.float64, pin_memory=True)
            h.copy_(shares[c])
            hosts.append(h)
        nchunks = max(2, -(-n_loc // 6_250_000))

        def step_host():
            eng.begin(conf)
            for c in range(ncat):
                eng.add_host(c, hosts[c], nchunks)
            return eng.finish([float(n_total)] * ncat)
        ms_e2e, st_e, dst_e, _, pk_e = timed(step_host, args.steps, max(args.warmup, 3))
        e2e = {"value": n_total * ncat / (ms_e2e * 1e-3), "unit": "particles/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": n_total * 32 * ncat, "d2h_bytes_per_step":
               (2 + 4 * pk_e.nl) * pk_e.nbin * 8 * world, "host_memory": "pinned, 1/N of the catalogue per rank",
               "upload_chunks_per_rank": nchunks, "stages_ms_rank0": dst_e}
        del hosts

    # ---- parity inside the run: rank 0 transforms the SAME catalogue on one GPU
    parity = None
    if n_total * 32 * ncat <= 16e9 and w["ng"] <= 1024:
        if rank == 0:
            full = [gen(1 + c, 0, n_total) for c in range(ncat)]
            one = powspec_b200.Context(local_rank)
            m = one.genr_mesh(conf, Cata(data=full, wdata=[float(n_total)] * ncat))
            pk1 = one.powspec(conf, None, m)
            parity = parity_block(pk, pk1, "the one-GPU path on the same catalogue, inside this run", 1e-6
                                  if args.precision == 8 else 1e-4)
            one.close()
            del full
        barrier()

    # ---- replicas: N independent catalogues, one per GPU, no exchange (the no-collective figure)
    replicas = None
    if not args.no_replicas and w["npart"] * 32 * ncat <= 8e9 and w["ng"] <= 1024:
        del shares
        own = [ctx.generate_catalog(w["npart"], w["box"], kind=args.kind, seed=11 + rank + c) for c in range(ncat)]
        cata = Cata(data=own, wdata=[float(w["npart"])] * ncat)

        def step_one():
            return ctx.powspec(conf, cata, ctx.genr_mesh(conf, cata))
        for _ in range(3):
            step_one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(args.steps):
            step_one()
        e1.record(lib_stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rms = float(t.item()) / args.steps
        replicas = {"what": f"{world} independent catalogues of the full workload, one per GPU, no data-path "
                            "collective (weak scaling)", "ms_per_step": rms,
                    "value": world * w["npart"] * ncat / (rms * 1e-3), "unit": "particles/s"}
        for o in own:
            ctx.free_catalog(o)

    if rank != 0:
        eng.close()
        dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    s_real = args.precision
    F = 2 if w["interlace"] else 1
    ng = w["ng"]
    ngk = ng // 2 + 1
    ncmplx = ng * ng * ngk
    nx = ng // world
    # algorithmic bytes of one rank (per rank: N/G particles, Ntot/G cells)
    b_assign = ncat * (32 * n_loc + F * nx * ng * ng * s_real)
    b_fft = ncat * F * 3 * (ng ** 3 * s_real + ncmplx * 2 * s_real) / world
    b_bin = ncat * F * ncmplx * 2 * s_real / world
    sent = pk.traffic["transpose_bytes_sent"]
    peer = pk.traffic["peer_stores"]
    t_link = dmax.get("fft_zy", 0.0) if peer else dmax.get("transpose", 0.0)

    def frac(b, ms):
        return {"ms": ms, "algorithmic_bytes_per_rank": b, "GBps": b / max(ms, 1e-9) / 1e6,
                "frac": b / max(ms, 1e-9) / 1e6 / peak}
    roofline = {"kernel": "k_assign_coop (mass assignment of this rank's slab, %d field(s))" % F,
                "bound": "hbm", "achieved": b_assign / max(kern_assign_ms, 1e-9) / 1e6, "peak": peak,
                "unit": "GB/s", "frac": b_assign / max(kern_assign_ms, 1e-9) / 1e6 / peak,
                "peak_source": peak_src, "traffic": None, "algorithmic_bytes": b_assign,
                "launch_ms": kern_assign_ms,
                "stages_per_rank": {"assign_stage(route+sort+scatter)": frac(b_assign, dmax.get("route", 0) + dmax.get("assign", 0)),
                                    "fft(zy+x)": frac(b_fft, dmax.get("fft_zy", 0) + dmax.get("fft_x", 0)),
                                    "bin": frac(b_bin, dmax.get("bin", 0))}}
    nvlink = {"bound": "nvlink", "what": "distributed-FFT transposes, bytes SENT per rank and step: "
              "(G-1)/G * Ncmplx * 2s / G per field (SURVEY.md §8d)",
              "transport": "peer stores fused into the y pass (no all-to-all)" if peer else "NCCL all-to-all "
              "(second stream, overlapped with the next field's z/y passes)",
              "bytes_per_rank_per_step": sent, "route_bytes_per_rank_per_step": pk.traffic["route_bytes_sent"],
              "ms": t_link, "ms_is": "the y-pass kernels that carry the stores (FFT work included: a lower bound "
              "on the link rate)" if peer else "the all-to-all intervals on the communication stream",
              "achieved": sent / max(t_link, 1e-9) / 1e6, "peak": NVLINK_GBS, "peak_nominal": 900.0,
              "unit": "GB/s", "frac": sent / max(t_link, 1e-9) / 1e6 / NVLINK_GBS,
              "peak_source": "measured peer copy per direction, B200_PROFILING.md"}
    config = dict(config)
    config["parallelism"] = (f"ONE {ng}^3 mesh x-slab-decomposed over {world} GPUs: {nx} x-planes per rank, particles "
                             f"routed to the owner of their base cell, halo ring, distributed FFT, allreduce of the bins")
    config["npart_total"] = n_total
    config["npart_per_gpu"] = n_loc
    config["transport"] = "library-issued NCCL (dlopen) + " + ("CUDA-IPC peer stores" if peer else "all-to-all")
    line = {"metric": METRIC, "value": n_total * ncat / (ms_step * 1e-3), "unit": "particles/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.precision == 8 else "f32", "data": "synthetic", "config": config,
            "mode": "slab", "stages_ms_max_over_ranks": dmax, "kernel_ms_max_over_ranks":
            {"k_assign_coop": kern_assign_ms, "row_sort": sort_ms},
            "particles_per_s_assigned": n_total * ncat / max((dmax.get("route", 0) + dmax.get("assign", 0)) * 1e-3, 1e-12),
            "mesh_cells_per_s_binned": ncat * ncmplx / max(dmax.get("bin", 1e-9) * 1e-3, 1e-12),
            "roofline": roofline, "nvlink": nvlink, "parity": parity, "replicas": replicas,
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "P0_first_bins": [float(x) for x in pk.pl[0][0][:3]]}
    print(json.dumps(line))
    eng.close()
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
    ap.add_argument("--chunk", type=int, default=1 << 26, help="slab mode: particles per generated chunk and rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-clustered", action="store_true")
    ap.add_argument("--no-replicas", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="context option name=value (ablations)")
    ap.add_argument("--mode", default=None, choices=["replica", "slab"],
                    help="N>1: 'slab' (default) = ONE mesh x-slab-decomposed over the GPUs (strong scaling); "
                         "'replica' = one independent catalogue per GPU (weak scaling, no collective)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    mode = args.mode or ("slab" if world > 1 else "replica")
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
        # a full-size step takes ~20 s on 16 cores: cap the counts, and say so
        full_steps = min(max(args.steps, 1), 2) if WORKLOADS[args.workload]["ng"] > 256 else max(args.steps, 1)
        r, _ = run_reference(args.workload, full_steps, min(max(args.warmup, 0), 1))
        config["parallelism"] = f"host CPU, {r['cores']} OpenMP threads"
        config["reference_ran"] = r["sample"]
        line = {"impl": "reference", "metric": METRIC,
                "value": r["value"], "unit": "particles/s", "n_gpus": args.gpus, "steps": r["steps"],
                "warmup": r["warmup"], "steps_requested": args.steps, "warmup_requested": args.warmup,
                "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "particles/s", "cores": r["cores"],
                                 "kind": r["kind"], "sample": r["sample"], "backend": r["backend"],
                                 "timed_span": "genr_mesh() + powspec() only (oracle/ref_driver.c)",
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
        if name != "p2p":
            ctx.set_option(name, int(val))
    if args.opt:
        config["options"] = args.opt
    n = w["npart"]
    ncat = w["ncat"]
    conf = Conf(ndata=ncat, issim=True, bsize=(w["box"],) * 3, gsize=w["ng"],
                assign=powspec_b200.powspec_assign_names.index(w["assign"]), intlace=w["interlace"],
                poles=tuple(w["poles"]), kbin=w["kbin"], isauto=(True, ncat == 2), iscross=ncat == 2,
                precision=args.precision, device=local_rank)

    if mode == "slab" and world > 1:
        return bench_slab(args, ctx, conf, w, world, rank, local_rank, config, barrier)
    if args.workload in ("c4", "c5"):
        raise SystemExit("bench.py: workloads c4 / c5 are slab-decomposed: launch with torchrun on >= 4 / 8 GPUs")

    lib_stream = ctx.torch_stream()
    cat_dev = ctx.generate_catalog(n, w["box"], kind=args.kind, seed=1 + rank)

    def step(cata):
        mesh = ctx.genr_mesh(conf, cata)
        return ctx.powspec(conf, cata, mesh)

    # ---- the very first call of the process, from host memory: allocation of the meshes
    # (17 GB), cuFFT plans, chunk and staging buffers — what a one-shot caller (the
    # reference's C host calls genr_mesh / powspec once) pays; wall clock
    host = None
    e2e_cold_ms = None
    if not args.no_e2e:
        host = torch.empty((n, 4), dtype=torch.float64, pin_memory=True)
        ctx.L.psb_copy_to_host(ctx.h, host.data_ptr(), cat_dev[0], n * 32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(Cata(data=[host], wdata=[float(n)]))
        e2e_cold_ms = (time.perf_counter() - t0) * 1e3

    def timed(cata, steps, warmup):
        for _ in range(warmup):
            step(cata)
        barrier()
        stage = {}
        launches = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        t0 = time.perf_counter()
        pk = None
        for _ in range(steps):
            pk = step(cata)
            launches += pk.launches
            for k_, v in pk.timings_ms.items():
                stage[k_] = stage.get(k_, 0.0) + v
        e1.record(lib_stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall * 1e3, {k_: v / steps for k_, v in stage.items()}, launches, pk

    cata_dev = Cata(data=[cat_dev], wdata=[float(n)])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, wall_ms, stages, launches, pk = timed(cata_dev, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- end to end: catalogue in pinned host memory, H2D inside the timed region
    e2e = None
    if host is not None:
        cata_host = Cata(data=[host], wdata=[float(n)])
        ms_e2e, _, stages_e2e, _, pk_e = timed(cata_host, args.steps, max(args.warmup, 3))
        ms_e2e_step = ms_e2e / args.steps
        d2h = (2 + 4 * pk_e.nl) * pk_e.nbin * 8 + 6 * 8 * 148 * 8
        e2e = {"value": world * n / (ms_e2e_step * 1e-3), "unit": "particles/s",
               "ms_per_step": ms_e2e_step, "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": d2h,
               "host_memory": "pinned", "stages_ms": stages_e2e, "e2e_cold_ms": e2e_cold_ms,
               "e2e_cold_is": "first call of the process from pinned host memory: mesh allocation, cuFFT plans, "
                              "chunk buffers (wall clock)"}
        # the same from PAGEABLE host memory (what the reference's C host hands over:
        # plain malloc), staged through pinned buffers by the library
        if world == 1:
            pageable = host.numpy().copy()
            cata_pg = Cata(data=[pageable], wdata=[float(n)])
            ms_pg, _, stages_pg, _, _ = timed(cata_pg, max(1, args.steps // 2), 1)
            e2e["pageable_ms_per_step"] = ms_pg / max(1, args.steps // 2)
            e2e["pageable_stages_ms"] = stages_pg
            del pageable

    # ---- the same step on a clustered catalogue (blobs of sigma = 2 cells
    # + 20 % uniform background)
    F = 2 if w["interlace"] else 1
    ntot = w["ng"] ** 3
    ncmplx = w["ng"] ** 2 * (w["ng"] // 2 + 1)
    clustered = None
    if world == 1 and args.kind == 0 and not args.no_clustered:
        cl = ctx.generate_catalog(n, w["box"], kind=1, seed=1)
        ms_c, _, st_c, _, _ = timed(Cata(data=[cl], wdata=[float(n)]), min(args.steps, 5), 3)
        ms_c /= min(args.steps, 5)
        ctx.free_catalog(cl)
        clustered = {"catalogue": "clustered: N/1000 blobs of sigma = 2 Mpc/h + 20 % uniform background",
                     "ms_per_step": ms_c, "value": n / (ms_c * 1e-3), "unit": "particles/s", "stages_ms": st_c,
                     "particles_per_s_assigned": n / max(st_c.get("assign", 1e-9) * 1e-3, 1e-12),
                     "mesh_cells_per_s_binned": ncmplx / max(st_c.get("bin", 1e-9) * 1e-3, 1e-12)}

    # ---- roofline of the dominant hand-written kernel (assignment)
    peak, peak_src = measured_peaks()
    s_real = args.precision
    b_assign = 32 * n + F * ntot * s_real
    b_bin = F * ncmplx * 2 * s_real
    b_fft = F * 3 * (ntot * s_real + ncmplx * 2 * s_real)
    t_assign = stages.get("assign", 0.0) * 1e-3
    t_stage = (stages.get("assign", 0.0) + stages.get("sort", 0.0) + stages.get("memset", 0.0) +
               stages.get("bounds", 0.0)) * 1e-3
    ach = b_assign / t_assign / 1e9 if t_assign > 0 else 0.0
    tiles = ctx.L.psb_assign_path(ctx.h) == 1
    kname = "k_tile_accumulate" if tiles else "k_assign_coop"
    tr = None if (args.npart or args.ng or args.opt) else ncu_traffic(kname, args.workload, args.precision)
    if tiles:
        limiter = ("shared-memory atomic wavefronts: 54 native u32 ATOMS.ADD per particle and field (two limbs of "
                   "27 cells), ~3.5 bank wavefronts per warp instruction on random cells (ncu: l1tex 66 %, issue "
                   "48 %); DRAM traffic = lists read once + mesh written once; not HBM")
        what = ("the assignment STAGE as SURVEY.md §8(d) counts it: tile-list count + scan + fill "
                "(k_tile_lists) and the owner-computes accumulation (k_tile_accumulate; no memset, no "
                "read-modify-write) over the same algorithmic bytes")
    else:
        limiter = ("L2 atomic sector-request rate: 2.7e9 requests (18 rows x 1.5 sectors per particle) "
                   "per launch against ~190e9/s measured by tools/red_probe.cu on B200; not HBM")
        what = ("the assignment STAGE as SURVEY.md §8(d) counts it: particle sort + mesh "
                "memsets + scatter over the same algorithmic bytes")
    roofline = {"kernel": "%s<%s,%s,%s> (mass assignment, %d field(s))" % (
                    kname, w["assign"], "double" if args.precision == 8 else "float",
                    "interlaced" if w["interlace"] else "single", F),
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "peak_source": peak_src,
                "traffic": tr["traffic_bytes"] if tr else None,
                "traffic_source": tr["source"] if tr else None,
                "limiter": limiter,
                "algorithmic_bytes": b_assign, "launch_ms": stages.get("assign", 0.0),
                "stage": {"what": what,
                          "ms": t_stage * 1e3, "achieved": b_assign / max(t_stage, 1e-12) / 1e9,
                          "frac": b_assign / max(t_stage, 1e-12) / 1e9 / peak},
                "other_stages": {
                    "fft": {"ms": stages.get("fft", 0.0), "algorithmic_bytes": b_fft,
                            "GBps": b_fft / max(stages.get("fft", 1e-9), 1e-9) / 1e6,
                            "frac": b_fft / max(stages.get("fft", 1e-9), 1e-9) / 1e6 / peak,
                            "note": "z pass: cuFFT batched 1-D r2c, y and x passes: k_fft_strided "
                                    "(hand-written; Ng in 512/1024/1536/2048); else cuFFT 3-D"},
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

    # ---- the reference on the host cores, on the SAME particles, and the parity of the
    # GPU result against it (full size when the host has the memory)
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        try:
            same = host.numpy() if host is not None else ctx.catalog_to_host(cat_dev)
            r, ref = run_reference(args.workload, 1, 0, catalogue=same)
            cpu = {"value": r["value"], "unit": "particles/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample"] + ", 1 run, no warm-up", "backend": r["backend"],
                   "s_per_step": r["s_per_step"], "t_genr_mesh_s": r["t_genr_mesh_s"],
                   "t_powspec_s": r["t_powspec_s"],
                   "timed_span": "genr_mesh() + powspec() only (oracle/ref_driver.c)"}
            if r["same_catalogue"] and args.precision == 8:
                parity = parity_block(pk, ref, f"the unmodified reference ({r['kind']}) on the same {n} particles, "
                                      f"{w['ng']}^3, on {r['cores']} host cores, inside this run", 1e-6)
        except Exception as ex:      # the oracle is a checker; never let it break the product's line
            cpu = {"value": None, "unit": "particles/s", "cores": os.cpu_count(), "kind": "unavailable",
                   "sample": f"oracle failed: {ex}"}

    line = {"metric": METRIC, "value": value,
            "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 8 else "f32", "data": "synthetic",
            "config": config, "stages_ms": stages, "wall_ms_per_step": wall_ms / args.steps,
            "particles_per_s_assigned": world * n / max(t_assign, 1e-12),
            "particles_per_s_assigned_stage": world * n / max(t_stage, 1e-12),
            "mesh_cells_per_s_binned": world * ncmplx / max(stages.get("bin", 1e-9) * 1e-3, 1e-12),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "clustered": clustered,
            "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "P0_first_bins": [float(x) for x in pk.pl[0][0][:3]]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
