#!/usr/bin/env python
"""Capability / timing runs of the BASELINE.json configurations.

    python tools/run_config.py --config c3 [--scale 1.0]
    torchrun --nproc-per-node 8 tools/run_config.py --config c5 --precision 4

c1  periodic box, 1e6 particles, 256^3, CIC, P0/P2                       (1 GPU)
c2  periodic box, 1e8 particles, 1024^3, TSC+interlace, P0/P2/P4         (1 GPU, or slab over N)
c3  survey data 2e7 + randoms 1e8, FKP weights, 1536^3, PCS, P0/P2/P4    (1 GPU)
c4  cross spectrum of two 1e9-particle boxes, 2048^3, PCS+interlace      (slab over N GPUs)
c5  periodic box, 8e9 particles, 4096^3, TSC+interlace                   (slab over 8 GPUs)

--scale s shrinks the mesh side by s and the particle numbers by s^3 (same
particles per cell).  Large catalogues are generated on the device chunk by
chunk (counter-based Philox, psb_generate_into), never held as a whole.

Checks that do not need an oracle run at these sizes: mode counts bit-exact
against the streaming CPU restatement (oracle_mode_counts; skipped above 2048^3
unless --check-counts), and P_0 of a Poisson catalogue consistent with zero
(|P_0| against the expected scatter shot * sqrt(2 / nmodes))."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def streaming_counts(ng, box, kbin, kmin=0.0):
    import ctypes as C

    from oracle import load_oracle
    lib = load_oracle("port").lib
    lib.oracle_mode_counts.restype = C.c_int
    lib.oracle_mode_counts.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double,
                                       C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    cnt = np.zeros(8192, dtype=np.uint64)
    km = np.zeros(8192)
    bs = (C.c_double * 3)(box, box, box)
    nb = lib.oracle_mode_counts(ng, bs, kmin, -1.0, kbin, 0, 8192, cnt.ctypes.data, km.ctypes.data)
    return cnt[:nb], km[:nb]


def poisson_check(pk, nsigma=6.0):
    """P_0 of a uniform random catalogue: consistent with zero.  Only bins below
    half the Nyquist frequency are used (above it the residual aliasing of the
    window-corrected estimator is not Gaussian-small)."""
    shot = pk.shot[0]
    p0 = pk.pl[0][0]
    sigma = shot * np.sqrt(2.0 / np.maximum(pk.cnt.astype(float), 1.0))
    z = (np.abs(p0) / sigma)[1:pk.nbin // 2]
    return float(np.max(z)), bool(np.all(z < nsigma))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--precision", type=int, default=8, choices=[4, 8])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--chunk", type=int, default=1 << 27, help="particles per generated chunk and rank")
    ap.add_argument("--check-counts", action="store_true")
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="context option name=value (ablations)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import powspec_b200 as pb
    from powspec_b200.distributed import GpuSlabEngine, TorchComm, slab_power

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pb.Context(local)
    for kv in args.opt:
        name, val = kv.split("=")
        ctx.set_option(name, int(val))
    s = args.scale
    out = {"config": args.config, "scale": s, "precision": args.precision, "n_gpus": world}

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    if args.config in ("c1", "c2"):
        ng, n, box = (256, 10 ** 6, 1000.0) if args.config == "c1" else (1024, 10 ** 8, 1000.0)
        ng = int(round(ng * s)); n = int(round(n * s ** 3)); box *= s
        kw = dict(ng=ng, assign="CIC" if args.config == "c1" else "TSC", interlace=args.config == "c2",
                  poles=(0, 2) if args.config == "c1" else (0, 2, 4), box=box, kbin=0.01,
                  precision=args.precision)
        cat = ctx.generate_catalog(n, box, kind=args.kind, seed=1)
        ts = []
        for _ in range(args.steps + 1):
            t0 = time.perf_counter()
            pk = pb.run(cat, ctx=ctx, wdata=[float(n)], **kw)
            ts.append(time.perf_counter() - t0)
        out.update(ng=ng, npart=n, s_per_run=min(ts[1:]), timings_ms=pk.timings_ms)
    elif args.config == "c3":
        ng = int(round(1536 * s))
        nd, nr = int(round(2e7 * s ** 3)), int(round(1e8 * s ** 3))
        from oracle.oracle import survey_scalars

        def wedge(seed, n):
            r = np.random.default_rng(seed)
            ra = np.deg2rad(r.uniform(100, 260, n)); dec = np.deg2rad(r.uniform(-10, 70, n))
            # comoving distance range of z in [0.4, 1.1] for Omega_m = 0.31 (Mpc/h), uniform in volume
            d = np.cbrt(r.uniform(1065.0 ** 3, 2560.0 ** 3, n)) * s
            cat = np.empty((n, 4))
            cat[:, 0] = d * np.cos(dec) * np.cos(ra); cat[:, 1] = d * np.cos(dec) * np.sin(ra)
            cat[:, 2] = d * np.sin(dec)
            nz = np.full(n, 3e-4)
            wfkp = 1 / (1 + 1e4 * nz)
            cat[:, 3] = wfkp
            return cat, np.ones(n), wfkp, nz
        D, dwc, dwf, dnz = wedge(1, nd)
        R, rwc, rwf, rnz = wedge(2, nr)
        sc = survey_scalars(dwc, dwf, dnz, rwc, rwf, rnz)
        kw = dict(ng=ng, assign="PCS", interlace=False, poles=(0, 2, 4), issim=False, kbin=0.005,
                  precision=args.precision, rand=[R], scalars=[sc])
        ts = []
        for _ in range(args.steps + 1):
            t0 = time.perf_counter()
            pk = pb.run(D, ctx=ctx, **kw)
            ts.append(time.perf_counter() - t0)
        out.update(ng=ng, ndata=nd, nrand=nr, s_per_run=min(ts[1:]), timings_ms=pk.timings_ms,
                   P0_first=[float(x) for x in pk.pl[0][0][:4]], nbin=pk.nbin)
        box = None
    else:
        ng, n_each, box, ncat, assign = (2048, 10 ** 9, 2000.0, 2, 3) if args.config == "c4" else \
            (4096, 8 * 10 ** 9, 4000.0, 1, 2)
        ng = int(round(ng * s)); n_each = int(round(n_each * s ** 3)); box *= s
        ng -= ng % world
        conf = pb.Conf(ndata=ncat, issim=True, bsize=(box,) * 3, gsize=ng, assign=assign, intlace=True,
                       poles=(0, 2, 4), kbin=0.01, isauto=(True, ncat == 2), iscross=ncat == 2,
                       precision=args.precision, device=local)
        eng = GpuSlabEngine(ctx, conf, world, rank)

        class NoComm:
            size, rank = 1, 0
        comm = TorchComm() if world > 1 else NoComm()
        n_loc = n_each // world

        def chunks(seed):
            """this rank's share of catalogue `seed`, generated chunk by chunk"""
            done = 0
            while done < n_loc:
                m = min(args.chunk, n_loc - done)
                t = torch.empty((m, 4), dtype=torch.float64, device="cuda")
                ctx.generate_into(t, box, kind=args.kind, seed=seed, first_index=rank * n_loc + done)
                yield t
                done += m

        class Lazy:      # re-iterable per step
            def __init__(self, seed): self.seed = seed
            def __iter__(self): return chunks(self.seed)
        cats = [Lazy(1 + c) for c in range(ncat)]
        ts = []
        for _ in range(args.steps + 1):
            sync(); t0 = time.perf_counter()
            pk = slab_power(eng, comm, cats, [float(n_loc * world)] * ncat)
            sync(); ts.append(time.perf_counter() - t0)
        out.update(ng=ng, npart_each=n_loc * world, ncat=ncat, s_per_run=min(ts[1:]),
                   peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)

    if rank == 0:
        if args.config != "c3":
            zmax, ok = poisson_check(pk)
            out.update(poisson_max_sigma=zmax, poisson_ok=ok)
            if out["ng"] <= 2048 or args.check_counts:
                t0 = time.perf_counter()
                cnt, km = streaming_counts(out["ng"], box, 0.01)
                out.update(counts_bit_exact=bool(np.array_equal(cnt, pk.cnt)),
                           kavg_max_rel=float(np.max(np.abs(km - pk.km) / np.maximum(km, 1e-300))),
                           oracle_counts_s=time.perf_counter() - t0)
        out["nbin"] = int(pk.nbin)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
