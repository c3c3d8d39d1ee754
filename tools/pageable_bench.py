#!/usr/bin/env python
"""End-to-end step of BASELINE config 2 from PAGEABLE host memory (what the reference's C
host hands genr_mesh(): plain malloc'd DATA arrays, src/read_cata.c:86-189) for the variants
of the staging copy (csrc/hostcopy.cpp): streaming stores on / off, 8 / 16 staging threads,
beside the same step from pinned memory.  One JSON line per variant.

    python tools/pageable_bench.py [npart] [steps]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import powspec_b200 as pb
    from powspec_b200.api import Cata, Conf
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ctx = pb.Context(0)
    conf = Conf(ndata=1, issim=True, bsize=(1000.0,) * 3, gsize=1024, assign=2, intlace=True,
                poles=(0, 2, 4), kbin=0.01, isauto=(True, False), iscross=False, precision=8, device=0)
    dev = ctx.generate_catalog(n, 1000.0, kind=0, seed=1)
    pinned = torch.empty((n, 4), dtype=torch.float64, pin_memory=True)
    ctx.L.psb_copy_to_host(ctx.h, pinned.data_ptr(), dev[0], n * 32)
    torch.cuda.synchronize()
    ctx.free_catalog(dev)
    pageable = pinned.numpy().copy()

    def run(data, label, **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        cata = Cata(data=[data], wdata=[float(n)])
        ts, pk = [], None
        for it in range(2 + steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pk = ctx.powspec(conf, cata, ctx.genr_mesh(conf, cata))
            torch.cuda.synchronize()
            if it >= 2:
                ts.append((time.perf_counter() - t0) * 1e3)
        print(json.dumps({"source": label, "options": opts, "npart": n, "ms_per_step_wall": float(np.mean(ts)),
                          "ms_min": float(np.min(ts)), "h2d_stage_ms": pk.timings_ms.get("h2d"),
                          "GBps_over_h2d_stage": n * 32 / max(pk.timings_ms.get("h2d", 0.0), 1e-9) / 1e6,
                          "P0_first": float(pk.pl[0][0][0])}), flush=True)

    run(pinned, "pinned")
    # (streaming stores, MiB per piece, pieces, threads); twice, interleaved: the host's cores and
    # last-level cache are shared with other tenants of the box, single runs scatter by +-5 ms
    for nt, piece, slots, thr in 2 * ((0, 8, 4, 16), (0, 16, 3, 16), (1, 64, 2, 16), (0, 12, 4, 16)):
        run(pageable, "pageable", h2d_nt=nt, h2d_piece_mb=piece, h2d_slots=slots, h2d_threads=thr)
    run(pinned, "pinned")


if __name__ == "__main__":
    main()
