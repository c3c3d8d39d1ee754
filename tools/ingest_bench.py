#!/usr/bin/env python
"""Throughput of the binary catalogue ingest (psb_catalog_load) on a 1e8-row, 4-column
float64 .npy file (3.2 GB, BASELINE config 2's catalogue) from the page cache, beside
the reference's ASCII reader (POWSPEC_ref, a bounded 2e6-line sample, all host cores)."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import powspec_b200 as pb
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 8
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(tmp, "cat.npy")
    r = np.random.default_rng(1)
    a = np.lib.format.open_memmap(path, mode="w+", dtype=np.float64, shape=(n, 4))
    for i in range(0, n, 1 << 24):
        m = min(1 << 24, n - i)
        a[i:i + m, :3] = r.random((m, 3)) * 1000.0
        a[i:i + m, 3] = 1.0
    a.flush(); del a
    ctx = pb.Context(0)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        cat, sums = ctx.load_catalog(path, pos=(0, 1, 2), wcomp=3, issim=True)
        best = min(best, time.perf_counter() - t0)
        ctx.free_catalog(cat)
    out = {"rows": n, "file_GB": n * 32 / 1e9, "s": best, "rows_per_s": n / best,
           "GBps": n * 32 / best / 1e9, "sumw": sums["sumw"]}
    ref = os.path.join(ROOT, "oracle", "_ref", "POWSPEC_ref")
    if os.path.exists(ref):
        m = 2_000_000
        np.savetxt(os.path.join(tmp, "cat.txt"), np.c_[r.random((m, 3)) * 1000.0, np.ones(m)], fmt="%.10g")
        conf = os.path.join(tmp, "c.conf")
        open(conf, "w").write("""DATA_CATALOG = cat.txt
DATA_FORMATTER = "%lf %lf %lf %lf"
DATA_POSITION = [$1,$2,$3]
DATA_WT_COMP = $4
CUBIC_SIM = T
BOX_SIZE = 1000
GRID_SIZE = 16
PARTICLE_ASSIGN = 0
GRID_INTERLACE = F
MULTIPOLE = [0]
KMIN = 0
BIN_SIZE = 0.05
OUTPUT_AUTO = out.txt
OVERWRITE = 1
VERBOSE = F
""")
        t0 = time.perf_counter()
        subprocess.run([ref, "-c", conf], cwd=tmp, capture_output=True)
        t = time.perf_counter() - t0
        out["reference_ascii"] = {"rows": m, "s_whole_program_16cube": t, "rows_per_s": m / t,
                                  "cores": os.cpu_count()}
    print(json.dumps(out))
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
