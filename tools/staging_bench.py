#!/usr/bin/env python
"""Rate of the pageable -> staging-buffer copy of csrc/hostcopy.cpp on this host (no GPU
work: the library's measurement hook = h2d_async without the DMA): 3.2 GB of particle
records streamed through two 64 MB buffers by one persistent pool, 4 / 8 / 16 threads, with
streaming stores and with plain memcpy in the pool's threads, beside a single-threaded numpy
copy.  One JSON line.

    python tools/staging_bench.py [GB]
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import powspec_b200 as pb
    lib = pb.load_library()
    lib.psb_test_host_stage.restype = C.c_double
    lib.psb_test_host_stage.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
    total = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 3_200_000_000
    piece = 64 << 20
    src = np.ones(total, dtype=np.uint8)
    out = {"bytes": total, "piece_bytes": piece, "host_cores": os.cpu_count()}
    for nt in (1, 0):
        for nthr in (4, 8, 16):
            best = min(lib.psb_test_host_stage(src.ctypes.data, total, piece, nthr, nt) for _ in range(3))
            out[f"{'stream' if nt else 'memcpy'}_{nthr}thr_GBps"] = total / best / 1e9
    stage = np.zeros(piece, dtype=np.uint8)
    t0 = time.perf_counter()
    for off in range(0, total, piece):
        n = min(piece, total - off)
        stage[:n] = src[off:off + n]
    out["numpy_1thr_GBps"] = total / (time.perf_counter() - t0) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
