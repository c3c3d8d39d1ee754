#!/usr/bin/env python
"""Per-pass timing of the hand-written FFT kernels against cuFFT (through torch.fft) on
full-size meshes: the strided y and x passes, the r2c z pass and the fused z + y kernel,
for every supported size and both precisions.  Prints one JSON line per (size, precision)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=3):
    import torch
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    import torch

    import powspec_b200 as pb
    ctx = pb.Context(0)
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 1536, 2048]
    for ng in sizes:
        ngk = ng // 2 + 1
        for prec in (8, 4):
            cdt = torch.complex128 if prec == 8 else torch.complex64
            rdt = torch.float64 if prec == 8 else torch.float32
            x = torch.zeros(ng, ng, ngk, dtype=cdt, device="cuda")
            xr = torch.view_as_real(x)
            xr.normal_()
            gb = 2 * x.numel() * x.element_size() / 1e9          # read + write of one pass
            out = {"ng": ng, "precision": prec, "pass_GB": gb}
            out["own_y_ms"] = timed(lambda: ctx.fft_axis(x, 1))
            out["own_x_ms"] = timed(lambda: ctx.fft_axis(x, 0))
            real = xr.view(ng * ng, 2 * ngk)
            out["own_z_r2c_ms"] = timed(lambda: ctx.fft_rows(real, ng))
            out["own_zy_fused_ms"] = timed(lambda: ctx.fft_zy(xr.view(ng, ng, 2 * ngk)))
            try:
                # cuFFT, out of place (torch allocates the result): same bytes moved
                out["cufft_y_ms"] = timed(lambda: torch.fft.fft(x, dim=1))
                out["cufft_x_ms"] = timed(lambda: torch.fft.fft(x, dim=0))
                rr = real.view(ng, ng, 2 * ngk)[:, :, :ng]
                out["cufft_z_r2c_ms"] = timed(lambda: torch.fft.rfft(rr, dim=2))
            except torch.OutOfMemoryError:
                out["cufft"] = "out of memory (out-of-place result + work area)"
            for k in list(out):
                if k.endswith("_ms") and "zy" not in k:
                    out[k.replace("_ms", "_TBps")] = round(gb / out[k], 3)
            print(json.dumps(out), flush=True)
            del x, xr, real
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
