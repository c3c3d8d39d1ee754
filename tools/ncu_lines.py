#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel of an ncu report
(captured with --import-source on, code built with -lineinfo).  Run here, no GPU needed.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep kernel_regex [top_n] [samples]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    by = 1 if (len(sys.argv) > 4 and sys.argv[4] == "samples") else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    per = defaultdict(lambda: [0.0, 0.0])
    fname, first, active = "", None, True
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Function Name":
            first = first or r[1]
            active = r[1] == first
            continue
        if not active:
            continue
        if r and r[0] == "Line No" and "Instructions Executed" in r:
            hdr = {h: i for i, h in enumerate(r)}
            continue
        # cuda,sass view: a row per source line (aggregated metrics) followed by its SASS rows
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        try:
            inst = float(r[hdr["Instructions Executed"]] or 0)
            smp = float(r[hdr["# Samples"]] or 0)
        except ValueError:
            continue
        key = (fname + ":" + r[0], r[1].strip()[:100])
        per[key][0] += inst
        per[key][1] += smp
    tot_i = sum(v[0] for v in per.values()) or 1
    tot_s = sum(v[1] for v in per.values()) or 1
    print(f"total warp instructions {tot_i:.4g}, samples {tot_s:.0f}")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][by])[:top]:
        print(f"{100 * v[0] / tot_i:5.1f}% inst {100 * v[1] / tot_s:5.1f}% smp  {k[0]:>24} {k[1]}")


if __name__ == "__main__":
    main()
