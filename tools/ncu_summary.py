#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): per captured launch the duration,
DRAM bytes, pipe utilisations and the top stall reasons.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--traffic kernel_regex workload precision]

--traffic appends / replaces the matching entry of profiles/ncu_traffic.json (the figure
bench.py reports as roofline.traffic)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum"]


def gb(value, unit):
    v = float(value)
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = None
    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        traffic = (re.compile(sys.argv[k + 1]), sys.argv[k + 2], int(sys.argv[k + 3]))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print("---", name[:150])
        for w in WANT:
            if w in idx and r[idx[w]]:
                print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
        st = [(h, float(r[i])) for h, i in idx.items()
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
        st.sort(key=lambda x: -x[1])
        for h, v in st[:5]:
            print("  stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                  f"{v:.2f} warps/issue")
        if traffic and traffic[0].search(name):
            rd = gb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = gb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
            try:
                ent = json.load(open(path))
            except Exception:
                ent = []
            key = traffic[0].pattern
            ent = [e for e in ent if not (e["kernel"] == key and e["workload"] == traffic[1] and e["precision"] == traffic[2])]
            ent.append({"kernel": key, "workload": traffic[1], "precision": traffic[2],
                        "traffic_bytes": (rd + wr) * 1e9, "dram_read_bytes": rd * 1e9, "dram_write_bytes": wr * 1e9,
                        "source": f"ncu --set full capture {os.path.basename(rep)} (summary under profiles/)"})
            json.dump(ent, open(path, "w"), indent=1)
            traffic = None


if __name__ == "__main__":
    main()
