#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few per-launch numbers DESIGN.md and bench.py quote.

usage: python profiles/summarize_ncu.py gpurun_out/prof_X.ncu-rep [--traffic profiles/traffic.json]
Prints one block per captured launch; with --traffic, writes dram bytes (read + write) per kernel into the JSON bench.py reads.
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    traffic = {}
    for r in body:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "")
        print("### %s   grid %s block %s" % (short, r[col["Grid Size"]], r[col["Block Size"]]))
        vals = {}
        for k in KEYS:
            hit = [h for h in head if h.endswith(k)]
            if not hit:
                continue
            i = col[hit[0]]
            vals[k] = (r[i], units[i])
            print("  %-88s %s %s" % (k, r[i], units[i]))
        try:
            rd = float(vals["dram__bytes_read.sum"][0]) * TO_BYTES[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * TO_BYTES[vals["dram__bytes_write.sum"][1]]
            key = short.split("<")[0].split("::")[-1]
            traffic.setdefault(key, []).append(rd + wr)
        except Exception:
            pass
    if "--traffic" in sys.argv:
        out = sys.argv[sys.argv.index("--traffic") + 1]
        try:
            cur = json.load(open(out))
        except Exception:
            cur = {}
        for k, v in traffic.items():
            cur[k] = sum(v) / len(v)
        json.dump(cur, open(out, "w"), indent=1)
        print("traffic ->", out, cur)


if __name__ == "__main__":
    main()
