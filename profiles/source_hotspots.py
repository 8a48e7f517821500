#!/usr/bin/env python
"""Per-CUDA-source-line hot spots of one kernel in an .ncu-rep captured with --set full --import-source on.

usage: python profiles/source_hotspots.py REPORT.ncu-rep KERNEL_REGEX [TOP_N]
Prints, for the TOP_N lines by warp-stall samples: share of samples, share of warp instructions, shared-memory wavefronts
(total / excessive = bank conflicts), the three leading stall reasons, file:line and the source text."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname, head, col, lines, first_fn = "", None, None, [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if len(r) >= 2 and r[0] == "Function Name":
            if first_fn is None:
                first_fn = r[1]
            cur_fn = r[1]; continue
        if len(r) > 4 and r[0] == "Line No":
            head = r
            col = {}
            for i, h in enumerate(head):
                col.setdefault(h, i)
            continue
        if head is None or len(r) != len(head) or r[0] == "" or cur_fn != first_fn:
            continue
        lines.append((fname, r))

    def f(r, k):
        try:
            return float(r[col[k]])
        except Exception:
            return 0.0
    tot_s = sum(f(r, "# Samples") for _, r in lines) or 1.0
    tot_i = sum(f(r, "Instructions Executed") for _, r in lines) or 1.0
    tot_w = sum(f(r, "L1 Wavefronts Shared") for _, r in lines); tot_x = sum(f(r, "L1 Wavefronts Shared Excessive") for _, r in lines)
    print("# %s" % first_fn[:120])
    print("# total: %d stall samples, %d warp instructions, %d shared wavefronts of which %d excessive (bank conflicts)" % (tot_s, tot_i, tot_w, tot_x))
    stall_keys = [h for h in head if h.startswith("stall_") and "Not Issued" not in h]
    agg = {k: sum(f(r, k) for _, r in lines) for k in stall_keys}
    print("# stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    by = "Instructions Executed" if "--by-inst" in sys.argv else "# Samples"
    lines.sort(key=lambda fr: -f(fr[1], by))
    for fn, r in lines[:topn]:
        top = sorted(((k, f(r, k)) for k in stall_keys), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% smp %5.1f%% inst  shw %8d exc %8d  %-42s %s:%s  %s" % (
            100 * f(r, "# Samples") / tot_s, 100 * f(r, "Instructions Executed") / tot_i, f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Excessive"),
            " ".join("%s:%d" % (k[6:], v) for k, v in top if v > 0), fn, r[0], r[1].strip()[:100]))


if __name__ == "__main__":
    main()
