#!/usr/bin/env python
"""Summaries of an .ncu-rep for profiles/: (1) selected raw metrics of one kernel launch as metric,unit,value CSV (the
metric list is taken from an existing summary so that rounds stay comparable), (2) per-opcode stall samples and executed
warp instructions from the source page.  usage: ncu_summary.py <rep> <metrics_template.csv> <out_metrics.csv> [<out_opcodes.csv>]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, template, out_metrics = sys.argv[1:4]
    want = [r[0] for r in list(csv.reader(open(template)))[1:]]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out_metrics, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(["metric", "unit", "value"])
        for m in want:
            if m in col:
                w.writerow([m, units[col[m]], vals[col[m]]])
    if len(sys.argv) > 4:
        src = page(rep, "source", ["--print-source", "sass"])
        hi = next(i for i, r in enumerate(src) if "Source" in r)      # the first line names the kernel
        h, src = src[hi], src[hi:]
        ci = {n: i for i, n in enumerate(h)}
        s_col = next(i for n, i in ci.items() if n.startswith("Source"))
        smp = next(i for n, i in ci.items() if n.startswith("Warp Stall Sampling (All"))
        exe = ci["Instructions Executed"]
        agg = defaultdict(lambda: [0, 0])
        for r in src[1:]:
            if len(r) <= max(s_col, smp, exe):
                continue
            toks = r[s_col].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            try:
                agg[op][0] += int(float(r[smp] or 0)); agg[op][1] += int(float(r[exe] or 0))
            except ValueError:
                pass
        tot = sum(v[0] for v in agg.values()) or 1
        with open(sys.argv[4], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["opcode", "stall_samples", "share", "warp_instructions_executed"])
            for op, (a, b) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
                w.writerow([op, a, round(a / tot, 4), b])


main()
