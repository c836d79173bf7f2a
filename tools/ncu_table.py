#!/usr/bin/env python
"""One CSV with a column per kernel from several .ncu-rep files (the last captured launch of each), metric rows taken
from an existing table.  usage: ncu_table.py <template.csv> <out.csv> label=rep [label=rep ...]"""
import csv
import io
import subprocess
import sys

template, out = sys.argv[1:3]
want = [r[0] for r in list(csv.reader(open(template)))[1:]]
cols = []
for spec in sys.argv[3:]:
    label, rep = spec.split("=", 1)
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    ci = {h: i for i, h in enumerate(hdr)}
    name = vals[ci["Kernel Name"]].replace("void ", "").replace("mrb::", "").split("(CUtensorMap")[0].split("(mrb::")[0].replace("(int)", "").replace("(bool)", "")
    cols.append((label + ": " + name, {m: (vals[ci[m]] + (" " + units[ci[m]] if units[ci[m]] else "")) for m in want if m in ci}))
with open(out, "w", newline="") as f:
    w = csv.writer(f, quoting=csv.QUOTE_ALL)
    w.writerow(["metric"] + [c[0] for c in cols])
    for m in want:
        w.writerow([m] + [c[1].get(m, "") for c in cols])
