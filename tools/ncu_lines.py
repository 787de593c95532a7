#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
samples, executed warp-instructions and the dominant stall reasons.  Usage:
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = defaultdict(lambda: defaultdict(float))
src_text = {}
tot_s = tot_i = 0.0
for row in csv.reader(open(path, errors="replace")):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) < len(hdr) - 2:
        continue
    d = dict(zip(hdr, row))
    if row[0] == "":                    # SASS rows repeat what the cuda line row already sums
        continue
    try:
        line = int(row[0])
    except ValueError:
        continue
    key = (cur_file, line)
    src_text[key] = row[1].strip()[:90]
    def f(k):
        try: return float(d.get(k, "0") or 0)
        except ValueError: return 0.0
    agg[key]["samples"] += f("# Samples")
    agg[key]["inst"] += f("Instructions Executed")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            agg[key][k] += f(k)
    tot_s += f("# Samples"); tot_i += f("Instructions Executed")
print("total samples %.0f, warp-instructions %.3g" % (tot_s, tot_i))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s | %s" % (
        100 * a["samples"] / max(tot_s, 1), 100 * a["inst"] / max(tot_i, 1), key[0], key[1], src_text[key],
        " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(a["samples"], 1)) for v, k in st)))
