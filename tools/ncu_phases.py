#!/usr/bin/env python
"""Per-phase stall samples / warp-instructions per frame from an ncu source-page CSV of the fp32 frame kernel.
Usage: python tools/ncu_phases.py src.csv frames"""
import csv, sys
from collections import defaultdict
def f(x):
    try: return float(x)
    except ValueError: return 0.0
def cat(file, line):
    if file == 'fpc_encode_fp32.cu':
        if 70 <= line <= 131: return 'gru gemm+gate epilogue'
        if 176 <= line <= 195: return 'producer'
        if 247 <= line <= 268: return 'fc/residual'
        if 269 <= line <= 312: return 'indicators+scalar'
        if 313 <= line <= 335: return 'vq dispatch/compaction'
        if 336 <= line <= 392: return 'feedback/output'
        return 'kernel other'
    if file in ('fpc_common.cuh', 'sm_20_intrinsics.hpp'): return 'mbarrier/barrier helpers'
    if file == 'fpc_vq_screen.cuh': return 'vq screened'
    if file in ('fpc_vq_search.cuh', 'fpc_vq.cuh'): return 'vq exact/scalar helpers'
    if file == 'fpc_math.cuh': return 'sigmoid/tanh'
    return 'misc'
path, frames = sys.argv[1], float(sys.argv[2])
cur = hdr = None
agg = defaultdict(lambda: [0.0, 0.0])
for row in csv.reader(open(path, errors='replace')):
    if not row: continue
    if row[0] == 'File Path': cur = row[1].split('/')[-1]; hdr = None; continue
    if row[0] == 'Function Name': continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None or row[0] == '': continue
    try: ln = int(row[0])
    except ValueError: continue
    d = dict(zip(hdr, row))
    c = cat(cur, ln)
    agg[c][0] += f(d.get('# Samples')); agg[c][1] += f(d.get('Instructions Executed'))
ts = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %5.1f%%  samples/frame %8.0f   warp-inst/frame %.3g" % (k, 100 * v[0] / ts, v[0] / frames, v[1] / frames))
