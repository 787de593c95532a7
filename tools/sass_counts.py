#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what a kernel runs on (B200_PROFILING.md "What proves a
Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, bulk async copies -> UBLKCP, packed fp32 FMA ->
FFMA2, 3-input min -> FMNMX3, ...   Reads the built library with cuobjdump (no GPU needed).

    python tools/sass_counts.py [path/to/libfpc_b200.so] > profiles/r2_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "feature-predictor-for-speech-codec_b200", "csrc", "libfpc_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "FFMA2", "FFMA", "FMNMX3", "FMNMX", "VIMNMX",
         "HMMA", "DFMA", "DADD", "DMUL", "REDG", "ATOMG", "RED", "LDS", "LDG", "STG", "LDL", "STL", "SYNCS", "BAR", "REDUX", "NANOSLEEP")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        total[kern] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
                break
        else:
            counts[kern][op.split(".")[0]] += 0
print("# SASS mnemonic counts per kernel of %s (cuobjdump -sass; static counts, not executed counts)" % os.path.basename(lib))
print("# tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, cp.async.bulk = UBLKCP, fma.rn.f32x2 = FFMA2")
for k in sorted(counts, key=lambda k: -total[k]):
    c = counts[k]
    shown = " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])
    print("%-70s %6d instr | %s" % (k[:70], total[k], shown))
