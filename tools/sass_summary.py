#!/usr/bin/env python3
"""Blackwell-native evidence from the built library: per kernel, the SASS instructions that only tcgen05 / TMA / TMEM /
cluster code produces (B200_PROFILING.md, "What proves a Blackwell-native kernel"), with counts and the first occurrence.
usage: tools/sass_summary.py [lib.so] [kernel-name-regex] > profiles/rNN_sass_<kernel>.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mixed-gemmul8_b200/libgemmul8_b200.so"
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "oz_gemm")
KEY = re.compile(r"\b(UTC[A-Z0-9]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAPF[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|"
                 r"SYNCS[.\w]*|UCGABAR[.\w]*|ATOMG[.\w]*CAS[.\w]*|ACQBULK|USETMAXREG[.\w]*|DFMA[.\w]*|DADD[.\w]*|HMMA[.\w]*|IMMA[.\w]*)")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, stats, first, total = None, {}, {}, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn is None or not pat.search(fn):
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if not m:
        continue
    total[fn] += 1
    k = KEY.search(m.group(2))
    if k:
        stats.setdefault(fn, collections.Counter())[k.group(1)] += 1
        first.setdefault((fn, k.group(1)), f"/*{m.group(1)}*/ {m.group(2)};")
for fn in sorted(stats):
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    print(f"== {dem[:180]}\n   ({total[fn]} SASS instructions)")
    for op, c in sorted(stats[fn].items()):
        print(f"   {c:5d}  {op:44s} first: {first[(fn, op)][:110]}")
