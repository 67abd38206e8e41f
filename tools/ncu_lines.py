#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line:
python tools/ncu_lines.py file.csv [top]  -> warp-instructions executed and stall samples per line."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file, hdr = None, None
inst = defaultdict(float)
samp = defaultdict(float)
src = {}
tot_i = tot_s = 0.0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        isamp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    if r[2] == "-":      # the per-line summary row (no SASS address): take it
        key = (cur_file, int(r[0]))
        src[key] = r[1]
        try:
            inst[key] += float(r[ii]); samp[key] += float(r[isamp])
            tot_i += float(r[ii]); tot_s += float(r[isamp])
        except ValueError:
            pass
print(f"total warp-instructions {tot_i:.0f}, samples {tot_s:.0f}")
for key in sorted(inst, key=lambda k: -inst[k])[:top]:
    print(f"{inst[key]:10.0f} {100*inst[key]/tot_i:5.1f}%  smp {100*samp[key]/max(tot_s,1):5.1f}%  {key[0]}:{key[1]:<5d} {src[key].strip()[:110]}")

# optional grouping of morph_fused.cu / tile_nets.cuh lines into stages
if len(sys.argv) > 3 and sys.argv[3] == "stages":
    R = [("blur", 144, 190), ("adaptive", 192, 264), ("lbp_var", 266, 350), ("act", 352, 373),
         ("mag(T2)", 375, 407), ("sobel3+nms_bin", 99, 131), ("nms(T3)", 409, 444), ("otsu", 446, 497),
         ("hysteresis", 499, 538), ("div helpers", 84, 97), ("prologue+L load", 543, 716), ("N normalise", 717, 733),
         ("T1 dispatch", 734, 793), ("T2-T4 dispatch", 794, 831), ("T5 counts", 832, 903), ("phi assembly", 904, 964),
         ("nets dispatch", 965, 1035)]
    g = defaultdict(float); gs = defaultdict(float)
    for (f, ln), v in inst.items():
        name = f
        if f == "morph_fused.cu":
            name = "morph_fused.cu:other"
            for n, a, b in R:
                if a <= ln <= b:
                    name = n
                    break
        g[name] += v; gs[name] += samp[(f, ln)]
    print()
    for n in sorted(g, key=lambda k: -g[k]):
        print(f"{g[n]:10.0f} {100*g[n]/tot_i:5.1f}%  smp {100*gs[n]/max(tot_s,1):5.1f}%  {n}")
