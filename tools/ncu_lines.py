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

# optional grouping of morph_fused.cu lines into stages: ranges start at marker comments / function heads
if len(sys.argv) > 3 and sys.argv[3] == "stages":
    import os, re
    srcp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcaq_yolo_b200", "csrc", "morph_fused.cu")
    marks = [("helpers", r"^__device__ __forceinline__ int clampi"), ("nms_bin+sobel3", r"^// literal atan2 binning"),
             ("blur", r"^// ---- T1a"), ("adaptive", r"^// ---- T1b"), ("lbp_var", r"^// ---- T1c"), ("act", r"^// ---- T1d"),
             ("mag(T2)", r"^// ---- T2"), ("nms(T3)", r"^// ---- T3"), ("otsu", r"^// Otsu threshold from"),
             ("hysteresis", r"^// ---- T4"), ("kernel prologue", r"^morph_fused_kernel"), ("L minmax", r"^  // ---- L:"),
             ("N normalise", r"^  // ---- N:"), ("T1 dispatch", r"^  // ---- T1:"), ("hist merge+T2 dispatch", r"^  for \(int pr = 0; pr < ns; \+\+pr\) \{  "),
             ("T3 dispatch+otsu call", r"^  // ---- T3:"), ("T4 dispatch", r"^  // ---- T4:"), ("T5 counts", r"^  // ---- T5:"),
             ("phi assembly", r"^  // ---- phi1"), ("nets dispatch", r"^  // all-gather of a per-tile array"), ("host", r"^}  // namespace mcaq")]
    lines = open(srcp).read().split("\n")
    starts = []
    for name, rx in marks:
        for i, l in enumerate(lines, 1):
            if re.search(rx, l):
                starts.append((i, name))
                break
    starts.sort()
    def stage(ln):
        name = "morph_fused.cu:head"
        for st, n in starts:
            if ln >= st:
                name = n
        return name
    g = defaultdict(float); gs = defaultdict(float)
    for (f, ln), v in inst.items():
        name = stage(ln) if f == "morph_fused.cu" else f
        g[name] += v; gs[name] += samp[(f, ln)]
    print()
    for n in sorted(g, key=lambda k: -g[k]):
        print(f"{g[n]:10.0f} {100*g[n]/tot_i:5.1f}%  smp {100*gs[n]/max(tot_s,1):5.1f}%  {n}")
