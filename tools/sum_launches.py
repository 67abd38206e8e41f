#!/usr/bin/env python
"""Sum `ncu --metrics gpu__time_duration.sum --csv` launch lists per kernel name: python tools/sum_launches.py file.csv [nsteps]
(the list covers nsteps identical steps; totals are divided by nsteps)."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    name = re.sub(r"\(.*", "", r[ik])[:90]
    tot[name] += v; cnt[name] += 1
allt = sum(tot.values())
print(f"total {allt / nsteps:.1f} us per step over {sum(cnt.values()) // nsteps} launches")
for n in sorted(tot, key=lambda k: -tot[k])[:30]:
    print(f"{tot[n] / nsteps:9.1f} us {cnt[n] // nsteps:4d}x  {n}")
