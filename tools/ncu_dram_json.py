#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> per-kernel DRAM bytes / duration / issue statistics and the per-step DRAM
traffic: python tools/ncu_dram_json.py raw.csv out.json nsteps "how"."""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
nsteps = int(sys.argv[3])
col = {n: i for i, n in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def val(r, name):
    i = col[name]
    try:
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    except ValueError:
        return None


kern = []
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    kern.append({"kernel": name, "grid": r[col.get("Grid Size", 0)] if "Grid Size" in col else None,
                 "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                 "duration_us": val(r, "gpu__time_duration.sum"),
                 "inst_executed": val(r, "smsp__inst_executed.sum") if "smsp__inst_executed.sum" in col else None,
                 "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in col else None,
                 "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in col else None,
                 "registers": val(r, "launch__registers_per_thread") if "launch__registers_per_thread" in col else None})
tot = sum((k["dram_read_bytes"] or 0) + (k["dram_write_bytes"] or 0) for k in kern)
json.dump({"how": sys.argv[4] if len(sys.argv) > 4 else "", "steps_captured": nsteps, "dram_bytes_per_step": tot / nsteps,
           "kernels": kern}, open(sys.argv[2], "w"), indent=1)
print("dram bytes per step", tot / nsteps)
