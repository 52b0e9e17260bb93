#!/usr/bin/env python
"""profiles/rollout_traffic_<kind><N>_b<B>.json from an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch): the DRAM bytes of the decode launches of the LAST rollout in the capture — every
launch after the last score-table kernel (k_score_table) up to k_rollout_finish, each one MEASURED (no computed shares).

    python tools/traffic_from_ncu.py <launches.csv> <kind> <N> <B> <out.json>
"""
import collections
import csv
import json
import sys

path, kind, N, B, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ci = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launches = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ci["Metric Unit"]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
             "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1.0)
    d = launches.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]]})
    d[r[ci["Metric Name"]]] = v * scale
ids = list(launches)
last_table = max(i for i in ids if "k_score_table" in launches[i]["name"])
finish = max(i for i in ids if "k_rollout_finish" in launches[i]["name"])
assert finish > last_table
sel = [launches[i] for i in ids if last_table < i <= finish]
per = collections.OrderedDict()
for d in sel:
    k = d["name"].split("(")[0].replace("void ", "")
    a = per.setdefault(k, {"launches": 0, "us": 0.0, "read": 0.0, "write": 0.0})
    a["launches"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["read"] += d.get("dram__bytes_read.sum", 0.0)
    a["write"] += d.get("dram__bytes_write.sum", 0.0)
rd, wr = sum(a["read"] for a in per.values()), sum(a["write"] for a in per.values())
# the weight transposes / splits run before the timed bracket of the decode loop (negligible bytes); kept in the list
json.dump({
    "kernel": "decode loop: every launch after the score-table prologue (steps 0-1: mean / gather, query-fold GEMMs, first-step "
              "glimpse; steps >= 2: k_step_glimpse, GEMM-B, k_step_pointer)",
    "config": f"{kind} N={N} B={B} greedy, score-table mode, split-step launches",
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "per_kernel": per,
    "source": f"{path}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; "
              "every launch measured, summed over the decode launches of one rollout.  ncu serialises the kernels and "
              "flushes caches between replays, so L2 reuse between consecutive kernels is not credited (upper bound).",
}, open(out, "w"), indent=1)
print(out, f"{(rd + wr) / 1e9:.2f} GB over {len(sel)} launches")
for k, a in per.items():
    print(f"  {k[:50]:50s} n={a['launches']:3d} {a['us'] / 1e3:8.3f} ms  rd {a['read'] / 1e9:7.2f} GB wr {a['write'] / 1e9:6.2f} GB")
