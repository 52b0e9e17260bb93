#!/bin/bash
# Per-kernel durations of three split decode steps of the bench workload (ncu, cold-cache / serialised: use for shares).
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__warps_active.avg.per_cycle_active --clock-control none \
    -k regex:"k_step_glimpse|k_step_pointer|k_gemm_tc4" -s 700 -c 6 --csv --log-file gpurun_out/steps9.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/steps9.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mn=hdr.index('Metric Name'); idc=hdr.index('ID')
per=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)>mv: per.setdefault(r[idc],{'k':r[kn].split('(')[0]})[r[mn]]=float(r[mv].replace(',',''))
for i,d in per.items():
    print(f"{d['k'][:36]:36s} {d['gpu__time_duration.sum']/1e3:8.1f} us  dram read {d['dram__bytes_read.sum']/1e6:8.1f} MB  warps/SM {d['sm__warps_active.avg.per_cycle_active']:5.1f}")
PY
