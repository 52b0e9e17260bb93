#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x -k ff_fused 2>&1 | tail -15 > gpurun_out/r02f_ff_test.log
tail -5 gpurun_out/r02f_ff_test.log
grep -q passed gpurun_out/r02f_ff_test.log || exit 1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02f_tests.log; tail -4 gpurun_out/r02f_tests.log
python tools/parity_diag.py ckpt 2>&1 | grep tables_split | cut -c1-250 > gpurun_out/r02f_diag_ckpt.log; cat gpurun_out/r02f_diag_ckpt.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench.json')); print(d['ms_per_step'], d['breakdown_ms'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02f_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    k=r[ki][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:25]:
    print(f'{k:60s} n={n:4d} total={t/1e3:8.3f} ms avg={t/n:9.1f} us')
PY
