#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02g_tests.log; tail -6 gpurun_out/r02g_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench.json')); print(d['ms_per_step'], d['breakdown_ms'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])"
tail -3 gpurun_out/r02g_bench.err
