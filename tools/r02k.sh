#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_policy.py -q -x -k "score_table_mode or teacher_forced_logits" 2>&1 | tail -8 > gpurun_out/r02k_t1.log; tail -4 gpurun_out/r02k_t1.log
grep -q "passed" gpurun_out/r02k_t1.log || exit 1
grep -q "failed" gpurun_out/r02k_t1.log && exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02k_tests.log; tail -4 gpurun_out/r02k_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02k_bench.json')); print(d['ms_per_step'], d['breakdown_ms'], d['roofline']['frac'], d['clocks'])"
tail -3 gpurun_out/r02k_bench.err
