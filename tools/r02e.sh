#!/bin/bash
# round-2 measurement batch: tests, precision probes, bench, sanitizers
set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02e_tests.log
python tools/gemm_error_probe.py > gpurun_out/r02e_gemm_probe.log 2>&1
python tools/parity_diag.py ckpt 2>&1 | cut -c1-330 > gpurun_out/r02e_diag_ckpt.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02e_sanitize_$tool.log 2>&1
  tail -5 gpurun_out/r02e_sanitize_$tool.log
done
tail -3 gpurun_out/r02e_tests.log; cat gpurun_out/r02e_bench.json | cut -c1-3000; tail -5 gpurun_out/r02e_bench.err
