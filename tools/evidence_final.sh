#!/bin/bash
# final round-2 evidence on the committed binary: tests, smoke, bench line, ncu of the fused attention kernel, sanitizers
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -c 300 gpurun_out/r02f_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_qkv_attention" -s 4 -c 1 -f -o gpurun_out/r02f_qkv_attention python tools/one_rollout.py tsp 50 65536 2 > gpurun_out/r02f_ncu_qkv.log 2>&1
tail -1 gpurun_out/r02f_ncu_qkv.log
timeout 600 ncu --set full --clock-control none -k "regex:k_gemm_tn_tc" -c 1 -f -o gpurun_out/r02f_gemm_tn python tools/gemm_tn_time.py > gpurun_out/r02f_ncu_tn.log 2>&1
tail -1 gpurun_out/r02f_ncu_tn.log
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02f_sanitize_$tool.log 2>&1
  tail -2 gpurun_out/r02f_sanitize_$tool.log
done
