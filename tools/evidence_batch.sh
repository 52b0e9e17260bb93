#!/bin/bash
# round-2 evidence batch (1 GPU): full bench line, ncu launch list with DRAM bytes, ncu --set full of the top kernels, sanitizers
set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
tail -c 600 gpurun_out/r02z_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02z_launches_tsp50_b65536.csv python tools/one_rollout.py tsp 50 65536 2 > gpurun_out/r02z_ncu1.log 2>&1
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02z_launches_vrp100_b131072.csv python tools/one_rollout.py vrp 100 131072 1 > gpurun_out/r02z_ncu2.log 2>&1
for spec in "ff:k_ff_fused:4" "qkv_attention:k_qkv_attention:4" "glimpse:^k_step_glimpse$:70" "pointer:^k_step_pointer$:70" "gemm_b:k_gemm_tc4:61" "gemm_out:k_gemm_tc4:55" "score_table_fused:k_score_table_fused:1" "glimpse_first:k_step_glimpse_first:2"; do
  name=${spec%%:*}; rest=${spec#*:}; pat=${rest%%:*}; skip=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c 1 -f -o gpurun_out/r02z_$name python tools/one_rollout.py tsp 50 65536 2 > gpurun_out/r02z_ncu_$name.log 2>&1
  tail -2 gpurun_out/r02z_ncu_$name.log
done
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02z_sanitize_$tool.log 2>&1
  tail -2 gpurun_out/r02z_sanitize_$tool.log
done
ls -la gpurun_out/ | grep r02z
