cd $GRAFT_REPO_ROOT
for tool in synccheck racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02f_sanitize_$tool.log 2>&1
  tail -1 gpurun_out/r02f_sanitize_$tool.log
done
timeout 100 python tools/attn_kernel_once.py 50 65536 4 2>&1 | tail -2
timeout 100 python tools/attn_kernel_once.py 100 32768 3 2>&1 | tail -1
timeout 240 python -m pytest tests/test_gpu_gemm.py -q -x -k "qkv_attention" 2>&1 | tail -1
