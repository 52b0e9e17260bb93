#!/bin/bash
set -x
cd $GRAFT_REPO_ROOT
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_collectives.py > gpurun_out/r02n_collectives.json 2> gpurun_out/r02n_collectives.err
cat gpurun_out/r02n_collectives.json; tail -5 gpurun_out/r02n_collectives.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench_2gpu.json 2> gpurun_out/r02n_bench_2gpu.err
cut -c1-600 gpurun_out/r02n_bench_2gpu.json; tail -3 gpurun_out/r02n_bench_2gpu.err
timeout 600 python -m pytest tests/test_gpu_entries.py -q -k non_current_device 2>&1 | tail -3
