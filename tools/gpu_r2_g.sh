#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/g_bench_n2.json 2> gpurun_out/g_bench_n2.err
echo "bench rc $?"; tail -c 3000 gpurun_out/g_bench_n2.json; tail -5 gpurun_out/g_bench_n2.err
