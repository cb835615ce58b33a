#!/bin/bash
# round 2, call y (4 GPUs): configs[1] line at N = 4 with the final kernels
mkdir -p gpurun_out
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551"
timeout 420 $TR4 bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2y_bench_c2_4gpu.json 2> gpurun_out/r2y_bench_c2_4gpu.err
tail -c 300 gpurun_out/r2y_bench_c2_4gpu.err; head -c 1500 gpurun_out/r2y_bench_c2_4gpu.json
