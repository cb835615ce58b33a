#!/bin/bash
# round 2, call n (8 GPUs): configs[1] default line after the rank-0-only result copy (e2e)
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR8 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2n_bench_c2_8gpu.json 2> gpurun_out/r2n_bench_c2_8gpu.err
tail -c 400 gpurun_out/r2n_bench_c2_8gpu.err; head -c 1800 gpurun_out/r2n_bench_c2_8gpu.json
