#!/bin/bash
# 2-GPU validation: sharded ingest == single-GPU ingest, short strong-scaling bench line
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_multigpu.py > gpurun_out/r1m_multigpu_check_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r1m_multigpu_check_2gpu.log
tail -3 gpurun_out/r1m_multigpu_check_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --frames 4096 > gpurun_out/bench_r1m_2gpu.json 2> gpurun_out/bench_r1m_2gpu.err; tail -c 400 gpurun_out/bench_r1m_2gpu.json
