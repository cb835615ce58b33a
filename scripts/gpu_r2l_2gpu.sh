#!/bin/bash
# round 2, call l (2 GPUs): multi-GPU tests with the rewritten equality script, sharded bench line after the rank-0-only result copy
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_reference_golden.py -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2l_pytest_comm.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus 2 --frames 4096 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2l_bench_2gpu_4096frames.json 2> gpurun_out/r2l_bench_2gpu.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2l_bench_reference_2gpu.json 2> gpurun_out/r2l_bench_reference_2gpu.err
tail -5 gpurun_out/r2l_pytest_comm.log; tail -c 300 gpurun_out/r2l_bench_2gpu.err; head -c 1500 gpurun_out/r2l_bench_2gpu_4096frames.json; echo; head -c 300 gpurun_out/r2l_bench_reference_2gpu.json
