#!/bin/bash
# round 2, call i (2 GPUs): multi-GPU equality tests (C-ABI NCCL collectives and the torch.distributed form), sharded bench lines
python -m pytest tests/test_gpu_comm.py -q --timeout 900 2>&1 | tail -8 > gpurun_out/r2i_pytest_comm.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --frames 4096 --steps 1 --warmup 1 --no-cpu --no-knn > gpurun_out/r2i_bench_2gpu_4096frames.json 2> gpurun_out/r2i_bench_2gpu.err
$TR bench.py --gpus 2 --frames 4096 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation --collective torch > gpurun_out/r2i_bench_2gpu_4096frames_torchcoll.json 2> gpurun_out/r2i_bench_2gpu_torchcoll.err
$TR bench.py --gpus 2 --config c5 --frames 2048 --merge-frames 64 --queries 100 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2i_bench_c5_2gpu.json 2> gpurun_out/r2i_bench_c5_2gpu.err
tail -5 gpurun_out/r2i_pytest_comm.log; tail -c 300 gpurun_out/r2i_bench_2gpu.err; tail -c 300 gpurun_out/r2i_bench_c5_2gpu.err; cat gpurun_out/r2i_bench_2gpu_4096frames.json | head -c 600
