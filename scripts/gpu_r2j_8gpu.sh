#!/bin/bash
# round 2, call j (8 GPUs): sharded == single-GPU check, configs[1] default line, configs[3] (c4: 50 k x 1280x720), configs[4] (c5)
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
$TR8 scripts/check_multigpu.py --collective c --frames 96 > gpurun_out/r2j_check_8gpu_c.log 2> gpurun_out/r2j_check_8gpu_c.err
$TR8 scripts/check_multigpu.py --collective torch --frames 96 > gpurun_out/r2j_check_8gpu_torch.log 2> gpurun_out/r2j_check_8gpu_torch.err
$TR8 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2j_bench_c2_8gpu.json 2> gpurun_out/r2j_bench_c2_8gpu.err
$TR8 bench.py --gpus 8 --config c4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-knn > gpurun_out/r2j_bench_c4_8gpu.json 2> gpurun_out/r2j_bench_c4_8gpu.err
$TR8 bench.py --gpus 8 --config c5 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2j_bench_c5_8gpu.json 2> gpurun_out/r2j_bench_c5_8gpu.err
cat gpurun_out/r2j_check_8gpu_c.log gpurun_out/r2j_check_8gpu_torch.log; grep -h "Error\|error" gpurun_out/r2j_*.err | tail -5
for f in c2 c4 c5; do head -c 300 gpurun_out/r2j_bench_${f}_8gpu.json; echo; done
