#!/bin/bash
# round 2, call d: A7 global path without the contended counter, branch-free GELU; c4 shape sanity; launch list of the A7 kernels
python -m pytest tests/test_gpu_masks3d.py tests/test_gpu_dropin.py tests/test_gpu_objects.py tests/test_gpu_encoder.py -q --timeout 900 2>&1 | tail -30 > gpurun_out/r2d_pytest.log
python bench.py --frames 2048 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python bench.py --config c4 --frames 512 --steps 1 --warmup 1 --no-cpu --no-knn > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err
python bench.py --api graph --frames 256 --steps 1 --warmup 1 > gpurun_out/r2d_bench_graph.json 2> gpurun_out/r2d_bench_graph.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(k_m3d|Device|k_gemm|k_crop|k_attention|k_nn|k_scatter)' -c 500 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --frames 128 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2d_ncu_bench.log 2>&1
tail -6 gpurun_out/r2d_pytest.log; tail -c 300 gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench_c4.err; tail -c 300 gpurun_out/r2d_bench_graph.err
