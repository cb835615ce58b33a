#!/bin/bash
# round 2, call b: folded-LN encoder + bucketed A7 sort: parity tests, A/B bench lines, launch list
python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_vitl.py tests/test_gpu_masks3d.py tests/test_gpu_features.py tests/test_gpu_dropin.py tests/test_gpu_reference_golden.py tests/test_gpu_comm.py -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2b_pytest.log
python bench.py --frames 2048 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
HMSG_LN_FOLD=0 python bench.py --frames 2048 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2b_bench_nofold.json 2> gpurun_out/r2b_bench_nofold.err
python bench.py --api graph --frames 256 --steps 1 --warmup 1 > gpurun_out/r2b_bench_graph.json 2> gpurun_out/r2b_bench_graph.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --frames 128 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2b_ncu_bench.log 2>&1
tail -8 gpurun_out/r2b_pytest.log; tail -c 400 gpurun_out/r2b_bench.err; tail -c 400 gpurun_out/r2b_bench_graph.err
