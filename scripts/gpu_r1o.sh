#!/bin/bash
# attention trim validation + short bench + ncu launch list of the ingest kernels inside bench.py
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_vitl.py tests/test_gpu_reference_golden.py -m gpu -q -p no:cacheprovider > gpurun_out/r1o_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r1o_pytest_gpu.log
tail -3 gpurun_out/r1o_pytest_gpu.log
timeout 300 python bench.py --frames 2048 --no-cpu --no-knn --no-e2e > gpurun_out/bench_r1o_2048.json 2> gpurun_out/bench_r1o.err; tail -c 700 gpurun_out/bench_r1o_2048.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r1o_bench_launches_640frames.csv \
   python bench.py --steps 1 --warmup 0 --frames 640 --batch 64 --no-cpu --no-e2e --no-knn > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/r1o_bench_launches_640frames.csv
