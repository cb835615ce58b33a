#!/bin/bash
# evidence run: full GPU parity suite, default bench line, ncu launch list of a reduced bench invocation
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r1n_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r1n_pytest_gpu.log
tail -4 gpurun_out/r1n_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; tail -c 300 gpurun_out/bench_r1n.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1n_bench_launches_640frames.csv \
   python bench.py --steps 1 --warmup 0 --frames 640 --batch 64 --no-cpu --no-e2e --no-knn > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/r1n_bench_launches_640frames.csv
