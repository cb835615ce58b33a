#!/bin/bash
# round 2, call c: per-job on-chip A7, folded-LN encoder: full GPU suite, bench A/B, launch list of our kernels
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2c_pytest.log
python bench.py --frames 2048 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python bench.py --api graph --frames 256 --steps 1 --warmup 1 > gpurun_out/r2c_bench_graph.json 2> gpurun_out/r2c_bench_graph.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(k_|Device)' -c 700 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --frames 128 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2c_ncu_bench.log 2>&1
tail -8 gpurun_out/r2c_pytest.log; tail -c 300 gpurun_out/r2c_bench.err; tail -c 600 gpurun_out/r2c_bench_graph.err
