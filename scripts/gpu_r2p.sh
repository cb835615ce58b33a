#!/bin/bash
# round 2, call p: error-convention tests, ViT-L/14 probe with the round-2 kernels, ncu launch list of the final step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_errors.py -q --timeout 500 > gpurun_out/r2p_pytest_errors.log 2>&1; cat gpurun_out/r2p_pytest_errors.log | tail -6
timeout 600 python scripts/probe_vitl.py > gpurun_out/r2p_probe_vitl.log 2>&1; cp gpurun_out/probe_vitl.json gpurun_out/r2p_probe_vit_l14.json; tail -3 gpurun_out/r2p_probe_vitl.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(k_|Device)' -c 1500 --csv --log-file gpurun_out/r2p_launches_128frames.csv python bench.py --frames 128 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2p_ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2p_launches_128frames.csv > gpurun_out/r2p_launches_128frames_summary.txt 2>&1; head -30 gpurun_out/r2p_launches_128frames_summary.txt
