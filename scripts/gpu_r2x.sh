#!/bin/bash
# round 2, call x: ncu evidence of the final attention kernel (set full + source) and the final launch list of a 128-frame step
mkdir -p gpurun_out
B="python bench.py --frames 64 --steps 1 --warmup 0 --no-cpu --no-knn --no-e2e --no-a7-ablation"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_attention_mma3' -s 4 -c 1 -o gpurun_out/r2x_attention_full $B > gpurun_out/r2x_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name 'regex:^(k_|Device)' -c 1500 --csv --log-file gpurun_out/r2x_launches_128frames.csv python bench.py --frames 128 --steps 1 --warmup 1 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2x_ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2x_launches_128frames.csv > gpurun_out/r2x_launches_128frames_summary.txt 2>&1; head -12 gpurun_out/r2x_launches_128frames_summary.txt
ls -la gpurun_out/r2x_attention_full.ncu-rep
