#!/bin/bash
# round 2, call t: ncu --set full with source of the T <= 64 attention kernel inside a 64-frame step
mkdir -p gpurun_out
B="python bench.py --frames 64 --steps 1 --warmup 0 --no-cpu --no-knn --no-e2e --no-a7-ablation"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_attention_mma3' -s 4 -c 1 -o gpurun_out/r2t_attention_full $B > gpurun_out/r2t_ncu.log 2>&1
tail -3 gpurun_out/r2t_ncu.log; ls -la gpurun_out/r2t_attention_full.ncu-rep
