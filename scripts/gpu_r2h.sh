#!/bin/bash
# round 2, call h: ncu source-level captures of the FC / out-proj GEMMs after the smem-staged epilogue operands
B="python bench.py --frames 64 --steps 1 --warmup 0 --no-cpu --no-knn --no-e2e --no-a7-ablation"
NCU="ncu --set full --clock-control none --import-source on"
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<\(int\)5>' -s 3 -c 1 -o gpurun_out/r2h_gemm_fc_lngelu_full $B > gpurun_out/r2h_ncu3.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<\(int\)6>' -s 6 -c 1 -o gpurun_out/r2h_gemm_outproj_full $B > gpurun_out/r2h_ncu4.log 2>&1
ls -la gpurun_out/r2h*.ncu-rep
