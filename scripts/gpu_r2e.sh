#!/bin/bash
# round 2, call e: A7 scan with frame-interleaved blocks; c5 flow on one GPU; ncu --set full captures (geometry, A7, GEMM epilogues)
python -m pytest tests/test_gpu_masks3d.py tests/test_gpu_geometry.py -q --timeout 900 2>&1 | tail -8 > gpurun_out/r2e_pytest.log
python bench.py --frames 2048 --steps 1 --warmup 1 --no-cpu --no-knn > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python bench.py --config c5 --frames 512 --queries 200 --steps 1 --warmup 1 --no-cpu --no-knn --no-a7-ablation > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err
B="python bench.py --frames 64 --steps 1 --warmup 0 --no-cpu --no-knn --no-e2e --no-a7-ablation"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k 'regex:^(k_bounds|k_mark|k_accumulate|k_radius_count)$' -c 4 -o gpurun_out/r2e_geometry_full $B > gpurun_out/r2e_ncu1.log 2>&1
$NCU -k 'regex:^(k_m3d_scan|k_nn_winner|k_scatter_batch|k_crop_rows_mma)$' -c 4 -o gpurun_out/r2e_a7_nn_scatter_crops_full $B > gpurun_out/r2e_ncu2.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<5>' -s 3 -c 1 -o gpurun_out/r2e_gemm_fc_lngelu_full $B > gpurun_out/r2e_ncu3.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<6>' -s 6 -c 2 -o gpurun_out/r2e_gemm_resid_full $B > gpurun_out/r2e_ncu4.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<4>' -s 3 -c 1 -o gpurun_out/r2e_gemm_qkv_ln_full $B > gpurun_out/r2e_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail; tail -4 gpurun_out/r2e_pytest.log; tail -c 300 gpurun_out/r2e_bench_c5.err; tail -3 gpurun_out/r2e_ncu3.log
