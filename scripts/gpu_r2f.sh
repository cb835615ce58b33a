#!/bin/bash
# round 2, call f: c5 flow on one GPU (bounded object stage), new parity tests, ncu --set full of the three fused-epilogue GEMMs
python -m pytest tests/test_gpu_geometry.py tests/test_gpu_reference_golden.py tests/test_gpu_dropin.py -q --timeout 900 2>&1 | tail -12 > gpurun_out/r2f_pytest.log
python bench.py --config c5 --frames 1024 --merge-frames 128 --queries 200 --steps 1 --warmup 1 --no-cpu --no-knn --no-a7-ablation --no-e2e > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err
B="python bench.py --frames 64 --steps 1 --warmup 0 --no-cpu --no-knn --no-e2e --no-a7-ablation"
NCU="ncu --set full --clock-control none --import-source on"
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<\(int\)5>' -s 3 -c 1 -o gpurun_out/r2f_gemm_fc_lngelu_full $B > gpurun_out/r2f_ncu3.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<\(int\)6>' -s 6 -c 2 -o gpurun_out/r2f_gemm_resid_full $B > gpurun_out/r2f_ncu4.log 2>&1
$NCU --kernel-name-base demangled -k 'regex:k_gemm_f16_2sm<\(int\)4>' -s 3 -c 1 -o gpurun_out/r2f_gemm_qkv_ln_full $B > gpurun_out/r2f_ncu5.log 2>&1
ls -la gpurun_out/r2f*.ncu-rep; tail -5 gpurun_out/r2f_pytest.log; tail -c 400 gpurun_out/r2f_bench_c5.err; tail -2 gpurun_out/r2f_ncu3.log
