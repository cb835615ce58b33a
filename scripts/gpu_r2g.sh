#!/bin/bash
# round 2, call g: GEMM epilogues with smem-staged column vectors + prefetched row statistics; head dim 80 tower
python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_vitl.py tests/test_gpu_geometry.py -q --timeout 900 2>&1 | tail -15 > gpurun_out/r2g_pytest.log
python bench.py --frames 2048 --steps 1 --warmup 1 --no-cpu --no-knn > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
B="python bench.py --frames 64 --steps 1 --warmup 0 --no-cpu --no-knn --no-e2e --no-a7-ablation"
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name 'regex:^(k_gemm|k_rowstats|k_attention)' -c 80 --csv --log-file gpurun_out/r2g_gemm_launches.csv $B > gpurun_out/r2g_ncu.log 2>&1
tail -8 gpurun_out/r2g_pytest.log; tail -c 300 gpurun_out/r2g_bench.err
