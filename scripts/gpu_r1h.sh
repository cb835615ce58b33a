#!/bin/bash
# round-1 GPU session h: parity suite, default bench line, ViT-L/14 probe, ncu of the non-GEMM kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 200 python scripts/probe_vitl.py > gpurun_out/probe_vitl.log 2>&1; tail -3 gpurun_out/probe_vitl.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_crop|k_frame_rows|k_nn_|k_scatter_batch|k_fuse|k_masks_boxes' \
   --launch-skip 18 --launch-count 9 -o gpurun_out/r1h_nongemm_full -f python scripts/ncu_ingest.py > gpurun_out/ncu_nongemm.log 2>&1; tail -3 gpurun_out/ncu_nongemm.log
ls -la gpurun_out
