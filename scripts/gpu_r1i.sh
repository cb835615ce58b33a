#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_crops.py tests/test_gpu_encoder_vitl.py tests/test_gpu_dropin.py tests/test_gpu_reference_golden.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu3.log
tail -6 gpurun_out/pytest_gpu3.log
timeout 600 python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -c 400 gpurun_out/bench_r1i.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_crop|k_frame_rows' \
   --launch-skip 8 --launch-count 4 -o gpurun_out/r1i_crops_full -f python scripts/ncu_ingest.py > gpurun_out/ncu_crops.log 2>&1; tail -2 gpurun_out/ncu_crops.log
