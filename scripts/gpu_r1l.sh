#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_crops.py tests/test_gpu_reference_golden.py tests/test_gpu_dropin.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu6.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu6.log
tail -4 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -c 300 gpurun_out/bench_r1l.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_crop|k_frame_rows' \
   --launch-skip 8 --launch-count 4 -o gpurun_out/r1l_crops_full -f python scripts/ncu_ingest.py > gpurun_out/ncu_crops.log 2>&1; tail -2 gpurun_out/ncu_crops.log
