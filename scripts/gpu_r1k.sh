#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu5.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu5.log
tail -4 gpurun_out/pytest_gpu5.log
timeout 600 python bench.py > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -c 300 gpurun_out/bench_r1k.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_crop|k_frame_rows' \
   --launch-skip 8 --launch-count 4 -o gpurun_out/r1k_crops_full -f python scripts/ncu_ingest.py > gpurun_out/ncu_crops.log 2>&1; tail -2 gpurun_out/ncu_crops.log
