#!/bin/bash
# round 2, call o: frame-batch size A/B on one box (64 / 96 / 128 frames per batch), QuickGELU parity test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py -q --timeout 500 -k "quick_gelu or vit_forward" 2>&1 | tail -4 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
for fb in 64 128 96 64; do
  timeout 600 python bench.py --frames 2048 --batch $fb --steps 2 --warmup 2 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2o_bench_fb$fb.json 2> gpurun_out/r2o_bench_fb$fb.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2o_bench_fb$fb.json").read().strip().splitlines()[-1])
print("fb", $fb, "value", round(d["value"],1), "gemm TF", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], d["roofline"].get("other_kernels_ms_per_step"))
PY
done
