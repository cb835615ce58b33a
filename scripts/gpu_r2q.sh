#!/bin/bash
# round 2, call q: error-convention tests; attention with L2 prefetch of the group's next tile (encoder parity + step A/B against r2o: attn 144.9 - 146.8 ms per 2048 frames)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_errors.py -q --timeout 500 > gpurun_out/r2q_pytest_errors.log 2>&1; tail -25 gpurun_out/r2q_pytest_errors.log
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_vitl.py -q --timeout 800 2>&1 | tail -5 > gpurun_out/r2q_pytest_encoder.log; cat gpurun_out/r2q_pytest_encoder.log
for i in 1 2; do
timeout 600 python bench.py --frames 2048 --batch 64 --steps 2 --warmup 2 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2q_bench_$i.json 2> gpurun_out/r2q_bench_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_bench_$i.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "gemm TF", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], d["roofline"].get("other_kernels_ms_per_step"))
PY
done
