#!/bin/bash
# round 2, call u (K/V fragments shared by the two query tiles of a warp, staged coalesced output): attention with four warps per (image, head) tile (5 / 6 tiles per SM): parity, A/B inside the step on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_vitl.py tests/test_gpu_reference_golden.py -q --timeout 800 > gpurun_out/r2u_pytest.log 2>&1; tail -4 gpurun_out/r2u_pytest.log
for v in 0 6 0; do
HMSG_ATTN_WIDE=$v timeout 600 python bench.py --frames 2048 --batch 64 --steps 2 --warmup 2 --no-cpu --no-knn --no-e2e --no-a7-ablation > gpurun_out/r2u_bench_w$v.json 2> gpurun_out/r2u_bench_w$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2u_bench_w$v.json").read().strip().splitlines()[-1])
print("wide", $v, "value", round(d["value"],1), "gemm TF", round(d["roofline"]["achieved"],1), "clk", d["clocks"]["sm_mhz"], "attn", round(d["roofline"]["other_kernels_ms_per_step"]["attn"],1))
PY
done
