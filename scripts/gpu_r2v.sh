#!/bin/bash
# round 2, call v: full validation on one B200 -- GPU test suite, smoke(), default bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2v_pytest_gpu.log
tail -5 gpurun_out/r2v_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2v_smoke.log
tail -5 gpurun_out/r2v_smoke.log
timeout 900 python bench.py > gpurun_out/r2v_bench_default.json 2> gpurun_out/r2v_bench_default.err; echo "bench rc=$?"
cat gpurun_out/r2v_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/r2v_bench_reference.json
