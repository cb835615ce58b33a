#!/bin/bash
# round 2, call m: compute-sanitizer memcheck over the parity tests of the kernels written this round
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -q --timeout 500 2>&1 | tail -5 > gpurun_out/r2m_pytest_dropin.log; cat gpurun_out/r2m_pytest_dropin.log
CS="compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 0 --target-processes all"
for t in test_gpu_masks3d test_gpu_features test_gpu_geometry test_gpu_objects; do
  timeout 900 $CS python -m pytest tests/$t.py -x -q --timeout 800 > gpurun_out/r2m_memcheck_$t.log 2>&1; echo "$t rc=$?" | tee -a gpurun_out/r2m_memcheck_summary.log
  grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2m_memcheck_$t.log | tee -a gpurun_out/r2m_memcheck_summary.log
  tail -3 gpurun_out/r2m_memcheck_$t.log
done
timeout 900 $CS python -m pytest tests/test_gpu_encoder.py -x -q --timeout 800 -k "test_vit_forward or test_gemm_tcgen05" > gpurun_out/r2m_memcheck_test_gpu_encoder.log 2>&1; echo "encoder rc=$?" | tee -a gpurun_out/r2m_memcheck_summary.log
tail -3 gpurun_out/r2m_memcheck_test_gpu_encoder.log
grep -h "ERROR SUMMARY" gpurun_out/r2m_memcheck_*.log | sort | uniq -c | tee -a gpurun_out/r2m_memcheck_summary.log
