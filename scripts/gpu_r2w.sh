#!/bin/bash
# round 2, call w: any-T attention with paired query tiles (head dim 64): parity tests + ViT-L/14 probe (r2p: attention 24.75 of 95.8 ms per 520 images)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder_vitl.py tests/test_gpu_encoder.py tests/test_gpu_parity_sizes.py -q --timeout 800 > gpurun_out/r2w_pytest.log 2>&1; tail -4 gpurun_out/r2w_pytest.log
timeout 600 python scripts/probe_vitl.py > gpurun_out/r2w_probe_vitl.log 2>&1; cp gpurun_out/probe_vitl.json gpurun_out/r2w_probe_vit_l14.json; tail -2 gpurun_out/r2w_probe_vitl.log
