#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_canary_gpu.py -q > gpurun_out/r02_pytest_canary.log 2>&1
echo "canary rc=$?" >> gpurun_out/r02_pytest_canary.log
tail -30 gpurun_out/r02_pytest_canary.log
bash tools/gpu_profile.sh
