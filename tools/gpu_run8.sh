#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_gemm8.log 2>&1
echo "gemm rc=$?" >> gpurun_out/r02_pytest_gemm8.log
RUART_GELU_MODE=2 timeout 300 python tools/bench_gemm.py > gpurun_out/r02_gemm8.txt 2>&1
tail -5 gpurun_out/r02_pytest_gemm8.log; cat gpurun_out/r02_gemm8.txt
