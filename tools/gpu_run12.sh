#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_bert_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_12.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r02_pytest_12.log
timeout 200 python tools/bench_qkv_attn.py 2>&1 | tail -1 | tee gpurun_out/r02_qkv_attn_12.txt
RUART_GELU_MODE=2 timeout 300 python tools/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r02_gemm12.txt
