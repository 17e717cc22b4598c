#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bert_kernels_gpu.py -m gpu -q -x -k "qkv or folded" > gpurun_out/r02_pytest_12.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/r02_pytest_12.log
timeout 200 python tools/bench_qkv_attn.py 2>&1 | tail -1 | tee gpurun_out/r02_qkv_attn_12.txt
