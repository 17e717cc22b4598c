#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bert_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_9b.log 2>&1
echo "rc=$?" >> gpurun_out/r02_pytest_9b.log
tail -25 gpurun_out/r02_pytest_9b.log
RUART_GELU_MODE=2 timeout 300 python tools/bench_gemm.py > gpurun_out/r02_gemm9.txt 2>&1
tail -6 gpurun_out/r02_gemm9.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err
tail -3 gpurun_out/r02_bench9.err; cut -c1-600 gpurun_out/r02_bench9.json
