#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_bert_kernels_gpu.py tests/test_model_gpu.py tests/test_sdnet_kernels_gpu.py tests/test_canary_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_11.log 2>&1
echo "rc=$?" >> gpurun_out/r02_pytest_11.log
tail -4 gpurun_out/r02_pytest_11.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err
tail -3 gpurun_out/r02_bench11.err; cut -c1-300 gpurun_out/r02_bench11.json
timeout 300 python tools/phase_times.py > gpurun_out/r02_phase_times11.txt 2>&1; tail -16 gpurun_out/r02_phase_times11.txt
