#!/bin/bash
mkdir -p gpurun_out
RUART_GELU_MODE=2 timeout 300 python tools/bench_gemm.py 2>&1 | tail -11 | tee gpurun_out/r02_gemm12b.txt
