#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 0 1; do echo "RUART_QA_DEBUG=$d"; RUART_QA_DEBUG=$d timeout 200 python tools/bench_qkv_attn.py 2>&1 | tail -1; done | tee gpurun_out/r02_qkv_attn_12.txt
