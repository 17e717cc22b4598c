#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02b_bench_default.err; cut -c1-400 gpurun_out/r02b_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02b_bench_reference.json
