#!/bin/bash
# round-2 GPU session 1: whole GPU suite, bench line, timeline tools
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --deselect tests/test_autograd_gpu.py --deselect tests/test_train_gpu.py > gpurun_out/r02_pytest_infer.log 2>&1
echo "infer rc=$?" >> gpurun_out/r02_pytest_infer.log
timeout 900 python -m pytest tests/test_autograd_gpu.py tests/test_train_gpu.py -q --maxfail=40 > gpurun_out/r02_pytest_train.log 2>&1
echo "train rc=$?" >> gpurun_out/r02_pytest_train.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
timeout 300 python tools/phase_times.py > gpurun_out/r02_phase_times.txt 2>&1
timeout 300 python tools/gap_analysis.py > gpurun_out/r02_gap.txt 2>&1
tail -5 gpurun_out/r02_pytest_infer.log; tail -30 gpurun_out/r02_pytest_train.log; cat gpurun_out/r02_bench1.json | head -c 3000
