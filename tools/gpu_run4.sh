#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r02_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02_pytest_all.log
python tools/bench_ln.py > gpurun_out/r02_ln.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
timeout 300 python tools/phase_times.py > gpurun_out/r02_phase_times4.txt 2>&1
timeout 300 python tools/train_profile.py cfg5 > gpurun_out/r02_train_profile.txt 2>&1
tail -4 gpurun_out/r02_pytest_all.log; cat gpurun_out/r02_ln.txt; cat gpurun_out/r02_phase_times4.txt | tail -20; head -12 gpurun_out/r02_train_profile.txt
