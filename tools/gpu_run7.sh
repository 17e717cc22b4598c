#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r02_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02_pytest_all.log
python tools/bench_subword.py > gpurun_out/r02_subword_async16.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err
timeout 300 python tools/phase_times.py > gpurun_out/r02_phase_times7.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches.csv python tools/profile_step.py cfg3 > gpurun_out/r02_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
tail -4 gpurun_out/r02_pytest_all.log; cat gpurun_out/r02_subword_async16.txt; tail -16 gpurun_out/r02_phase_times7.txt; head -24 gpurun_out/r02_launches_summary.txt
