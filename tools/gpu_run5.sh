#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r02_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02_pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline --raw-collate --sync-check > gpurun_out/r02_bench5_raw_sync.json 2>> gpurun_out/r02_bench5.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-phoc --no-cpu-baseline --raw-collate > gpurun_out/r02_bench5_raw.json 2>> gpurun_out/r02_bench5.err
timeout 300 python tools/phase_times.py > gpurun_out/r02_phase_times5.txt 2>&1
timeout 300 python tools/train_profile.py cfg5 > gpurun_out/r02_train_profile.txt 2>&1
timeout 300 python bench.py --train --cfg cfg5 --steps 5 --warmup 3 > gpurun_out/r02_train_n1.json 2> gpurun_out/r02_train_n1.err
tail -4 gpurun_out/r02_pytest_all.log; cat gpurun_out/r02_phase_times5.txt | tail -18; head -14 gpurun_out/r02_train_profile.txt; tail -3 gpurun_out/r02_bench5.err
