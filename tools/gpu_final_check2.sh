#!/bin/bash
# whole -m gpu suite again after the collate validation + ncu launch list of one cfg-5 training step (forward + backward)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r02f_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02f_pytest_all.log
tail -4 gpurun_out/r02f_pytest_all.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02f_train_launches.csv python tools/train_profile.py cfg5 > gpurun_out/r02f_train_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r02f_train_launches.csv > gpurun_out/r02f_train_launches_summary.txt 2>&1
head -45 gpurun_out/r02f_train_launches_summary.txt
