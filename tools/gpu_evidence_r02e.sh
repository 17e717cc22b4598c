#!/bin/bash
# final-tree evidence on one GPU: ncu launch list of one cfg-3 step, cfg-5 training step line + profiles, reference arm line
T=r02e
mkdir -p gpurun_out
python tools/profile_step.py cfg3 > gpurun_out/${T}_profile_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py cfg3 > gpurun_out/${T}_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt 2>&1
timeout 200 python bench.py --cfg cfg5 --train --steps 10 --warmup 3 > gpurun_out/${T}_train_n1.json 2> gpurun_out/${T}_train_n1.err; echo "train rc=$?"
timeout 200 python tools/train_profile.py cfg5 > gpurun_out/${T}_train_profile.txt 2>&1
timeout 200 python tools/train_host_profile.py cfg5 > gpurun_out/${T}_train_host_profile.txt 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_line.json 2> gpurun_out/${T}_bench_reference.err; echo "ref rc=$?"
head -12 gpurun_out/${T}_launches_summary.txt; cut -c1-300 gpurun_out/${T}_train_n1.json; head -30 gpurun_out/${T}_train_profile.txt; head -60 gpurun_out/${T}_train_host_profile.txt; cut -c1-300 gpurun_out/${T}_bench_reference_line.json
