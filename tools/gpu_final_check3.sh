#!/bin/bash
# whole -m gpu suite with the sorted embedding gradient + vectorised subword backward, then the cfg-5 training line and
# the ncu launch list of one training step
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r02g_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02g_pytest_all.log
tail -25 gpurun_out/r02g_pytest_all.log | cut -c1-220
timeout 200 python bench.py --cfg cfg5 --train --steps 10 --warmup 3 > gpurun_out/r02g_train_n1.json 2> gpurun_out/r02g_train_n1.err; echo "train rc=$?"
cut -c1-330 gpurun_out/r02g_train_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02g_train_launches.csv python tools/train_profile.py cfg5 > gpurun_out/r02g_train_profile.txt 2>&1
python tools/summarize_launches.py gpurun_out/r02g_train_launches.csv > gpurun_out/r02g_train_launches_summary.txt 2>&1
head -24 gpurun_out/r02g_train_launches_summary.txt | cut -c1-140; grep -E "eg_|embedding_grad|subword_layers" gpurun_out/r02g_train_launches_summary.txt | cut -c1-140
