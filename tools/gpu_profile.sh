#!/bin/bash
# round-2 ncu evidence: launch list of one bench step + --set full captures of the dominant and the changed kernels
mkdir -p gpurun_out
python tools/profile_step.py cfg3 > gpurun_out/r02_profile_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches.csv python tools/profile_step.py cfg3 > gpurun_out/r02_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
# --set full: the CTA-pair GEMM (two layers = 8 launches), LayerNorm, BERT attention, subword mean, attention tails
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_bf16_2cta|ln_bf16|bert_attention_mma16_async|subword_avg_layers_async|split_concat" \
    --launch-skip 40 --launch-count 24 -f -o gpurun_out/r02_full python tools/profile_step.py cfg3 > gpurun_out/r02_ncu_full.log 2>&1
python tools/summarize_ncu_full.py gpurun_out/r02_full.ncu-rep > gpurun_out/r02_ncu_full_summary.txt 2>&1
ncu --set full --clock-control none --profile-from-start off \
    -k regex:"subword_avg_layers_async|lstm_recurrence2|attention_tail_mma" --launch-count 8 -f -o gpurun_out/r02_full2 \
    python tools/profile_step.py cfg3 > gpurun_out/r02_ncu_full2.log 2>&1
python tools/summarize_ncu_full.py gpurun_out/r02_full2.ncu-rep > gpurun_out/r02_ncu_full2_summary.txt 2>&1
rm -f gpurun_out/r02_full2.ncu-rep
ls -la gpurun_out/r02_full.ncu-rep; cat gpurun_out/r02_launches_summary.txt | head -40; cat gpurun_out/r02_ncu_full_summary.txt; cat gpurun_out/r02_ncu_full2_summary.txt
