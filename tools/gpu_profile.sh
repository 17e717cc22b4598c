#!/bin/bash
# round-2 ncu evidence (final code of the round): launch list of one bench step + --set full captures of the
# BERT kernels (fused qkv+attention, the three folded-LayerNorm GEMMs, subword mix) and of the SDNet-stack kernels
T=${1:-r02b}
mkdir -p gpurun_out
python tools/profile_step.py cfg3 > gpurun_out/${T}_profile_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py cfg3 > gpurun_out/${T}_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launches_summary.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_bf16_2cta|qkv_attn_2cta|subword_avg_layers_fold|bert_embed_raw" \
    --launch-skip 16 --launch-count 12 -f -o gpurun_out/${T}_full python tools/profile_step.py cfg3 > gpurun_out/${T}_ncu_full.log 2>&1
python tools/summarize_ncu_full.py gpurun_out/${T}_full.ncu-rep > gpurun_out/${T}_ncu_full_summary.txt 2>&1
ncu --set full --clock-control none --profile-from-start off \
    -k regex:"subword_avg_layers_fold|lstm_recurrence|attention_tail_mma|seq_tiles|split_concat" --launch-count 10 -f -o gpurun_out/${T}_full2 \
    python tools/profile_step.py cfg3 > gpurun_out/${T}_ncu_full2.log 2>&1
python tools/summarize_ncu_full.py gpurun_out/${T}_full2.ncu-rep > gpurun_out/${T}_ncu_full2_summary.txt 2>&1
rm -f gpurun_out/${T}_full2.ncu-rep
ls -la gpurun_out/${T}_full.ncu-rep; cat gpurun_out/${T}_launches_summary.txt | head -40; cat gpurun_out/${T}_ncu_full_summary.txt; cat gpurun_out/${T}_ncu_full2_summary.txt
