#!/bin/bash
# compute-sanitizer evidence.  PYTORCH_NO_CUDA_MEMORY_CACHING=1: every torch
# allocation is its own cudaMalloc, so an out-of-bounds access cannot hide inside the caching allocator's pool.
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
CS=/usr/local/cuda/bin/compute-sanitizer
(timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()"; echo "exit=$?") > gpurun_out/r02_sanitizer_memcheck_smoke.txt 2>&1
(timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_forward.py cfg1; echo "exit=$?") > gpurun_out/r02_sanitizer_memcheck_cfg1_streams.txt 2>&1
(timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_forward.py small train; echo "exit=$?") > gpurun_out/r02_sanitizer_memcheck_train.txt 2>&1
(RUART_SANITIZE_BERT_LAYERS=2 timeout 1500 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python tools/sanitize_forward.py small; echo "exit=$?") > gpurun_out/r02_sanitizer_racecheck_small_streams.txt 2>&1
(RUART_SANITIZE_BERT_LAYERS=2 timeout 900 $CS --tool synccheck --error-exitcode 9 --print-limit 20 python tools/sanitize_forward.py small; echo "exit=$?") > gpurun_out/r02_sanitizer_synccheck_small.txt 2>&1
for f in gpurun_out/r02_sanitizer_*.txt; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit=|ok|Error|error" $f | head -12; done
