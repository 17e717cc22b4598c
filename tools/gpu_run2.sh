#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_autograd_gpu.py tests/test_train_gpu.py -q --maxfail=40 > gpurun_out/r02_pytest_train.log 2>&1
echo "train rc=$?" >> gpurun_out/r02_pytest_train.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -k "full_size or degenerate or deterministic_at_batch or predict_goldens" > gpurun_out/r02_pytest_full.log 2>&1
echo "full rc=$?" >> gpurun_out/r02_pytest_full.log
python tools/bench_subword.py > gpurun_out/r02_subword_sync.txt 2>&1
RUART_SUBWORD_ASYNC=1 python tools/bench_subword.py > gpurun_out/r02_subword_async.txt 2>&1
RUART_SUBWORD_ASYNC=1 timeout 300 python -m pytest tests/test_bert_kernels_gpu.py tests/test_model_gpu.py -q -k "not full_size" > gpurun_out/r02_pytest_async.log 2>&1
echo "async rc=$?" >> gpurun_out/r02_pytest_async.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench2_sync.json 2> gpurun_out/r02_bench2.err
RUART_SUBWORD_ASYNC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-phoc --no-cpu-baseline > gpurun_out/r02_bench2_async.json 2>> gpurun_out/r02_bench2.err
timeout 600 python bench.py --train --cfg cfg5 --steps 5 --warmup 3 > gpurun_out/r02_train_n1.json 2> gpurun_out/r02_train_n1.err
tail -25 gpurun_out/r02_pytest_train.log; tail -8 gpurun_out/r02_pytest_full.log; cat gpurun_out/r02_subword_sync.txt gpurun_out/r02_subword_async.txt; tail -3 gpurun_out/r02_pytest_async.log; cat gpurun_out/r02_train_n1.json; tail -5 gpurun_out/r02_train_n1.err
