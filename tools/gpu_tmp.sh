#!/bin/bash
timeout 300 python -m pytest tests/test_sdnet_kernels_gpu.py tests/test_canary_gpu.py -x -q 2>&1 | tail -2
for k in 2 3 0; do echo "KLO=$k"; RUART_LSTM_KLO=$k timeout 120 python tools/bench_lstm.py; done
