#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest_all13.log 2>&1
echo "all rc=$?" >> gpurun_out/r02_pytest_all13.log
tail -6 gpurun_out/r02_pytest_all13.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke13.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke13.log
timeout 300 python bench.py --cfg cfg5 --train --steps 5 --warmup 3 > gpurun_out/r02_train13.json 2> gpurun_out/r02_train13.err; echo "train rc=$?"; cut -c1-250 gpurun_out/r02_train13.json
timeout 400 python bench.py --cfg cfg4 --steps 3 --warmup 3 --no-phoc --no-cpu-baseline > gpurun_out/r02_cfg4_n1_13.json 2> gpurun_out/r02_cfg4_n1_13.err; echo "cfg4 rc=$?"; cut -c1-250 gpurun_out/r02_cfg4_n1_13.json; tail -2 gpurun_out/r02_cfg4_n1_13.err
