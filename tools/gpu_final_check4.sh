#!/bin/bash
# last call of the round: whole -m gpu suite + smoke() + cfg-5 training line on the final library
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r02h_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02h_pytest_all.log
tail -4 gpurun_out/r02h_pytest_all.log | cut -c1-200
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02h_smoke.log
timeout 40 python bench.py --cfg cfg5 --train --steps 10 --warmup 3 > gpurun_out/r02h_train_n1.json 2> gpurun_out/r02h_train_n1.err; echo "train rc=$?"
cut -c1-330 gpurun_out/r02h_train_n1.json
