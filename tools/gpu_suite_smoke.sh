#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r02_pytest_all13.log 2>&1
echo "all rc=$?" >> gpurun_out/r02_pytest_all13.log
tail -4 gpurun_out/r02_pytest_all13.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke13.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke13.log
