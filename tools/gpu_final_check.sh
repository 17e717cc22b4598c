#!/bin/bash
# whole -m gpu suite + smoke() + default bench line on one fresh box (the round-end driver sequence)
mkdir -p gpurun_out
timeout 560 python -m pytest tests -m gpu -q --maxfail=10 --durations=15 > gpurun_out/r02e_pytest_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r02e_pytest_all.log
tail -30 gpurun_out/r02e_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02e_smoke.log
timeout 300 python bench.py > gpurun_out/r02e_bench_default.json 2> gpurun_out/r02e_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02e_bench_default.err; cut -c1-400 gpurun_out/r02e_bench_default.json
