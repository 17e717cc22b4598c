#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py > gpurun_out/b_lstm$i.json 2> gpurun_out/b_lstm$i.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/b_lstm$i.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), round(d['host_loop_ms_per_step'],2), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), round(d['roofline']['instrumented_ms_per_step'],2), d['clocks'])
PY
done
timeout 300 python tools/phase_times.py cfg3 2>&1 | tail -16
