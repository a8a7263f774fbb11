#!/bin/bash
# Round 2: whole GPU suite + the default bench (with the retrieval extra) after the rotary / retrieval changes.
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2m.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_r2m.log | cut -c1-250
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --json-out gpurun_out/bench_r2m.json > gpurun_out/bench_r2m.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_r2m.log | cut -c1-200
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2m.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['clocks'])
for k, v in d['extra'].items():
    print(k, json.dumps(v)[:900])
PY
