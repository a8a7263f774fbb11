#!/bin/bash
# Round 2: sentence token cache in the host feed -- worker tests (outputs still equal the reference's) and the bench.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_worker.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_r2w.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_r2w.log | cut -c1-200
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --json-out gpurun_out/bench_r2w.json > gpurun_out/bench_r2w.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2w.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['clocks']['sm_mhz'])
for k in ('e2e_worker', 'c1'):
    print(k, json.dumps(d['extra'][k])[:600])
print(json.dumps(d['cpu_baseline'])[:700])
PY
