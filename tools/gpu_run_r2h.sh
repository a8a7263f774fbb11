#!/bin/bash
# Round 2: FP32 instruction-form issue rates, attention after the decode split, and a same-box A/B of the
# whole step (packed layout on / off, attention variant 0 / 5).
mkdir -p gpurun_out
timeout -s KILL 120 tools/bin/fp32_rate > gpurun_out/fp32_rate.log 2>&1; echo "fp32_rate rc=$?"; cat gpurun_out/fp32_rate.log
timeout -s KILL 600 python tools/att_bench.py 0,5,37 > gpurun_out/att_bench_r2h.log 2>&1; echo "att bench rc=$?"; grep "B=" gpurun_out/att_bench_r2h.log
for cfg in "B2E_PACKED=1 B2E_ATT3=5" "B2E_PACKED=0 B2E_ATT3=5" "B2E_PACKED=1 B2E_ATT3=0" "B2E_PACKED=1 B2E_ATT3=5"; do
  echo "== $cfg"
  env $cfg timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['clocks'], d['roofline']['achieved'])"
done 2>&1 | tee gpurun_out/step_ab_r2h.log
