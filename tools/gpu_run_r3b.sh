#!/bin/bash
# Round 2: same-box A/B of the whole C2 / C5 step: two-warpgroup attention with its stamps (261, the kernel as it was
# mid-round) vs the shipping kernel (65).
mkdir -p gpurun_out
for v in 261 65 261 65; do
  echo "== B2E_ATT3=$v C2"
  B2E_ATT3=$v timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['clocks']['sm_mhz'])"
  echo "== B2E_ATT3=$v C5"
  B2E_ATT3=$v timeout -s KILL 300 python tools/bench_esm2.py 64 2>/dev/null | tail -n 1 | cut -c1-160
done 2>&1 | tee gpurun_out/step_ab_final.log
