#!/bin/bash
# Round 2: final validation of the tree as committed: whole GPU suite, smoke(), default bench, reference arm.
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_final.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout -s KILL 900 python bench.py --json-out gpurun_out/bench_final2_r02.json > gpurun_out/bench_final2_r02.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_final2_r02.log | cut -c1-250
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final_r02.json 2> gpurun_out/bench_ref_final_r02.err; echo "reference arm rc=$?"; tail -n 1 gpurun_out/bench_ref_final_r02.json | cut -c1-400
