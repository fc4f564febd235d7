#!/bin/bash
# final check of the tree: GPU tests, smoke, default bench
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -2 gpurun_out/smoke.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json')); m=d['mcmc']
print(d['value'], d['e2e']['value'], d['roofline']['frac'], {k:m[k] for k in ('value','e2e','kernel_only')}, m['roofline']['frac'], d['extra']['c2']['mcmc_walker_steps_per_s'])
PY
