#!/bin/bash
# standard GPU pass: parity tests, smoke, bench (both arms), probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -3 gpurun_out/smoke.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" = "probe" ]; then timeout 600 python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/probe.txt; tail -12 gpurun_out/probe.txt; fi
