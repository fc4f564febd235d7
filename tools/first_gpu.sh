#!/bin/bash
# first end-to-end GPU pass: peaks, parity tests, probe timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
./build/fp64_peak > gpurun_out/fp64_peak.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1
echo "probe exit $?" >> gpurun_out/probe.txt
tail -5 gpurun_out/fp64_peak.txt; tail -30 gpurun_out/pytest_gpu.txt; tail -12 gpurun_out/probe.txt
