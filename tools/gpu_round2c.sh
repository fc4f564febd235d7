#!/bin/bash
mkdir -p gpurun_out
./build/tma_gemm_bench > gpurun_out/tma_gemm_bench.txt 2>&1; cat gpurun_out/tma_gemm_bench.txt
python tools/var_sched_probe.py > gpurun_out/var_sched.txt 2>&1; tail -3 gpurun_out/var_sched.txt
python tools/kernel_tour.py > gpurun_out/tour_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tour_launches.csv python tools/kernel_tour.py > gpurun_out/tour_ncu1.log 2>&1; echo "tour exit $?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -6 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['mcmc']['value'], d['mcmc']['e2e'], d['mcmc']['kernel_only'], d['mcmc']['roofline']['frac']); print(d['kernels']['c4']); print(d['extra']['c2'])"
