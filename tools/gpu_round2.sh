#!/bin/bash
# round-2 GPU pass (one GPU): parity tests, smoke, bench (both arms), launch list of the bench,
# full ncu capture of the two headline kernels.  Each ncu run follows a plain run that exited 0.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -30 gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -3 gpurun_out/smoke.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" = "ref" ]; then timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json; fi
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-table"
$CMD > gpurun_out/ncu_plain1.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list exit $?"
python tools/prof_headline.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'predict_var|ensemble_kernel|predict_mean_kernel' -o gpurun_out/prof_headline -f python tools/prof_headline.py > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ncu -i gpurun_out/prof_headline.ncu-rep --page raw --csv > gpurun_out/prof_headline_raw.csv 2> gpurun_out/prof_raw.err
# gpurun copies back at most 64 MiB: the report itself stays on the box when it is large, its raw page travels
if [ $(stat -c %s gpurun_out/prof_headline.ncu-rep) -gt 30000000 ]; then rm gpurun_out/prof_headline.ncu-rep; fi
ls -la gpurun_out | tail -15
tail -n 3 gpurun_out/ncu_launch.log; tail -n 3 gpurun_out/ncu_full.log
