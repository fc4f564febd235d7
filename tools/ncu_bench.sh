#!/bin/bash
# ncu evidence for bench.py (one GPU): launch list of the whole command, then a
# full capture of the dominant kernel.  Each ncu run directly follows a plain
# run of the same command line that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain1.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_var -s 20 -c 3 -o gpurun_out/prof_predict_var $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ls -la gpurun_out | tail -12
tail -n 3 gpurun_out/ncu_launch.log; tail -n 3 gpurun_out/ncu_full.log
