#!/bin/bash
# ncu evidence only (launch list of the bench + full capture of the headline kernels)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-table"
$CMD > gpurun_out/ncu_plain1.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch-list exit $?"
python tools/prof_headline.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'predict_var|ensemble_kernel|predict_mean_kernel' -o gpurun_out/prof_headline -f python tools/prof_headline.py > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ncu -i gpurun_out/prof_headline.ncu-rep --page raw --csv > gpurun_out/prof_headline_raw.csv 2> gpurun_out/prof_raw.err
rm -f gpurun_out/prof_headline.ncu-rep
ls -la gpurun_out | tail -8
