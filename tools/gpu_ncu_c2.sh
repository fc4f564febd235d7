#!/bin/bash
mkdir -p gpurun_out
python tools/prof_c2_sampler.py > gpurun_out/prof_c2_plain.log 2>&1 && cat gpurun_out/prof_c2_plain.log | tail -2 &&
timeout 600 ncu --set full --clock-control none -k regex:'ensemble_kernel' -o gpurun_out/prof_c2 -f python tools/prof_c2_sampler.py > gpurun_out/ncu_c2.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_c2.ncu-rep --page raw --csv > gpurun_out/prof_c2_raw.csv 2> gpurun_out/prof_c2.err
rm -f gpurun_out/prof_c2.ncu-rep
ls -la gpurun_out/prof_c2*
