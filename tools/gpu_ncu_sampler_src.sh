#!/bin/bash
# per-instruction stall samples of the wide sampler kernel (c5 share) for analysis here: the report is too large
# to travel, its source page (SASS + samples) is exported on the box
mkdir -p gpurun_out
python tools/prof_headline.py c5 > gpurun_out/prof_plain_c5.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ensemble_kernel' -c 1 -o gpurun_out/prof_sampler -f python tools/prof_headline.py c5 > gpurun_out/ncu_sampler.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_sampler.ncu-rep --page source --csv > gpurun_out/prof_sampler_source.csv 2> gpurun_out/prof_sampler_src.err
ncu -i gpurun_out/prof_sampler.ncu-rep --page raw --csv > gpurun_out/prof_sampler_raw.csv 2>> gpurun_out/prof_sampler_src.err
rm -f gpurun_out/prof_sampler.ncu-rep
ls -la gpurun_out/prof_sampler*
