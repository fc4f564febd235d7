#!/bin/bash
# per-instruction samples of the covariance strip kernel (c4 size) for analysis here
mkdir -p gpurun_out
python tools/cov_probe.py > gpurun_out/cov_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cov_strip' -c 2 -o gpurun_out/prof_cov -f python tools/cov_probe.py > gpurun_out/ncu_cov.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_cov.ncu-rep --page source --csv > gpurun_out/prof_cov_source.csv 2> gpurun_out/prof_cov.err
ncu -i gpurun_out/prof_cov.ncu-rep --page raw --csv > gpurun_out/prof_cov_raw.csv 2>> gpurun_out/prof_cov.err
rm -f gpurun_out/prof_cov.ncu-rep
ls -la gpurun_out/prof_cov*
