#!/bin/bash
# ncu evidence for every kernel family (one GPU).  Each ncu run follows a plain
# run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python tools/kernel_tour.py"
$CMD > gpurun_out/tour_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/tour_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tour_launches.csv $CMD > gpurun_out/tour_ncu1.log 2>&1
echo "launch-list exit $?"
timeout 1500 ncu --set full --clock-control none \
  -k regex:'cov_kernel|cov_strip|predict_var_split|chol_dataflow|trsv_dataflow|kinv_kernel|grad_tiles|predict_mean|predict_var|tri_gemm|predict_grad_kernel|utility_kernel|ensemble_kernel|trinv_kernel' \
  -o gpurun_out/tour_full -f $CMD > gpurun_out/tour_ncu2.log 2>&1
echo "full exit $?"
# the report itself can exceed what gpurun copies back: keep its raw page as CSV
ncu -i gpurun_out/tour_full.ncu-rep --page raw --csv > gpurun_out/tour_full_raw.csv 2> gpurun_out/tour_raw.err
if [ $(stat -c %s gpurun_out/tour_full.ncu-rep) -gt 40000000 ]; then rm gpurun_out/tour_full.ncu-rep; fi
ls -la gpurun_out | grep tour
tail -3 gpurun_out/tour_ncu2.log
