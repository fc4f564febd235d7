#!/bin/bash
mkdir -p gpurun_out
python tools/var_sched_probe.py > gpurun_out/var_sched.txt 2>&1; cat gpurun_out/var_sched.txt | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'predict_var' --csv --log-file gpurun_out/var_sched_ncu.csv python tools/var_sched_probe.py c4 > gpurun_out/var_sched_ncu.log 2>&1; echo "ncu exit $?"
ENS_NW=8192 python tools/ens_probe.py 2>&1 | grep c5_like | tail -3
ENS_NW=65536 python tools/ens_probe.py 2>&1 | grep c5_like | tail -3
python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_gp.py tests/test_gpu_full_size.py -m gpu -q --timeout 900 2>&1 | tail -5
