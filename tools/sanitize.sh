#!/bin/bash
# compute-sanitizer passes over tools/sanitize_tour.py (race detection / sanitizers).
# Each pass runs under its own timeout; outputs land in gpurun_out/sanitize_<tool>.txt.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 120 python tools/sanitize_tour.py > gpurun_out/sanitize_plain.txt 2>&1; echo "plain exit $?" | tee -a gpurun_out/sanitize_plain.txt
for tool in memcheck racecheck synccheck; do
  T=${SAN_TIMEOUT:-240}
  timeout $T $CS --tool $tool --print-limit 30 --error-exitcode 7 python tools/sanitize_tour.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitize_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|TOUR DONE" gpurun_out/sanitize_$tool.txt | tail -3
done
