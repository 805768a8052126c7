#!/bin/bash
# compute-sanitizer passes over the kernel parity checks (the reference has no race / memory checking).
# Slow (10-50x): run on the small parity groups only.   usage: tools/gpu_sanitize.sh [memcheck|racecheck|synccheck]
TOOL=${1:-memcheck}
mkdir -p gpurun_out
for grp in tc_1x1 tc_epilogue tc_bn elementwise; do
  timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/sanitize_${TOOL}_${grp}.log \
      python tools/gpu_diag.py $grp > gpurun_out/sanitize_${TOOL}_${grp}.out 2>&1
  echo "$TOOL $grp rc=$?"; grep -E "ERROR SUMMARY|checks," gpurun_out/sanitize_${TOOL}_${grp}.log gpurun_out/sanitize_${TOOL}_${grp}.out | tail -2
done
