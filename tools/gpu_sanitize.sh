#!/bin/bash
# compute-sanitizer passes over the kernel parity checks (the reference has no race / memory checking).
# Slow (10-50x): run on the small parity groups only.
#   usage: tools/gpu_sanitize.sh [memcheck|racecheck|synccheck|initcheck] ["group group ..."] [per-group timeout s]
TOOL=${1:-memcheck}
GROUPS_=${2:-"tc_1x1 tc_epilogue tc_bn loss small_ops"}
TMO=${3:-600}
mkdir -p gpurun_out
for grp in $GROUPS_; do
  GS_DIAG_TIMEOUT=$TMO timeout $((TMO + 60)) compute-sanitizer --tool $TOOL --target-processes all --error-exitcode 9 \
      --log-file gpurun_out/sanitize_${TOOL}_${grp}.%p.log \
      python tools/gpu_diag.py $grp > gpurun_out/sanitize_${TOOL}_${grp}.out 2>&1
  echo "$TOOL $grp rc=$?"
  grep -h -E "ERROR SUMMARY|== $grp" gpurun_out/sanitize_${TOOL}_${grp}.*.log gpurun_out/sanitize_${TOOL}_${grp}.out | tail -4
done
