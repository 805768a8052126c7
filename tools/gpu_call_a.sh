#!/bin/bash
# Round-2 session A: PDL bit A/B on the bench, BASELINE config 3 with graph replay, compute-sanitizer passes.
mkdir -p gpurun_out
AB_STEPS=10 bash tools/gpu_ab.sh "-" "GS_PDL=30" "GS_PDL=94" "GS_PDL=126" 2>&1 | grep -v "^    " 
timeout 300 python tools/config_cases.py 3 > gpurun_out/config3.log 2>&1; echo "config3 rc=$?"; tail -3 gpurun_out/config3.log
for tool in memcheck racecheck; do
  grp="conv bn loss misc sgd"; [ $tool = racecheck ] && grp="conv bn loss"
  timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitize_$tool.log \
      python tools/sanitize_target.py $grp > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitize_$tool.out; grep -E "ERROR SUMMARY" gpurun_out/sanitize_$tool.log | tail -1
done
