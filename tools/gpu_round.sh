#!/bin/bash
# One GPU-box session: parity diag + pytest + bench + ncu launch list + ncu full capture of the top kernel.
mkdir -p gpurun_out
timeout 400 python tools/gpu_diag.py stage_tc model_tc 2>&1 | tail -40
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --ncu-cycle --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
  timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:igemm_kernel -s 100 -c 3 -f -o gpurun_out/prof_igemm $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; ls -la gpurun_out/
fi
