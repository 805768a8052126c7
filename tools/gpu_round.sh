#!/bin/bash
# One GPU-box session: pytest -m gpu + smoke + bench + kineto kernel breakdown (+ optional ncu full captures).
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-profile --no-infer --graphs 0 --kineto > gpurun_out/bench_kineto.json 2> gpurun_out/bench_kineto.err
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --ncu-cycle --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-infer --graphs 0"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"igemm_kernel|wgrad_kernel|bn_bwd_apply_kernel|bn_bwd_reduce_kernel|bn_apply_kernel" -s 300 -c 10 -f -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out | tail -12
