#!/bin/bash
# Round profile on ONE GPU: default bench line, then the ncu launch list of one steady-state sandwich cycle
# (gpu__time_duration + DRAM bytes per launch) and `--set full` captures of the hot kernels mid-cycle.
# Every ncu pass runs only after the same command exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
timeout 900 python bench.py --kineto > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
cp gpurun_out/kineto_kernels.json gpurun_out/kineto_n1.json
CMD="python bench.py --ncu-cycle --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-infer"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
# `--set full` captures mid-cycle (MAX iteration).  Launch indices: forward convs 35..123 = stage 3 (1x1 1280->320, 3x3 320 d2,
# 1x1 320->1280); conv launches 174..350 = backward of stage 3 (dgrad + wgrad interleaved); same regions for the DynBN kernels.
cap() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 -f \
      -o gpurun_out/$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "ncu full $1 rc=$?"; tail -1 gpurun_out/ncu_$1.log
}
cap prof_conv_fwd "igemm_kernel" 60 6
cap prof_conv_bwd "igemm_kernel|wgrad_kernel" 220 8
cap prof_bn_fwd "bn_apply_kernel" 60 4
cap prof_bn_bwd "bn_bwd_reduce_kernel|bn_bwd_apply_kernel" 100 8
ls -la gpurun_out | tail -8
