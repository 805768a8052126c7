#!/bin/bash
# Round profile on ONE GPU: default bench line, then the ncu launch list of one steady-state sandwich cycle
# (gpu__time_duration + DRAM bytes per launch) and one `--set full` capture of a few conv launches mid-cycle.
# Every ncu pass runs only after the same command exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
timeout 900 python bench.py --kineto > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
cp gpurun_out/kineto_kernels.json gpurun_out/kineto_n1.json
CMD="python bench.py --ncu-cycle --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-infer"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"igemm_kernel" -s 120 -c 6 -f \
    -o gpurun_out/prof_igemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -8
