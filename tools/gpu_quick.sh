#!/bin/bash
# parity diag (all groups) + bench, no profilers
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py tc_1x1 tc_3x3 tc_s2 tc_epilogue tc_bn elementwise stage_tc model_tc 2>&1 | tail -40
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
