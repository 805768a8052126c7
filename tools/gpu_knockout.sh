#!/bin/bash
# usage: tools/gpu_knockout.sh "<entry,entry>" ...   ('-' = nothing knocked out) -> gpurun_out/knockout.txt
mkdir -p gpurun_out
: > gpurun_out/knockout.txt
i=0
for k in "$@"; do
  i=$((i+1))
  [ "$k" = "-" ] && k=""
  GS_KNOCKOUT="$k" timeout 150 python tools/knockout_bench.py --steps ${SWEEP_STEPS:-10} --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/ko_$i.json 2> gpurun_out/ko_$i.err
  rc=$?
  python - "$k" $rc gpurun_out/ko_$i.json <<'PY' | tee -a gpurun_out/knockout.txt
import json, sys
k, rc, f = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    r = json.loads(open(f).read().strip().splitlines()[-1])
    print(f"[{k or 'nothing'}] rc={rc} ms/step {r['ms_per_step']:.2f}")
except Exception as ex:
    print(f"[{k}] rc={rc} FAILED {ex}")
PY
done
