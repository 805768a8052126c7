#!/bin/bash
# Env-knob sweep on the N=1 bench (no profilers): usage tools/gpu_sweep.sh "<ENV1>" "<ENV2>" ...  ('-' = defaults).
# Prints ms per sandwich cycle for each setting; SWEEP_STEPS (default 10) timed steps, 3 warm-up steps.
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
i=0
for e in "$@"; do
  i=$((i+1))
  [ "$e" = "-" ] && e=""
  env $e timeout 150 python bench.py --steps ${SWEEP_STEPS:-10} --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  rc=$?
  python - "$e" $rc gpurun_out/sweep_$i.json <<'PY' | tee -a gpurun_out/sweep.txt
import json, sys
e, rc, f = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    r = json.loads(open(f).read().strip().splitlines()[-1])
    print(f"[{e or 'defaults'}] rc={rc} ms/step {r['ms_per_step']:.2f} value {r['value']:.1f} e2e {r['e2e']['value']:.1f} host_enqueue {r.get('host_enqueue_ms_per_step', 0):.1f}")
except Exception as ex:
    print(f"[{e}] rc={rc} FAILED {ex}")
PY
done
