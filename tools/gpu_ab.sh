#!/bin/bash
# A/B of env knobs on the bench: usage tools/gpu_ab.sh "<ENV1>" "<ENV2>" ...   (each arg = env assignments, '-' = none)
mkdir -p gpurun_out
i=0
for e in "$@"; do
  i=$((i+1))
  [ "$e" = "-" ] && e=""
  env $e timeout 400 python bench.py --steps ${AB_STEPS:-10} --warmup 3 --no-cpu-baseline --no-infer --kineto > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  echo "== [$e] rc=$?"
  python - <<PY
import json
r=json.loads(open('gpurun_out/ab_$i.json').read().strip().splitlines()[-1])
print('value',round(r['value'],2),'ms/step',round(r['ms_per_step'],2),'e2e',round(r['e2e']['value'],2),'roofline',r.get('roofline'))
try:
    rows=json.load(open('gpurun_out/kineto_kernels.json'))
    for x in rows[:12]: print('   ',x)
except Exception as ex: print(ex)
PY
  cp gpurun_out/kineto_kernels.json gpurun_out/ab_kineto_$i.json 2>/dev/null
done
