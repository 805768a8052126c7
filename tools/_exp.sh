run() { env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json,sys; d=json.load(open('gpurun_out/tmp.json')); print('$*', round(d['value'],1), round(d['ms_per_step'],2))" || tail -5 gpurun_out/tmp.err; }
run GS_BN_FUSED_COOP=0
run GS_BN_FUSED_COOP=0 GS_BN_FUSED_BLOCKS_PER_SM=1
run GS_BN_FUSED_COOP=0 GS_WGRAD_STREAM=0
run GS_BN_FUSED_BWD=0 GS_WGRAD_STREAM=0
