export GS_COMM_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_rank_parity.py > gpurun_out/r2_mrp2.log 2>&1; echo "multi_rank_parity rc=$?"; grep -E "^\{" gpurun_out/r2_mrp2.log | tail -1; grep -E "Error|error" gpurun_out/r2_mrp2.log | head -5 | cut -c 1-300
run() { env "$@" timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/tmp.json 2> gpurun_out/tmp.err; tail -1 gpurun_out/tmp.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],2), (d.get('parity_multi') or {}).get('ok'))" || tail -5 gpurun_out/tmp.err; }
run GS_SYNCBN_FOLD=1
run GS_SYNCBN_FOLD=1 GS_BN_FUSED_BWD=0
run GS_SYNCBN_FOLD=0 GS_BN_FUSED_BWD=0
run GS_SYNCBN_FOLD=1 GS_BN_FUSED_BLOCKS_PER_SM=1
