export GS_COMM_TIMEOUT_S=60
N=${N:-2}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; tail -1 gpurun_out/r2_bench10.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['ms_per_step'])"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for f in 2 0; do
echo "== parity GS_SYNCBN_FOLD=$f"
GS_SYNCBN_FOLD=$f timeout 300 $TR tools/multi_rank_parity.py 2>&1 | grep -E '^\{|Error|error|timed out' | cut -c 1-1700
done
run() { env "$@" timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/tmpN.json 2> gpurun_out/tmpN.err; tail -1 gpurun_out/tmpN.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],2), (d.get('parity_multi') or {}).get('ok'))" || tail -5 gpurun_out/tmpN.err; }
run GS_SYNCBN_FOLD=2
run GS_SYNCBN_FOLD=0
run GS_SYNCBN_FOLD=2 GS_BN_FUSED_BWD=auto GS_BN_FUSED_MAX_MB=12
