export GS_COMM_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
run() { env "$@" timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/tmp8.json 2> gpurun_out/tmp8.err; tail -1 gpurun_out/tmp8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],2), json.dumps(d.get('parity_multi')))" || tail -5 gpurun_out/tmp8.err; }
run GS_SYNCBN_FOLD=0
run GS_SYNCBN_FOLD=1
