export GS_COMM_TIMEOUT_S=60
N=${N:-2}
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3 | cut -c 1-700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { env "$@" timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-infer --no-profile > gpurun_out/tmpN.json 2> gpurun_out/tmpN.err; tail -1 gpurun_out/tmpN.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],2), (d.get('parity_multi') or {}).get('ok'))" || tail -5 gpurun_out/tmpN.err; }
run GS_PDL=30
run GS_PDL=14
run GS_PDL=6
run GS_PDL=30
