python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_tests8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tests8.log | cut -c 1-400
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; tail -1 gpurun_out/r2_bench8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline_step']['frac']); b=d['breakdown']
for k,v in list(b['c_abi_calls_ms'].items())[:9]: print(k,v)"
