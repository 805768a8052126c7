python -m pytest tests/test_gpu_parity.py tests/test_gpu_path.py -m gpu -q -p no:cacheprovider -x -k "conv or stage or model or full" 2>&1 | tail -3 | cut -c 1-700
for d in 0 1; do GS_IGEMM_DEEP_RING=$d python tools/conv_trace.py 64,128,96,96,3,1,1 64,128,384,96,1,1,1 128,256,64,64,3,1,1 256,512,32,32,3,1,1 64,128,192,192,3,2,1 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('deep=$d', d['shape'], 'last_epi_ns', d['last_epi_ns'], 'cold_us', d['cold_us'])"; done
