python -m pytest tests/test_gpu_path.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "fcn_head or model or smoke" 2>&1 | tail -4 | cut -c 1-600
