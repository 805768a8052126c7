#!/bin/bash
# smoke of the reference-named entry points on 1 GPU (tiny run lengths) + per-shape conv profile
set -x
OPT="runner.max_iters=6 log_config.interval=2 checkpoint_config.interval=6 data.workers_per_gpu=0 data.train.length=16 data.test.length=2 data.test.size=[512,1024] data.val.length=2"
timeout 600 python tools/train_supernet.py configs/supernet_fcn_synthetic.py --work-dir /tmp/wd --no-validate --seed 0 --cfg-options $OPT 2>&1 | tail -6
ls -la /tmp/wd
timeout 600 python tools/test_supernet.py configs/supernet_fcn_synthetic.py /tmp/wd/latest.pth --work-dir /tmp/wd/test --cfg-options $OPT 2>&1 | tail -4
cat /tmp/wd/test/metrics.json | cut -c1-300
timeout 600 python tools/extract_subnet.py /tmp/wd/latest.pth /tmp/wd/sub configs/supernet_fcn_synthetic.py 2>&1 | tail -4
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.json
