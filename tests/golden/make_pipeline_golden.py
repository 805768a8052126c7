"""Generates tests/golden/pipeline_golden.npz: OpenCV (cv2 4.13, the build container's) outputs of the 8-bit primitives the
reference's train pipeline is made of (mmcv.imrescale == cv2.resize, mmcv.bgr2hsv / hsv2bgr == cv2.cvtColor), on small seeded
inputs, plus the measured agreement of oracle/ref_pipeline.py with them.  Run from the repo root:
    python tests/golden/make_pipeline_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_pipeline as R  # noqa: E402

rng = np.random.default_rng(20261018)
out = {'cv2_version': np.array(cv2.__version__)}
cases = [(48, 64, 31, 41), (48, 64, 24, 32), (40, 56, 77, 107), (64, 96, 50, 75), (32, 48, 64, 96)]
for i, (H, W, nh, nw) in enumerate(cases):
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    seg = rng.integers(0, 20, (H, W), dtype=np.uint8)
    out[f'resize{i}_img'], out[f'resize{i}_seg'], out[f'resize{i}_size'] = img, seg, np.array([nh, nw])
    out[f'resize{i}_linear'] = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    out[f'resize{i}_nearest'] = cv2.resize(seg, (nw, nh), interpolation=cv2.INTER_NEAREST)
bgr = rng.integers(0, 256, (64, 256, 3), dtype=np.uint8)
out['bgr'] = bgr
out['bgr2hsv'] = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
hsv = rng.integers(0, 256, (64, 256, 3), dtype=np.uint8)
hsv[..., 0] %= 180
out['hsv'] = hsv
out['hsv2bgr'] = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'pipeline_golden.npz'), **out)
# report the agreement of the restatement
for i in range(len(cases)):
    a = R.resize_linear_u8(out[f'resize{i}_img'], *out[f'resize{i}_size'])
    d = np.abs(a.astype(int) - out[f'resize{i}_linear'].astype(int))
    print('resize', cases[i], 'linear mismatches', int((d > 0).sum()), 'of', d.size, 'max', int(d.max()), '| nearest exact',
          bool((R.resize_nearest(out[f'resize{i}_seg'], *out[f'resize{i}_size']) == out[f'resize{i}_nearest']).all()))
print('bgr2hsv exact', bool((R.bgr2hsv_u8(bgr) == out['bgr2hsv']).all()))
d = np.abs(R.hsv2bgr_u8(hsv).astype(int) - out['hsv2bgr'].astype(int))
print('hsv2bgr mismatches', int((d > 0).sum()), 'of', d.size, 'max', int(d.max()), '| in the last 16 columns (scalar tail of cv2):',
      int((d[:, -16:] > 0).sum()))
