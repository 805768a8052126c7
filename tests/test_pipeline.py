"""Training data pipeline.  CPU: the numpy restatement (oracle/ref_pipeline.py) against the OpenCV golden
vectors, and the product's host logic (per-sample random decisions) against the restatement, bit for bit.  GPU: the fused
CUDA pipeline against the restatement -- crop choice, flip, label map and image values bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_pipeline as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'pipeline_golden.npz')


def test_restatement_matches_the_opencv_golden_vectors():
    g = np.load(GOLD)
    i = 0
    while f'resize{i}_img' in g:
        img, seg, (nh, nw) = g[f'resize{i}_img'], g[f'resize{i}_seg'], g[f'resize{i}_size']
        assert (R.resize_nearest(seg, int(nh), int(nw)) == g[f'resize{i}_nearest']).all()            # label maps: exact
        d = np.abs(R.resize_linear_u8(img, int(nh), int(nw)).astype(int) - g[f'resize{i}_linear'].astype(int))
        if nh <= img.shape[0]:
            assert d.max() == 0, (i, int((d > 0).sum()))      # down-scaling (the fixed-point path restated): exact
        else:
            assert d.max() <= 1 and (d > 0).mean() < 0.01     # up-scaling: cv2 4.13's vector path differs by 1 LSB on < 1 %
        i += 1
    assert i >= 5
    assert (R.bgr2hsv_u8(g['bgr']) == g['bgr2hsv']).all()                                              # integer HSV: exact
    d = np.abs(R.hsv2bgr_u8(g['hsv']).astype(int) - g['hsv2bgr'].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3      # fp32 HSV->BGR, truncated like cv2's vector body


def test_host_draws_equal_the_restatement(gs):
    """gaia_seg_b200.data_pipeline.GpuTrainPipeline.draw == oracle.ref_pipeline.draw_params, every integer and float."""
    from gaia_seg_b200.data_pipeline import GpuTrainPipeline
    pipe = GpuTrainPipeline(seed=7)
    seen_flip, seen_small = set(), False
    for sample in range(300):
        H0, W0 = (1024, 2048) if sample % 3 else (600 + sample, 900 + 2 * sample)
        p, q = pipe.draw(sample, H0, W0), R.draw_params(7, sample, H0, W0)
        assert (p.new_h, p.new_w, p.crop_h, p.crop_w) == (q['new_h'], q['new_w'], q['crop_h'], q['crop_w'])
        assert [(p.box_y[t], p.box_x[t]) for t in range(11)] == q['boxes']
        assert bool(p.flip) == q['flip']
        assert bool(p.has_brightness) == (q['brightness'] is not None) and bool(p.contrast_first) == q['contrast_first']
        if q['brightness'] is not None:
            assert np.float32(p.brightness) == np.float32(q['brightness'])
        if q['contrast'] is not None:
            assert np.float32(p.contrast) == np.float32(q['contrast'])
        if q['saturation'] is not None:
            assert np.float32(p.saturation) == np.float32(q['saturation'])
        assert (bool(p.has_hue), p.hue if p.has_hue else None) == (q['hue'] is not None, q['hue'])
        seen_flip.add(bool(p.flip))
        seen_small |= p.crop_h < 512 or p.crop_w < 1024
        for t in range(11):
            assert 0 <= p.box_y[t] <= p.new_h - p.crop_h and 0 <= p.box_x[t] <= p.new_w - p.crop_w
    assert seen_flip == {True, False} and seen_small      # both flip states and the pad path are exercised


def test_crop_redraw_rule_of_the_restatement():
    """RandomCrop's loop: the first candidate whose dominant class covers < 75 % of the non-ignored pixels wins; a map with
    one class never satisfies it -> the 11th (unchecked) candidate is used; cat_max_ratio = 1 disables the loop."""
    seg = np.zeros((600, 1100), np.uint8)
    seg[:, 550:] = 3
    p = dict(crop_h=512, crop_w=1024, boxes=[(0, 0)] * 11)
    assert R.choose_crop(seg, p) == 0                              # 2 classes, ~50 / 50
    p['boxes'] = [(0, 0)] * 4 + [(10, 60)] * 7
    seg2 = np.zeros((600, 1100), np.uint8)
    seg2[:, 1000:] = 5                                             # box (0,0): 97.6 % class 0; box (10,60): 92 % -> all fail
    assert R.choose_crop(seg2, p) == 10
    seg2[:, 700:] = 5                                              # (0,0): 68 % -> passes at t = 0
    assert R.choose_crop(seg2, p) == 0
    seg3 = np.full((600, 1100), 255, np.uint8)
    seg3[:300] = 1
    assert R.choose_crop(seg3, p) == 10                            # a single non-ignored class
    assert R.choose_crop(seg3, p, cat_max_ratio=1.0) == 0


# samples of the GPU test: Cityscapes-sized, small (-> padded) and elongated images; block layouts from fine to one class
# per crop so that the re-draw loop runs 0 .. 10 times
SIZES = [(1024, 2048), (1024, 2048), (300, 1200), (1024, 2048), (700, 1500), (1024, 2048), (1024, 2048), (1024, 2048),
         (200, 1600), (1024, 2048)]
BLOCKS = [(128, 256), (700, 1400), (100, 300), (1024, 2048), (350, 800)]


def test_samples_of_the_gpu_test_cover_redraws_flips_padding_and_every_distortion():
    picks, flips, dist, padded = set(), set(), set(), False
    for s, (H0, W0) in enumerate(SIZES):
        p = R.draw_params(11, s, H0, W0)
        _, seg = R.synthetic_sample(11, s, H0, W0, block=BLOCKS[s % len(BLOCKS)])
        picks.add(R.choose_crop(R.resize_nearest(seg, p['new_h'], p['new_w']), p))
        flips.add(p['flip'])
        dist |= {k for k in ('brightness', 'contrast', 'saturation', 'hue') if p[k] is not None}
        padded |= p['crop_h'] < 512 or p['crop_w'] < 1024
    assert len(picks) >= 3 and 10 in picks and 0 in picks, picks
    assert flips == {True, False} and dist == {'brightness', 'contrast', 'saturation', 'hue'} and padded


@pytest.mark.gpu
def test_gpu_pipeline_matches_the_restatement_bit_exactly(gs):
    from gaia_seg_b200.data_pipeline import GpuTrainPipeline
    gs._lib.require_device()
    seed = 11
    pipe = GpuTrainPipeline(seed=seed)
    sizes = SIZES
    imgs, segs, ids = [], [], list(range(len(sizes)))
    for s, (H0, W0) in zip(ids, sizes):
        img, seg = R.synthetic_sample(seed, s, H0, W0, block=BLOCKS[s % len(BLOCKS)])
        imgs.append(torch.from_numpy(img))
        segs.append(torch.from_numpy(seg))
    out = pipe(imgs, segs, ids)
    torch.cuda.synchronize()
    choice = out['crop_choice'].cpu().tolist()
    got_img, got_lab = out['img'].cpu().numpy(), out['gt_semantic_seg'].cpu().numpy()
    assert got_img.shape == (len(ids), 3, 512, 1024) and got_lab.dtype == np.int64
    picks, flips, photometric = set(), set(), set()
    for n, s in enumerate(ids):
        p = R.draw_params(seed, s, *sizes[n])
        ref_img, ref_lab, t, _ = R.pipeline(imgs[n].numpy(), segs[n].numpy(), p)
        assert choice[n] == t, (n, choice[n], t)                                   # crop box: exact
        assert out['img_metas'][n]['flip'] == p['flip']
        assert (got_lab[n] == ref_lab).all(), (n, int((got_lab[n] != ref_lab).sum()))   # labels incl. 255 padding: exact
        d = np.abs(got_img[n] - ref_img)
        assert d.max() == 0.0, (n, float(d.max()), int((d > 0).sum()))           # image: bit-identical fp32
        picks.add(t); flips.add(p['flip'])
        photometric |= {k for k in ('brightness', 'contrast', 'saturation', 'hue') if p[k] is not None}
    assert len(picks) >= 3
    assert flips == {True, False} and photometric == {'brightness', 'contrast', 'saturation', 'hue'}
    # the train step consumes the batch as it is
    assert out['img'].is_cuda and out['gt_semantic_seg'].shape == (len(ids), 1, 512, 1024)
