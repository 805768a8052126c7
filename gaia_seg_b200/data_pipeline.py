"""GPU training data pipeline: the mmseg train pipeline of the reference config
(configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:60-75) as device kernels.

    Resize(img_scale=(2048, 1024), ratio_range=(0.5, 2.0)) -> RandomCrop((512, 1024), cat_max_ratio=0.75) ->
    RandomFlip(0.5) -> PhotoMetricDistortion() -> Normalize(mean, std, to_rgb) -> Pad((512, 1024), 0 / 255) ->
    DefaultFormatBundle -> Collect(img, gt_semantic_seg)

The reference runs these transforms in DataLoader worker processes on the host (2 per GPU, `workers_per_gpu=2`), which
cannot feed a 10-50 ms training step.  Here the decoded uint8 image and label map are copied to the device as they are
(8 MB per Cityscapes sample instead of 25 MB of fp32) and ONE fused kernel per sample (csrc/gs_data.cu) writes the
normalised fp32 NCHW crop and the int64 label map the model consumes; the crop re-draw rule is evaluated on the device.

Host logic in this file: the per-sample random decisions.  Deviation from the reference, stated: mmseg draws from numpy's
global stream, whose consumption depends on the data (RandomCrop re-draws); every sample here owns a counter-based
splitmix64 stream keyed by (seed, sample index) and ALL its draws are made up front -- reproducible across workers / ranks
and independent of the device's crop choice.
"""
import ctypes

import torch

from . import _lib
from ._lib import AugParams, GsError, call

MASK64 = (1 << 64) - 1
N_CANDIDATES = 11


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


class _Stream:
    def __init__(self, seed, sample):
        self.base = _splitmix64(seed & MASK64) ^ _splitmix64((sample + 0x51ED27) & MASK64)
        self.k = 0

    def uniform(self, lo=0.0, hi=1.0):
        z = _splitmix64((self.base + self.k) & MASK64)
        self.k += 1
        return lo + (hi - lo) * ((z >> 11) * (1.0 / (1 << 53)))

    def randint(self, lo, hi):
        return lo + int(self.uniform() * (hi - lo)) if hi > lo else lo


def _rescale_size(h, w, scale):
    long_e, short_e = max(scale), min(scale)
    sf = min(long_e / max(h, w), short_e / min(h, w))
    return int(h * sf + 0.5), int(w * sf + 0.5)


class GpuTrainPipeline:
    """`pipeline(imgs, segs, sample_ids)` -> dict(img fp32 [N, 3, h, w], gt_semantic_seg int64 [N, 1, h, w], img_metas).

    imgs / segs: per-sample uint8 tensors [H0, W0, 3] (BGR, as cv2.imread / mmcv LoadImageFromFile decode them) and
    [H0, W0]; host tensors are copied to the device first (pinned memory -> non_blocking)."""

    def __init__(self, seed=0, crop_size=(512, 1024), img_scale=(2048, 1024), ratio_range=(0.5, 2.0), cat_max_ratio=0.75,
                 flip_ratio=0.5, mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), ignore_index=255,
                 brightness_delta=32, contrast_range=(0.5, 1.5), saturation_range=(0.5, 1.5), hue_delta=18,
                 photometric=True):
        self.seed, self.crop_size, self.img_scale, self.ratio_range = seed, tuple(crop_size), tuple(img_scale), ratio_range
        self.cat_max_ratio, self.flip_ratio, self.ignore_index = cat_max_ratio, flip_ratio, ignore_index
        self.mean, self.std = tuple(mean), tuple(std)
        self.brightness_delta, self.contrast_range = brightness_delta, contrast_range
        self.saturation_range, self.hue_delta, self.photometric = saturation_range, hue_delta, photometric
        self._ws = None

    # ---- host logic: every random decision of one sample, in the order the transforms consume them ----
    def draw(self, sample, H0, W0):
        rs = _Stream(self.seed, sample)
        ch_max, cw_max = self.crop_size
        ratio = rs.uniform(*self.ratio_range)
        scale = (int(self.img_scale[0] * ratio), int(self.img_scale[1] * ratio))
        new_h, new_w = _rescale_size(H0, W0, scale)
        p = AugParams()
        p.H0, p.W0, p.new_h, p.new_w = H0, W0, new_h, new_w
        p.crop_h, p.crop_w = min(ch_max, new_h), min(cw_max, new_w)
        p.out_h, p.out_w = ch_max, cw_max
        for t in range(N_CANDIDATES):
            p.box_y[t] = rs.randint(0, max(new_h - ch_max, 0) + 1)
            p.box_x[t] = rs.randint(0, max(new_w - cw_max, 0) + 1)
        p.flip = 1 if rs.uniform() < self.flip_ratio else 0
        bd, (c0, c1), (s0, s1), hd = self.brightness_delta, self.contrast_range, self.saturation_range, self.hue_delta
        # the draws are always consumed (the streams of two pipelines with / without photometric stay aligned)
        hb = rs.randint(0, 2); vb = rs.uniform(-float(bd), float(bd)) if hb else 0.0
        mode = rs.randint(0, 2)
        hc = rs.randint(0, 2); vc = rs.uniform(c0, c1) if hc else 1.0
        hs = rs.randint(0, 2); vs = rs.uniform(s0, s1) if hs else 1.0
        hh = rs.randint(0, 2); vh = rs.randint(-hd, hd) if hh else 0
        on = 1 if self.photometric else 0
        p.has_brightness, p.brightness = hb * on, vb
        p.contrast_first, p.has_contrast, p.contrast = mode, hc * on, vc
        p.has_saturation, p.saturation = hs * on, vs
        p.has_hue, p.hue = hh * on, vh
        for i in range(3):
            p.mean[i] = self.mean[i]
            p.inv_std[i] = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(self.std[i], dtype=torch.float32))
        p.cat_max_ratio, p.ignore_index = self.cat_max_ratio, self.ignore_index
        return p

    def __call__(self, imgs, segs, sample_ids, device=None):
        _lib.require_device()
        if not (len(imgs) == len(segs) == len(sample_ids)):
            raise GsError('GpuTrainPipeline: imgs, segs and sample_ids must have the same length')
        device = device or torch.device('cuda', torch.cuda.current_device())
        N = len(imgs)
        h, w = self.crop_size
        out_img = torch.empty((N, 3, h, w), dtype=torch.float32, device=device)
        out_lab = torch.empty((N, 1, h, w), dtype=torch.int64, device=device)
        if self._ws is None or self._ws.device != device:
            self._ws = torch.empty(int(_lib.load().gs_aug_workspace_bytes()) * max(N, 8), dtype=torch.uint8, device=device)
            self._chosen = torch.zeros(max(N, 8), dtype=torch.int32, device=device)
        if N > self._chosen.numel():
            self._ws = torch.empty(int(_lib.load().gs_aug_workspace_bytes()) * N, dtype=torch.uint8, device=device)
            self._chosen = torch.zeros(N, dtype=torch.int32, device=device)
        ws_each = int(_lib.load().gs_aug_workspace_bytes())
        st = torch.cuda.current_stream(device).cuda_stream
        metas, keep = [], []
        for n, (img, seg, sid) in enumerate(zip(imgs, segs, sample_ids)):
            if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or seg.dtype != torch.uint8 or \
                    tuple(seg.shape) != tuple(img.shape[:2]):
                raise GsError('GpuTrainPipeline: expected uint8 image [H, W, 3] (BGR) and uint8 label map [H, W]')
            img = img.to(device, non_blocking=True).contiguous()
            seg = seg.to(device, non_blocking=True).contiguous()
            keep += [img, seg]
            H0, W0 = int(img.shape[0]), int(img.shape[1])
            p = self.draw(int(sid), H0, W0)
            ws = self._ws[n * ws_each:(n + 1) * ws_each]
            chosen = self._chosen[n:n + 1]
            call('gs_aug_choose_crop', seg.data_ptr(), ctypes.byref(p), ws.data_ptr(), chosen.data_ptr(), st)
            call('gs_aug_fused', img.data_ptr(), seg.data_ptr(), ctypes.byref(p), chosen.data_ptr(), out_img[n].data_ptr(),
                 out_lab[n].data_ptr(), st)
            metas.append(dict(ori_shape=(H0, W0, 3), img_shape=(p.crop_h, p.crop_w, 3), pad_shape=(h, w, 3),
                              scale_factor=p.new_w / W0, flip=bool(p.flip), flip_direction='horizontal',
                              crop_candidates=[(p.box_y[t], p.box_x[t]) for t in range(N_CANDIDATES)],
                              img_norm_cfg=dict(mean=self.mean, std=self.std, to_rgb=True)))
        return dict(img=out_img, gt_semantic_seg=out_lab, img_metas=metas, crop_choice=self._chosen[:N])
