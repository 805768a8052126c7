"""CPU restatement (numpy) of the reference's TRAINING DATA PIPELINE -- test infrastructure, not the product.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path is
gaia_seg_b200/data_pipeline.py + csrc/gs_data.cu (ONE fused CUDA kernel per sample).

What is restated (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:60-75; the transforms themselves are [EXT] mmseg
0.x `mmseg/datasets/pipelines/transforms.py` + mmcv image ops, not in the reference tree):
    Resize(img_scale=(2048, 1024), ratio_range=(0.5, 2.0), keep_ratio)   mmcv.imrescale == cv2.resize INTER_LINEAR (image),
                                                                         INTER_NEAREST (label map)
    RandomCrop(crop_size=(512, 1024), cat_max_ratio=0.75)                up to 10 re-draws until no class covers >= 75 % of
                                                                         the non-ignored pixels of the crop
    RandomFlip(flip_ratio=0.5)                                           horizontal
    PhotoMetricDistortion()                                              brightness +-32, contrast / saturation x[0.5, 1.5],
                                                                         hue +-18 on uint8 BGR via 8-bit HSV round trips
    Normalize(mean, std, to_rgb=True) -> Pad(size=(512, 1024), pad_val=0, seg_pad_val=255) -> CHW fp32 / int64 labels
Pinned against OpenCV 4.13 executed in the build container (tests/golden/make_pipeline_golden.py ->
tests/golden/pipeline_golden.npz): resize_linear_u8 / resize_nearest / bgr2hsv_u8 / hsv2bgr_u8 are checked bit-exactly
against cv2.resize / cv2.cvtColor outputs stored in the fixture.

Deviation from the reference (stated): randomness.  mmseg draws from numpy's global MT19937 stream, and RandomCrop's
data-dependent number of re-draws shifts every later draw -- not reproducible across DataLoader workers even upstream.
Here every sample owns a counter-based stream (splitmix64 of (seed, sample index, draw index)), ALL draws of a sample are
made up front (11 crop candidates included), so the GPU can pick the crop itself without a host round trip and the
integer outputs (crop box, flip flag, label map) are bit-exactly reproducible on both sides.
"""
import math

import numpy as np

MASK64 = (1 << 64) - 1

MEAN = (123.675, 116.28, 103.53)        # RGB order (pspnet_ar50to101v2_gsync.py:57-58)
STD = (58.395, 57.12, 57.375)
IMG_SCALE = (2048, 1024)                # (long edge, short edge) of mmcv.imrescale
RATIO_RANGE = (0.5, 2.0)
CROP_SIZE = (512, 1024)                 # (h, w)
CAT_MAX_RATIO = 0.75
IGNORE_INDEX = 255
N_CANDIDATES = 11                       # get_crop_bbox() once + up to 10 re-draws


# ------------------------------------------------------------------------------------------------
# counter-based randomness
# ------------------------------------------------------------------------------------------------
def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


class SampleStream:
    """u_k = splitmix64(splitmix64(seed) ^ splitmix64(sample) + k) >> 11, scaled to [0, 1)."""

    def __init__(self, seed, sample):
        self.base = splitmix64(seed & MASK64) ^ splitmix64((sample + 0x51ED27) & MASK64)
        self.k = 0

    def uniform(self, lo=0.0, hi=1.0):
        z = splitmix64((self.base + self.k) & MASK64)
        self.k += 1
        return lo + (hi - lo) * ((z >> 11) * (1.0 / (1 << 53)))

    def randint(self, lo, hi):
        """integer in [lo, hi)  (numpy.random.randint convention)"""
        return lo + int(self.uniform() * (hi - lo)) if hi > lo else lo


def rescale_size(h, w, scale):
    """mmcv.rescale_size for a (long, short) tuple: keep the aspect ratio inside the box."""
    long_e, short_e = max(scale), min(scale)
    sf = min(long_e / max(h, w), short_e / min(h, w))
    return int(h * sf + 0.5), int(w * sf + 0.5)


def draw_params(seed, sample, H0, W0, crop_size=CROP_SIZE, img_scale=IMG_SCALE, ratio_range=RATIO_RANGE):
    """Every random decision of one sample, in the order the transforms consume them."""
    rs = SampleStream(seed, sample)
    ratio = rs.uniform(*ratio_range)                                   # Resize.random_sample_ratio
    scale = (int(img_scale[0] * ratio), int(img_scale[1] * ratio))
    new_h, new_w = rescale_size(H0, W0, scale)
    ch, cw = min(crop_size[0], new_h), min(crop_size[1], new_w)
    boxes = []
    for _ in range(N_CANDIDATES):                                       # RandomCrop.get_crop_bbox
        oy = rs.randint(0, max(new_h - crop_size[0], 0) + 1)
        ox = rs.randint(0, max(new_w - crop_size[1], 0) + 1)
        boxes.append((oy, ox))
    flip = rs.uniform() < 0.5                                           # RandomFlip
    p = dict(ratio=ratio, new_h=new_h, new_w=new_w, crop_h=ch, crop_w=cw, boxes=boxes, flip=bool(flip))
    # PhotoMetricDistortion: brightness, mode, contrast, saturation, hue
    p['brightness'] = rs.uniform(-32.0, 32.0) if rs.randint(0, 2) else None
    p['contrast_first'] = bool(rs.randint(0, 2))
    p['contrast'] = rs.uniform(0.5, 1.5) if rs.randint(0, 2) else None
    p['saturation'] = rs.uniform(0.5, 1.5) if rs.randint(0, 2) else None
    p['hue'] = rs.randint(-18, 18) if rs.randint(0, 2) else None
    return p


# ------------------------------------------------------------------------------------------------
# cv2-compatible 8-bit primitives (pinned against OpenCV in tests/golden)
# ------------------------------------------------------------------------------------------------
def _linear_coeffs(dst, src):
    """cv2 INTER_LINEAR index / fixed-point weight tables of one axis (resize.cpp: fx in float, weights * 2048 rounded)."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    f[lo], s[lo] = 0.0, 0
    hi = s >= src - 1
    f[hi], s[hi] = 0.0, src - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)              # saturate_cast<short> = round half to even
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    s1 = np.minimum(s + 1, src - 1)
    return s, s1, a0, a1


def resize_linear_u8(img, new_h, new_w):
    """cv2.resize(img, (new_w, new_h), interpolation=INTER_LINEAR) for uint8 HxWxC."""
    H, W = img.shape[:2]
    if (new_h, new_w) == (H, W):
        return img.copy()
    if H == 2 * new_h and W == 2 * new_w:
        # cv::resize switches INTER_LINEAR to its "area fast" path for an exact 2x decimation: rounded 2x2 mean
        s4 = img.astype(np.int64).reshape(new_h, 2, new_w, 2, -1).sum(axis=(1, 3))
        return ((s4 + 2) >> 2).astype(np.uint8).reshape(new_h, new_w, *img.shape[2:])
    sx0, sx1, a0, a1 = _linear_coeffs(new_w, W)
    sy0, sy1, b0, b1 = _linear_coeffs(new_h, H)
    src = img.astype(np.int64)
    rows = src[:, sx0] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # horizontal pass, scale 2^11
    r0, r1 = rows[sy0], rows[sy1]
    out = ((((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2)
    return np.clip(out, 0, 255).astype(np.uint8)


def nearest_index(dst, src):
    """cv2 INTER_NEAREST source index of every destination index: min(floor(d * src / dst), src - 1)."""
    ifx = 1.0 / (dst / src)
    return np.minimum(np.floor(np.arange(dst, dtype=np.float64) * ifx).astype(np.int64), src - 1)


def resize_nearest(seg, new_h, new_w):
    H, W = seg.shape[:2]
    return seg[nearest_index(new_h, H)][:, nearest_index(new_w, W)]


_SDIV = np.zeros(256, np.int64)
_HDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    _SDIV[_i] = int(np.rint((255 << 12) / (1.0 * _i)))
    _HDIV[_i] = int(np.rint((180 << 12) / (6.0 * _i)))


def bgr2hsv_u8(img):
    """cv2.cvtColor(img, COLOR_BGR2HSV) for uint8 (H in [0, 180)): the fixed-point RGB2HSV_b of OpenCV's color_hsv."""
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    s = (diff * _SDIV[v] + (1 << 11)) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << 11)) >> 12
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]], np.int64)


def hsv2bgr_u8(hsv):
    """cv2.cvtColor(hsv, COLOR_HSV2BGR) for uint8: OpenCV's float HSV2RGB formulas on (h, s / 255, v / 255) in fp32,
    result * 255 TRUNCATED.  (OpenCV's vectorised row body truncates, its scalar row tail -- the last < 16 pixels of a row --
    rounds: the library is not bit-consistent with itself here.  Truncation is what ~99 % of the pixels of a wide image
    get; the golden fixture pins the agreement with cv2 4.13: > 99.99 % of the values exact on the vector part, never
    more than 1 LSB apart anywhere.)"""
    f32 = np.float32
    h = hsv[..., 0].astype(f32)
    s = hsv[..., 1].astype(f32) * f32(1.0 / 255.0)
    v = hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    hh = h * f32(6.0 / 180.0)
    hh = np.where(hh >= 6, hh - 6, hh).astype(f32)
    sector = np.floor(hh).astype(np.int64)
    frac = (hh - sector.astype(f32)).astype(f32)
    bad = (sector < 0) | (sector >= 6)
    sector = np.where(bad, 0, sector)
    frac = np.where(bad, f32(0), frac).astype(f32)
    one = f32(1.0)
    tab = np.stack([v, (v * (one - s)).astype(f32), (v * (one - (s * frac).astype(f32))).astype(f32),
                    (v * (one - (s * (one - frac)).astype(f32))).astype(f32)], axis=-1)
    idx = _SECTOR[sector]                                            # [..., 3] -> b, g, r
    bgr = np.take_along_axis(tab, idx, axis=-1)
    gray = (hsv[..., 1] == 0)[..., None]
    bgr = np.where(gray, v[..., None], bgr).astype(f32)
    return np.clip(np.floor((bgr * f32(255.0)).astype(f32)), 0, 255).astype(np.uint8)


def _convert(img, alpha=1.0, beta=0.0):
    """PhotoMetricDistortion.convert: clip(img * alpha + beta, 0, 255).astype(uint8)  (fp32 arithmetic, truncation)."""
    x = img.astype(np.float32) * np.float32(alpha) + np.float32(beta)
    return np.clip(x, 0, 255).astype(np.uint8)


def photometric(img, p):
    """PhotoMetricDistortion.__call__ on a uint8 BGR image with the pre-drawn decisions of `p`."""
    if p['brightness'] is not None:
        img = _convert(img, beta=p['brightness'])
    if p['contrast_first'] and p['contrast'] is not None:
        img = _convert(img, alpha=p['contrast'])
    if p['saturation'] is not None:
        hsv = bgr2hsv_u8(img)
        hsv[..., 1] = _convert(hsv[..., 1], alpha=p['saturation'])
        img = hsv2bgr_u8(hsv)
    if p['hue'] is not None:
        hsv = bgr2hsv_u8(img)
        hsv[..., 0] = ((hsv[..., 0].astype(np.int64) + p['hue']) % 180).astype(np.uint8)
        img = hsv2bgr_u8(hsv)
    if (not p['contrast_first']) and p['contrast'] is not None:
        img = _convert(img, alpha=p['contrast'])
    return img


def choose_crop(seg_resized, p, crop_size=CROP_SIZE, cat_max_ratio=CAT_MAX_RATIO, ignore_index=IGNORE_INDEX):
    """RandomCrop.__call__'s re-draw loop over the pre-drawn candidates: index of the box that is used."""
    ch, cw = p['crop_h'], p['crop_w']
    if cat_max_ratio >= 1.0:
        return 0
    for t in range(N_CANDIDATES - 1):
        oy, ox = p['boxes'][t]
        cnt = np.bincount(seg_resized[oy:oy + ch, ox:ox + cw].ravel(), minlength=256)
        cnt[ignore_index] = 0
        cnt = cnt[cnt > 0]
        if len(cnt) > 1 and cnt.max() / cnt.sum() < cat_max_ratio:
            return t
    return N_CANDIDATES - 1


def pipeline(img_bgr_u8, seg_u8, p, crop_size=CROP_SIZE, mean=MEAN, std=STD):
    """The whole train pipeline for one sample.  Returns (img fp32 [3, 512, 1024] RGB-normalised, labels int64
    [1, 512, 1024], chosen candidate index, the uint8 BGR image after PhotoMetricDistortion (for tolerance accounting))."""
    img = resize_linear_u8(img_bgr_u8, p['new_h'], p['new_w'])
    seg = resize_nearest(seg_u8, p['new_h'], p['new_w'])
    t = choose_crop(seg, p, crop_size)
    oy, ox = p['boxes'][t]
    ch, cw = p['crop_h'], p['crop_w']
    img, seg = img[oy:oy + ch, ox:ox + cw], seg[oy:oy + ch, ox:ox + cw]
    if p['flip']:
        img, seg = img[:, ::-1], seg[:, ::-1]
    img = photometric(np.ascontiguousarray(img), p)
    rgb = img[..., ::-1].astype(np.float32)                              # to_rgb
    norm = (rgb - np.array(mean, np.float32)) * (np.float32(1.0) / np.array(std, np.float32))   # mmcv.imnormalize: * 1/std
    out = np.zeros((3, crop_size[0], crop_size[1]), np.float32)          # Pad(pad_val=0) AFTER Normalize
    out[:, :ch, :cw] = norm.transpose(2, 0, 1)
    lab = np.full((1, crop_size[0], crop_size[1]), 255, np.int64)        # seg_pad_val=255
    lab[0, :ch, :cw] = seg
    return out, lab, t, img


def synthetic_sample(seed, sample, H0=1024, W0=2048, num_classes=19, block=(128, 256)):
    """A Cityscapes-shaped uint8 BGR image and label map with uniform regions of `block` pixels (large blocks make
    RandomCrop's cat_max_ratio re-draws actually happen): blocky class layout + noise, 5 % ignore pixels."""
    rng = np.random.default_rng(seed * 1000003 + sample)
    bh, bw = block
    coarse = rng.integers(0, num_classes, (H0 // bh + 1, W0 // bw + 1))
    seg = np.kron(coarse, np.ones((bh, bw), np.int64))[:H0, :W0].astype(np.uint8)
    seg[rng.random((H0, W0)) < 0.05] = 255
    base = (seg.astype(np.int64)[..., None] * np.array([9, 5, 13]) + np.array([20, 60, 100])) % 256
    img = np.clip(base + rng.integers(-40, 40, (H0, W0, 3)), 0, 255).astype(np.uint8)
    return img, seg
