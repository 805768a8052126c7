"""Analytic complexity (MACs / parameters) of the currently applied sub-net -- the producer side of the model-space
table the reference builds with tools/count_flops.py:128-179 ([EXT] gaiavision.utils.get_model_complexity_info, a
hook-based counter that needs a forward pass per sub-net).

Nothing is executed: shapes are propagated through the module tree the way the forward pass wires it
(gaiaseg/models/backbones/dynamic_resnet.py:405-421, dynamic_res_layer.py:166-172, dynamic_fcn_head.py:128-135,
psp_head.py:228-241), using every dynamic module's `width_state` / `depth_state`.  Counting convention = mmcv's
flops counter (the one gaiavision's forks): convolution = multiply-accumulates (+ one add per output for a bias),
BatchNorm = 2 per element (affine), ReLU / pooling / resize = 1 per element; `params` counts the ACTIVE slices.
"""
import torch.nn as nn

from .core import DynamicBottleneck, DynamicConv2d, DynamicConvModule


class Counter:
    def __init__(self):
        self.flops, self.conv_macs, self.params = 0, 0, 0

    def conv(self, conv, shape, co=None):
        """shape = (C, H, W) of the input; returns the output shape."""
        C, H, W = shape
        kh, kw = conv.kernel_size
        s, p, d = conv.stride[0], conv.padding[0], conv.dilation[0]
        Ho = (H + 2 * p - d * (kh - 1) - 1) // s + 1
        Wo = (W + 2 * p - d * (kw - 1) - 1) // s + 1
        co = getattr(conv, 'width_state', conv.out_channels) if co is None else co
        macs = Ho * Wo * co * C * kh * kw // conv.groups
        self.conv_macs += macs
        self.flops += macs
        self.params += co * C * kh * kw // conv.groups
        if conv.bias is not None:
            self.flops += Ho * Wo * co
            self.params += co
        return (co, Ho, Wo)

    def norm(self, bn, shape):
        C, H, W = shape
        if bn is None:
            return shape
        affine = getattr(bn, 'affine', True)
        self.flops += C * H * W * (2 if affine else 1)
        if affine:
            self.params += 2 * C
        return shape

    def elementwise(self, shape):
        C, H, W = shape
        self.flops += C * H * W
        return shape

    def cba(self, mod, shape):
        """DynamicConvModule: conv -> norm -> act."""
        shape = self.conv(mod.conv, shape)
        if mod.with_norm:
            shape = self.norm(mod.norm, shape)
        if mod.with_activation:
            shape = self.elementwise(shape)
        return shape


def _bottleneck(cnt, blk, shape):
    out = cnt.conv(blk.conv1, shape)
    out = cnt.elementwise(cnt.norm(blk.norm1, out))
    out = cnt.conv(blk.conv2, out)
    out = cnt.elementwise(cnt.norm(blk.norm2, out))
    out = cnt.conv(blk.conv3, out)
    out = cnt.norm(blk.norm3, out)
    if blk.downsample is not None:
        idn = shape
        for m in blk.downsample:
            if isinstance(m, DynamicConv2d):
                idn = cnt.conv(m, idn)
            elif isinstance(m, nn.modules.batchnorm._BatchNorm):
                idn = cnt.norm(m, idn)
            elif isinstance(m, nn.AvgPool2d):
                raise NotImplementedError('avg_down is not on the GAIA-seg hot path')
        assert idn == out, (idn, out)
    return cnt.elementwise(out)          # residual add folded into the final ReLU count (mmcv counts modules only)


def backbone_complexity(cnt, bb, shape):
    """DynamicResNet.forward (dynamic_resnet.py:405-421) on a (C, H, W) input; returns the list of output shapes."""
    if bb.deep_stem:
        mods = list(bb.stem)
        for i in range(0, len(mods), 3):
            shape = cnt.conv(mods[i], shape)
            shape = cnt.elementwise(cnt.norm(mods[i + 1], shape))
    else:
        shape = cnt.conv(bb.conv1, shape)
        shape = cnt.elementwise(cnt.norm(bb.norm1, shape))
    C, H, W = shape
    cnt.flops += C * H * W                                     # max-pool 3x3 s2 p1: one per input element
    shape = (C, (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1)
    outs = []
    for i, name in enumerate(bb.res_layers):
        layer = getattr(bb, name)
        for b in range(layer.depth_state):
            assert isinstance(layer[b], DynamicBottleneck)
            shape = _bottleneck(cnt, layer[b], shape)
        if i in bb.out_indices:
            outs.append(shape)
    return outs


def _cls_seg(cnt, head, shape):
    return cnt.conv(head.conv_seg, shape, co=head.num_classes)


def head_complexity(cnt, head, feats):
    """DynamicFCNHead / DynamicPSPHead / DynamicASPPHead forward on the backbone feature shapes."""
    x = feats[head.in_index]
    name = type(head).__name__
    if name == 'DynamicFCNHead':
        out = x
        if head.num_convs > 0:
            for m in head.convs:
                out = cnt.cba(m, out)
        if head.concat_input:
            out = cnt.cba(head.conv_cat, (x[0] + out[0], out[1], out[2]))
        return _cls_seg(cnt, head, out)
    if name == 'DynamicPSPHead':
        total = (head.in_channels if head.channel_record_mode == 'segmented' else x[0])
        for s, ppm in zip(head.pool_scales, head.psp_modules):
            cnt.flops += x[0] * x[1] * x[2]                    # adaptive average pool: one per input element
            b = cnt.cba(ppm[1], (x[0], s, s))
            cnt.flops += b[0] * x[1] * x[2]                    # bilinear resize: one per output element
            total += b[0]
        # the segmented channel record reads the max-width layout; the zero gap costs no arithmetic in the reference
        eff_in = total - (head.in_channels - x[0]) if head.channel_record_mode == 'segmented' else total
        conv = head.bottleneck.conv
        out = cnt.conv(conv, (eff_in, x[1], x[2]))
        out = cnt.elementwise(cnt.norm(head.bottleneck.norm, out))
        return _cls_seg(cnt, head, out)
    if name == 'DynamicASPPHead':
        cnt.flops += x[0] * x[1] * x[2]
        b = cnt.cba(head.image_pool[1], (x[0], 1, 1))
        cnt.flops += b[0] * x[1] * x[2]
        total = b[0]
        for m in head.aspp_modules:
            total += cnt.cba(m, x)[0]
        out = cnt.cba(head.bottleneck, (total, x[1], x[2]))
        return _cls_seg(cnt, head, out)
    raise NotImplementedError(f'complexity of {name}')


def get_model_complexity_info(model, input_shape, only_backbone_flops=False, as_strings=False, **_ignored):
    """(flops, params) of `model` (a DynamicEncoderDecoder) with its CURRENT arch state for a (3, H, W) input;
    same call shape as the reference's tools/count_flops.py:147-149.  `flops` are multiply-accumulates in mmcv's
    convention; the decode head is included unless `only_backbone_flops` (the auxiliary head never runs at inference)."""
    module = model.module if hasattr(model, 'module') else model
    if len(input_shape) != 3:
        raise ValueError('input_shape must be (C, H, W)')
    cnt = Counter()
    feats = backbone_complexity(cnt, module.backbone, tuple(int(v) for v in input_shape))
    if not only_backbone_flops:
        out = head_complexity(cnt, module.decode_head, feats)
        C, H, W = out
        cnt.flops += C * input_shape[1] * input_shape[2]        # resize of the logits to the input size
    if as_strings:
        return f'{cnt.flops / 1e9:.2f} GFLOPs', f'{cnt.params / 1e6:.2f} M'
    return cnt.flops, cnt.params


def conv_macs(model, input_shape, only_backbone=False):
    """Convolution multiply-accumulates only (the figure SURVEY 8d quotes)."""
    module = model.module if hasattr(model, 'module') else model
    cnt = Counter()
    feats = backbone_complexity(cnt, module.backbone, tuple(int(v) for v in input_shape))
    if not only_backbone:
        head_complexity(cnt, module.decode_head, feats)
    return cnt.conv_macs
