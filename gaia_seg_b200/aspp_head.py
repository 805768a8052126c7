"""DynamicASPPHead -- DeepLabV3 head for BASELINE config 3 ("max sub-net dynamic ResNet-101 + DeepLabV3 ASPP head").

The reference ships no ASPP head; this is [EXT] mmseg `ASPPHead` / `ASPPModule`
(mmseg/models/decode_heads/aspp_head.py, the release line the reference pins) rebuilt from the reference's own dynamic
building blocks the way `DynamicPSPHead` rebuilds `PSPHead` (gaiaseg/models/decode_heads/dynamic_psp_head.py:75-147):

    x -> [ image_pool: AdaptiveAvgPool(1) -> 1x1 DynamicConvModule -> bilinear up to h x w,
           aspp_modules[i]: 1x1 (dilation 1) or 3x3 (dilation d, padding d) DynamicConvModule ]
      -> cat -> 3x3 `bottleneck` DynamicConvModule -> dropout -> conv_seg

Every branch takes the channel-PREFIX slice `x.size(1)` of its max-width weight in place (TMA tensor map, no copy); the
branch outputs are `channels` wide whatever the backbone width, so the concatenation needs no channel record.  Same
parameter names as mmseg (`image_pool.1.conv.weight`, `aspp_modules.{i}.conv.weight`, `bottleneck.conv.weight`,
`conv_seg.{weight,bias}`).
"""
import torch
import torch.nn as nn

from . import functional as F_gs
from ._lib import call
from .core import DynamicConv2d, DynamicConvModule, DynamicMixin
from .heads import FCNHead, HEADS, build_loss
from .psp_head import AdaptiveAvgPoolFn


class ResizeCatFn(torch.autograd.Function):
    """Channel concatenation into ONE buffer where every input is either already at the output size (channel-slice
    copy) or a smaller map that is bilinearly up-sampled (align_corners=False) straight into its slice.  Backward: the
    gradient of a full-size input is a channel-slice VIEW of the incoming gradient, that of a resized input the
    adjoint (gather) of the resize."""

    @staticmethod
    def forward(ctx, size, *inputs):
        inputs = [F_gs.as_act(t) for t in inputs]
        H, W = size
        N = inputs[0].shape[0]
        Cs = [t.shape[1] for t in inputs]
        total = sum(Cs)
        out = F_gs.new_act(N, total, H, W, inputs[0].device)
        P, st, off = N * H * W, F_gs._stream(), 0
        for t, C in zip(inputs, Cs):
            dst = out[:, off:off + C]
            if tuple(t.shape[2:]) == (H, W):
                call('gs_copy_channels', t.data_ptr(), F_gs.act_ld(t), dst.data_ptr(), total, P, C, st)
            else:
                call('gs_upsample_bf16_fwd', t.data_ptr(), F_gs.act_ld(t), N, t.shape[2], t.shape[3], C, dst.data_ptr(),
                     total, H, W, st)
            off += C
        ctx.meta = (H, W, [tuple(t.shape) for t in inputs])
        return out

    @staticmethod
    def backward(ctx, d):
        H, W, shapes = ctx.meta
        d = F_gs.as_act(d)
        N, st, off, grads = d.shape[0], F_gs._stream(), 0, [None]
        for shp in shapes:
            C = shp[1]
            if tuple(shp[2:]) == (H, W):
                grads.append(d[:, off:off + C])
            else:
                db = F_gs.new_act(shp[0], C, shp[2], shp[3], d.device)
                call('gs_upsample_bf16_bwd', d[:, off:off + C].data_ptr(), F_gs.act_ld(d), N, H, W, C, db.data_ptr(), C,
                     shp[2], shp[3], st)
                grads.append(db)
            off += C
        return tuple(grads)


class DynamicASPPModule(nn.ModuleList):
    """[EXT] mmseg ASPPModule: one DynamicConvModule per dilation (1x1 for dilation 1, else 3x3 with padding = dilation)."""

    def __init__(self, dilations, in_channels, channels, conv_cfg, norm_cfg, act_cfg):
        super().__init__()
        self.dilations, self.in_channels, self.channels = dilations, in_channels, channels
        for d in dilations:
            self.append(DynamicConvModule(in_channels, channels, 1 if d == 1 else 3, dilation=d,
                                          padding=0 if d == 1 else d, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                          act_cfg=act_cfg))

    def forward(self, x):
        return [m(x) for m in self]


@HEADS.register_module()
class DynamicASPPHead(FCNHead, DynamicMixin):
    search_space = set()

    def __init__(self, in_channels, channels, num_classes, dilations=(1, 6, 12, 18), dropout_ratio=0.1, conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1, input_transform=None,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0), ignore_index=255,
                 sampler=None, align_corners=False):
        nn.Module.__init__(self)
        assert isinstance(dilations, (list, tuple))
        self._init_inputs(in_channels, in_index, input_transform)
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.loss_decode = build_loss(loss_decode)
        self.ignore_index, self.align_corners = ignore_index, align_corners
        if sampler is not None or align_corners:
            raise NotImplementedError('pixel samplers / align_corners=True are not used by the GAIA-seg configs')
        self.sampler = None
        self.dilations = tuple(dilations)
        self.conv_seg = DynamicConv2d(channels, num_classes, kernel_size=1, padding=0)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.fp16_enabled = False
        self.image_pool = nn.Sequential(nn.AdaptiveAvgPool2d(1),
                                        DynamicConvModule(self.in_channels, channels, 1, conv_cfg=conv_cfg,
                                                          norm_cfg=norm_cfg, act_cfg=act_cfg))
        self.aspp_modules = DynamicASPPModule(self.dilations, self.in_channels, channels, conv_cfg=conv_cfg,
                                              norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.bottleneck = DynamicConvModule((len(self.dilations) + 1) * channels, channels, 3, padding=1,
                                            conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg)

    def forward(self, inputs):
        x = self._transform_inputs(inputs)
        pooled = self.image_pool[1](AdaptiveAvgPoolFn.apply(x, 1))          # [N, channels, 1, 1]; resize fused in the cat
        aspp_outs = ResizeCatFn.apply(tuple(x.shape[2:]), pooled, *self.aspp_modules(x))
        return self.cls_seg(self.bottleneck(aspp_outs))

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg, **kwargs):
        return self.losses(self.forward(inputs), gt_semantic_seg)
