"""DynamicPSPHead / DynamicPPM -- mirror of gaiaseg/models/decode_heads/dynamic_psp_head.py:25-173 on top of the in-tree
base gaiaseg/models/decode_heads/psp_head.py:228-241 (the head of the reference's only in-tree seg training config,
configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:24-37).

    x -> for s in pool_scales: AdaptiveAvgPool(s) -> 1x1 DynamicConvModule(-> channels) -> bilinear up to h x w
      -> cat[x, branches] -> 3x3 `bottleneck` DynamicConvModule called with channel_record -> dropout -> conv_seg

`channel_record = [C_x, channels, ...]` (psp_head.py:235-239): the gaiavision fork's meaning is not in the tree; the
self-consistent reading is a SEGMENTED input slice -- the feature keeps weight columns [0:C_x] and the
pyramid branches keep their MAX-width columns [in_channels : in_channels + 4*channels] when the backbone narrows.  It is
realised without any gather: the concat buffer always has the max-width layout and the gap [C_x : in_channels] is
zero-filled, so the plain prefix convolution over it IS the segmented convolution (zeros contribute nothing forward,
get zero weight gradient backward).  `channel_record_mode='prefix'` selects the other reading (plain prefix slice).
"""
import torch
import torch.nn as nn

from . import functional as F_gs
from ._lib import call
from .core import DynamicConv2d, DynamicConvModule, DynamicMixin
from .heads import FCNHead, HEADS, build_loss


class AdaptiveAvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, S):
        x = F_gs.as_act(x)
        N, C, H, W = x.shape
        y = F_gs.new_act(N, C, S, S, x.device)
        call('gs_adaptive_avgpool_fwd', x.data_ptr(), N, H, W, C, F_gs.act_ld(x), S, y.data_ptr(), C, F_gs._stream())
        ctx.dims = (N, C, H, W, S)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W, S = ctx.dims
        dy = F_gs.as_act(dy)
        dx = F_gs.new_act(N, C, H, W, dy.device)
        call('gs_adaptive_avgpool_bwd', dy.data_ptr(), F_gs.act_ld(dy), N, H, W, C, S, dx.data_ptr(), C, 0, F_gs._stream())
        return dx, None


class PSPCatFn(torch.autograd.Function):
    """cat[x, upsample(b_1), ..., upsample(b_k)] written straight into ONE buffer: x is copied to channels [0:C_x], the
    gap up to `x_slot` is zero-filled (segmented mode), every small branch map is bilinearly up-sampled directly into
    its channel slice.  Backward: dx is a channel-slice VIEW of the incoming gradient, the branch gradients are the
    adjoint (gather) of the resize."""

    @staticmethod
    def forward(ctx, x_slot, x, *branches):
        x = F_gs.as_act(x)
        branches = [F_gs.as_act(b) for b in branches]
        N, Cx, H, W = x.shape
        Cb = [b.shape[1] for b in branches]
        total = x_slot + sum(Cb)
        out = F_gs.new_act(N, total, H, W, x.device)
        P, st = N * H * W, F_gs._stream()
        call('gs_copy_channels', x.data_ptr(), F_gs.act_ld(x), out.data_ptr(), total, P, Cx, st)
        if x_slot > Cx:
            call('gs_zero_channels', out[:, Cx:x_slot].data_ptr(), total, P, x_slot - Cx, st)
        off = x_slot
        for b, C in zip(branches, Cb):
            call('gs_upsample_bf16_fwd', b.data_ptr(), F_gs.act_ld(b), N, b.shape[2], b.shape[3], C,
                 out[:, off:off + C].data_ptr(), total, H, W, st)
            off += C
        ctx.meta = (x_slot, Cx, Cb, [tuple(b.shape) for b in branches])
        return out

    @staticmethod
    def backward(ctx, d):
        x_slot, Cx, Cb, shapes = ctx.meta
        d = F_gs.as_act(d)
        N, _, H, W = d.shape
        grads, off, st = [None, d[:, :Cx]], x_slot, F_gs._stream()
        for C, shp in zip(Cb, shapes):
            db = F_gs.new_act(shp[0], C, shp[2], shp[3], d.device)
            call('gs_upsample_bf16_bwd', d[:, off:off + C].data_ptr(), F_gs.act_ld(d), N, H, W, C, db.data_ptr(), C,
                 shp[2], shp[3], st)
            grads.append(db)
            off += C
        return tuple(grads)


class DynamicPPM(nn.ModuleList):
    """Pooling Pyramid Module (dynamic_psp_head.py:25-73): per scale nn.Sequential(AdaptiveAvgPool2d, DynamicConvModule)."""

    def __init__(self, pool_scales, in_channels, channels, conv_cfg, norm_cfg, act_cfg, align_corners):
        super().__init__()
        self.pool_scales, self.align_corners = pool_scales, align_corners
        self.in_channels, self.channels = in_channels, channels
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        for s in pool_scales:
            self.append(nn.Sequential(nn.AdaptiveAvgPool2d(s),
                                      DynamicConvModule(in_channels, channels, 1, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                                        act_cfg=act_cfg)))

    def forward(self, x):
        """Returns the LOW-RESOLUTION branch outputs [N, channels, s, s]; the resize to x's size is fused into the
        concat (PSPCatFn) instead of materialising four full-resolution maps (dynamic_psp_head.py:62-73)."""
        outs = []
        for s, ppm in zip(self.pool_scales, self):
            pooled = AdaptiveAvgPoolFn.apply(x, s)
            outs.append(ppm[1](pooled))
        return outs


@HEADS.register_module()
class DynamicPSPHead(FCNHead, DynamicMixin):
    search_space = set()

    def __init__(self, in_channels, channels, num_classes, pool_scales=(1, 2, 3, 6), dropout_ratio=0.1, conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1, input_transform=None,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0), ignore_index=255,
                 sampler=None, align_corners=False, channel_record_mode='segmented'):
        nn.Module.__init__(self)
        assert channel_record_mode in ('segmented', 'prefix')
        self._init_inputs(in_channels, in_index, input_transform)
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.loss_decode = build_loss(loss_decode)
        self.ignore_index, self.align_corners = ignore_index, align_corners
        if sampler is not None or align_corners:
            raise NotImplementedError('pixel samplers / align_corners=True are not used by the GAIA-seg configs')
        self.sampler = None
        self.channel_record_mode = channel_record_mode
        self.conv_seg = DynamicConv2d(channels, num_classes, kernel_size=1, padding=0)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.fp16_enabled = False
        self.pool_scales = pool_scales
        self.psp_modules = DynamicPPM(pool_scales, self.in_channels, channels, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                      act_cfg=act_cfg, align_corners=align_corners)
        self.bottleneck = DynamicConvModule(self.in_channels + len(pool_scales) * channels, channels, 3, padding=1,
                                            conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg)

    def forward(self, inputs):
        x = self._transform_inputs(inputs)
        branches = self.psp_modules(x)
        x_slot = self.in_channels if self.channel_record_mode == 'segmented' else x.size(1)
        psp_outs = PSPCatFn.apply(x_slot, x, *branches)
        output = self.bottleneck(psp_outs)
        return self.cls_seg(output)

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg, **kwargs):
        if kwargs.get('teacher_logits') is not None or kwargs.get('aux_teacher_logits') is not None:
            raise NotImplementedError('in-place distillation (dynamic_psp_head.py:176-245) is outside the hot path')
        return self.losses(self.forward(inputs), gt_semantic_seg)
