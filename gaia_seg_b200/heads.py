"""Decode heads: DynamicFCNHead (gaiaseg/models/decode_heads/dynamic_fcn_head.py:23-231 on top of the in-tree
base gaiaseg/models/decode_heads/fcn_head.py:139-275).

Same constructor arguments, attribute / parameter names (`convs.{i}.conv.weight`, `conv_cat.conv.weight`,
`conv_seg.{weight,bias}`), `forward`, `losses`, `forward_train`, `forward_test`, `cls_seg`.  `losses()` keeps the
reference semantics -- bilinear resize of the logits to the label size (align_corners), pixel-wise CE with
ignore_index averaged over ALL pixels times loss_weight, and top-1 accuracy over all pixels -- but runs them
as ONE fused kernel pair on the low-resolution logits (gaia_seg_b200.functional.upsample_ce).
"""
import torch
import torch.nn as nn

from . import functional as F_gs
from .backbone import normal_init
from .core import DynamicConv2d, DynamicConvModule, DynamicMixin, Registry

HEADS = Registry('head')
LOSSES = Registry('loss')


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    """[EXT] mmseg CrossEntropyLoss for the soft-max case (restated in-tree at
    gaiaseg/models/losses/cross_entropy_loss.py:67-94 + utils.py:26-55): F.cross_entropy(reduction='none',
    ignore_index) -> mean over all elements -> * loss_weight.  Holds the hyper-parameters; the arithmetic is
    the fused upsample+CE kernel called by the head."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, loss_weight=1.0):
        super().__init__()
        if use_sigmoid or use_mask:
            raise NotImplementedError('sigmoid / mask CE variants are not on the GAIA-seg hot path')
        if class_weight is not None:
            raise NotImplementedError('class_weight is not used by any GAIA-seg config')
        if reduction != 'mean':
            raise NotImplementedError("only reduction='mean' (the reference's setting)")
        self.use_sigmoid, self.use_mask, self.reduction = use_sigmoid, use_mask, reduction
        self.class_weight, self.loss_weight = class_weight, loss_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=255,
                **kwargs):
        if weight is not None or avg_factor is not None or reduction_override not in (None, 'mean'):
            raise NotImplementedError('per-pixel weights / avg_factor are not on the GAIA-seg hot path')
        return F_gs.upsample_ce(cls_score, label, ignore_index, self.loss_weight)[0]


def build_loss(cfg):
    from .core import build_from_cfg
    return build_from_cfg(cfg, LOSSES)


class FCNHead(nn.Module):
    """Helpers of the reference's base head (fcn_head.py:139-253)."""

    def _init_inputs(self, in_channels, in_index, input_transform):
        if input_transform is not None:
            raise NotImplementedError('input_transform is unused by the dynamic heads')
        assert isinstance(in_channels, int) and isinstance(in_index, int)
        self.input_transform, self.in_index, self.in_channels = input_transform, in_index, in_channels

    def init_weights(self):
        normal_init(self.conv_seg, mean=0, std=0.01)

    def _transform_inputs(self, inputs):
        return inputs[self.in_index]

    def cls_seg(self, feat):
        if self.dropout is not None:
            feat = F_gs.dropout2d(feat, self.dropout_ratio, self.training)
        return F_gs.conv_seg(feat, self.conv_seg)

    def forward_test(self, inputs, img_metas, test_cfg):
        return self.forward(inputs)

    def losses(self, seg_logit, seg_label):
        """loss_seg / acc_seg from the LOW-RES logits (dynamic_fcn_head.py:137-159)."""
        if self.sampler is not None:
            raise NotImplementedError('pixel samplers (OHEM) are not used by the GAIA-seg configs')
        if self.align_corners:
            raise NotImplementedError('align_corners=True is not used by the GAIA-seg configs')
        loss_seg, acc_seg, _ = F_gs.upsample_ce(seg_logit, seg_label, self.ignore_index, self.loss_decode.loss_weight)
        return dict(loss_seg=loss_seg, acc_seg=acc_seg)


@HEADS.register_module()
class DynamicFCNHead(FCNHead, DynamicMixin):
    search_space = set()

    def __init__(self, in_channels, channels, num_classes, num_convs=2, kernel_size=3, concat_input=True,
                 dropout_ratio=0.1, conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1,
                 input_transform=None, loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                 ignore_index=255, sampler=None, align_corners=False):
        nn.Module.__init__(self)
        self._init_inputs(in_channels, in_index, input_transform)
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.loss_decode = build_loss(loss_decode)
        self.ignore_index, self.align_corners = ignore_index, align_corners
        if sampler is not None:
            raise NotImplementedError('pixel samplers (OHEM) are not used by the GAIA-seg configs')
        self.sampler = None
        self.conv_seg = DynamicConv2d(channels, num_classes, kernel_size=1, padding=0)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.fp16_enabled = False
        assert num_convs >= 0
        self.num_convs, self.concat_input, self.kernel_size = num_convs, concat_input, kernel_size
        if num_convs == 0:
            assert self.in_channels == self.channels
        convs = []
        for i in range(num_convs):
            convs.append(DynamicConvModule(self.in_channels if i == 0 else self.channels, self.channels,
                                           kernel_size=kernel_size, padding=kernel_size // 2, conv_cfg=self.conv_cfg,
                                           norm_cfg=self.norm_cfg, act_cfg=self.act_cfg))
        self.convs = nn.Identity() if num_convs == 0 else nn.Sequential(*convs)
        if self.concat_input:
            self.conv_cat = DynamicConvModule(self.in_channels + self.channels, self.channels, kernel_size=kernel_size,
                                              padding=kernel_size // 2, conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg,
                                              act_cfg=self.act_cfg)

    def forward(self, inputs):
        x = self._transform_inputs(inputs)
        carrier = None
        if (self.concat_input and self.num_convs > 0 and F_gs.SKIP_GRAD_CARRIER and torch.is_grad_enabled()
                and x.requires_grad):
            # x feeds convs[0] AND the concat: the concat's share of dL/dx is added inside convs[0]'s dgrad epilogue
            carrier = F_gs.GradCarrier()
            output = self.convs[0](x, grad_carrier=carrier)
            for m in list(self.convs)[1:]:
                output = m(output)
        else:
            output = self.convs(x)
        if self.concat_input:
            output = self.conv_cat(F_gs.cat_channels([x, output], skip_carrier=carrier))
        return self.cls_seg(output)

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg, **kwargs):
        if kwargs.get('aux_teacher_logits') is not None:
            raise NotImplementedError('in-place distillation (dynamic_fcn_head.py:186-226) is outside the hot path')
        seg_logits = self.forward(inputs)
        return self.losses(seg_logits, gt_semantic_seg)
