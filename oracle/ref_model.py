"""ORACLE -- CPU restatement of the reference's algorithm for the GAIA-seg hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product (gaia_seg_b200/) never does.  It is plain PyTorch on CPU in fp32 (fp64 on request): every op is the
library call the reference reaches (SURVEY.md 8c "Oracle definition"):

    DynamicConv2d      F.conv2d(x, W[:width_state, :x.size(1)], b[:width_state], stride, padding, dilation)
    DynamicBatchNorm   F.batch_norm(x, rm[:C], rv[:C], w[:C], b[:C], training, momentum 0.1, eps 1e-5)
    DynamicBottleneck  relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1 x))))))) + identity)
    heads              F.interpolate(bilinear, align_corners=False) -> F.cross_entropy(reduction='none',
                       ignore_index=255).mean() * loss_weight ; accuracy = topk(1) hits / numel * 100
    inference          resize -> softmax -> argmax

PARITY PINNING.  The arithmetic of DynamicConv2d / DynamicBatchNorm / DynamicBottleneck / DynamicConvModule lives
in `gaiavision` (unpinned, NOT vendored by the reference, not installable here: no network) and mmseg / mmcv
(absent); the reference ships no tests, golden vectors or fixtures.  Those classes are therefore restated from the
reference's call sites and are "parity unpinned".  What IS pinned against the reference's own code executed in
this container (tests/golden/make_golden.py, fixtures under tests/golden/):
  * cross_entropy / weight_reduce_loss / accuracy  <- gaiaseg/models/losses/{cross_entropy_loss,utils,accuracy}.py
  * DynamicResNet / DynamicResLayer / DynamicFCNHead wiring (module tree, parameter names, forward order,
    manipulate_stem / manipulate_body fan-out)  <- gaiaseg/models/backbones/dynamic_resnet.py,
    gaiaseg/models/utils/dynamic_res_layer.py, gaiaseg/models/decode_heads/dynamic_fcn_head.py, imported with
    stub mmcv / mmseg / gaiavision packages whose dynamic ops are THIS file's classes.

Each class cites the reference file:line it follows.
"""
from collections.abc import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.batchnorm import _BatchNorm


# ------------------------------------------------------------------------------------------------
# [EXT] gaiavision.core  (contracts: SURVEY.md 2.1)
# ------------------------------------------------------------------------------------------------
class DynamicMixin:
    """manipulate_arch routes key k to self.manipulate_<k> (dynamic_resnet.py:381-403, dynamic_encoder_decoder.py:31-42)."""
    search_space = set()

    def init_state(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, f'{k}_state', v)

    def manipulate_arch(self, arch_meta):
        for k, v in arch_meta.items():
            getattr(self, f'manipulate_{k}')(v)

    def deploy(self, mode=True):
        self._deploying = mode
        for m in self.children():
            _deploy(m, mode)


def _deploy(m, mode):
    if isinstance(m, DynamicMixin):
        m.deploy(mode)
    else:
        for c in m.children():
            _deploy(c, mode)


class DynamicConv2d(nn.Conv2d, DynamicMixin):
    """call sites: dynamic_fcn_head.py:76, dynamic_resnet.py:259-297 (via build_conv_layer 'DynConv2d')."""
    search_space = {'width'}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.init_state(width=self.out_channels)

    def manipulate_width(self, width):
        assert 0 < width <= self.out_channels
        self.width_state = width

    def forward(self, x, cols=None):
        co, ci = self.width_state, x.size(1)
        # `cols`: explicit input-channel columns (the segmented `channel_record` slice of the PSP bottleneck)
        w = self.weight[:co, :ci] if cols is None else self.weight[:co][:, cols]
        b = self.bias[:co] if self.bias is not None else None
        if getattr(self, '_deploying', False):
            self.weight = nn.Parameter(w.detach().clone())
            if b is not None:
                self.bias = nn.Parameter(b.detach().clone())
            self.out_channels, self.in_channels = co, ci
            w, b = self.weight, self.bias
        return F.conv2d(x, w, b, self.stride, self.padding, self.dilation, self.groups)


class DynamicBatchNorm2d(_BatchNorm, DynamicMixin):
    """'DynBN' / 'DynSyncBN' (dynamic_resnet.py:94,267,298; dynamic_res_layer.py:92).  SyncBN over R ranks is the
    same module applied to the concatenated R*N batch (equal per-rank batch sizes)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, **_ignored):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)

    def _check_input_dim(self, input):
        assert input.dim() == 4

    def forward(self, x):
        c = x.size(1)
        if getattr(self, '_deploying', False) and self.num_features != c:
            self.weight = nn.Parameter(self.weight.detach()[:c].clone())
            self.bias = nn.Parameter(self.bias.detach()[:c].clone())
            self.running_mean = self.running_mean[:c].clone()
            self.running_var = self.running_var[:c].clone()
            self.num_features = c
        use_batch = self.training or not self.track_running_stats or self.running_mean is None
        rm = self.running_mean[:c] if self.running_mean is not None and (not use_batch or self.training) else None
        rv = self.running_var[:c] if rm is not None else None
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked += 1
        w = self.weight[:c] if self.affine else None
        b = self.bias[:c] if self.affine else None
        return F.batch_norm(x, rm, rv, w, b, use_batch, self.momentum if self.momentum is not None else 0.0, self.eps)


def build_conv_layer(cfg, *args, **kwargs):
    return DynamicConv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.pop('group_size', None)
    layer = DynamicBatchNorm2d(num_features, **cfg)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return 'bn' + str(postfix), layer


class DynamicBottleneck(nn.Module, DynamicMixin):
    """constructed at dynamic_res_layer.py:106-125; norm3 at dynamic_resnet.py:362; width rule restated in-tree at
    elastic_convformer.py:334-341 (conv1, conv2 -> w; conv3, downsample -> 4w)."""
    expansion = 4
    search_space = {'width'}

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch', with_cp=False,
                 conv_cfg=None, norm_cfg=dict(type='DynBN'), dcn=None, plugins=None):
        super().__init__()
        s1, s2 = (1, stride) if style == 'pytorch' else (stride, 1)
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.norm3_name, norm3 = build_norm_layer(norm_cfg, planes * 4, postfix=3)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, kernel_size=1, stride=s1, bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, kernel_size=3, stride=s2, padding=dilation,
                                      dilation=dilation, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.conv3 = build_conv_layer(conv_cfg, planes, planes * 4, kernel_size=1, bias=False)
        self.add_module(self.norm3_name, norm3)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.init_state(width=planes)

    norm1 = property(lambda self: getattr(self, self.norm1_name))
    norm2 = property(lambda self: getattr(self, self.norm2_name))
    norm3 = property(lambda self: getattr(self, self.norm3_name))

    def manipulate_width(self, width):
        self.width_state = width
        self.conv1.manipulate_width(width)
        self.conv2.manipulate_width(width)
        self.conv3.manipulate_width(width * 4)
        if self.downsample is not None:
            for m in self.downsample:
                if isinstance(m, DynamicConv2d):
                    m.manipulate_width(width * 4)

    def forward(self, x):
        identity = x
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


class DynamicConvModule(nn.Module, DynamicMixin):
    """conv (bias iff no norm) -> norm -> act (dynamic_fcn_head.py:94-126)."""
    search_space = {'width'}

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), **_ignored):
        super().__init__()
        self.with_norm, self.with_activation = norm_cfg is not None, act_cfg is not None
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                     dilation=dilation, bias=not self.with_norm)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            self.activate = nn.ReLU(inplace=True)

    norm = property(lambda self: getattr(self, self.norm_name) if self.with_norm else None)

    def manipulate_width(self, width):
        self.conv.manipulate_width(width)

    def forward(self, x, channel_record=None):
        if channel_record is not None:
            # "segmented input slice" reading of the PSP bottleneck call (psp_head.py:235-239):
            # segment s of the input owns the weight columns starting at its MAX-width offset
            raise NotImplementedError('handled by DynamicPSPHead in this oracle')
        x = self.conv(x)
        if self.with_norm:
            x = self.norm(x)
        if self.with_activation:
            x = self.activate(x)
        return x


# ------------------------------------------------------------------------------------------------
# in-tree model code, restated
# ------------------------------------------------------------------------------------------------
class DynamicResLayer(nn.ModuleList, DynamicMixin):
    """gaiaseg/models/utils/dynamic_res_layer.py:16-172."""
    search_space = {'depth', 'width'}

    def __init__(self, block, inplanes, planes, depth, stride=1, dilation=1, avg_down=False, conv_cfg=None,
                 norm_cfg=None, downsample_first=True, contract_dilation=False, **kwargs):
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:           # :70
            mods, conv_stride = [], stride
            if avg_down:                                                    # :74-82
                conv_stride = 1
                mods.append(nn.AvgPool2d(kernel_size=stride, stride=stride, ceil_mode=True, count_include_pad=False))
            mods += [build_conv_layer(conv_cfg, inplanes, planes * block.expansion, kernel_size=1, padding=0,
                                      stride=conv_stride, bias=False),
                     build_norm_layer(norm_cfg, planes * block.expansion)[1]]
            downsample = nn.Sequential(*mods)
        first_dilation = dilation // 2 if (dilation > 1 and contract_dilation) else dilation   # :98-102
        layers = [block(inplanes=inplanes, planes=planes, stride=stride, dilation=first_dilation,
                        downsample=downsample, conv_cfg=conv_cfg, norm_cfg=norm_cfg, **kwargs)]
        for _ in range(1, depth):
            layers.append(block(inplanes=planes * block.expansion, planes=planes, stride=1, dilation=dilation,
                                conv_cfg=conv_cfg, norm_cfg=norm_cfg, **kwargs))
        super().__init__(layers)
        self.init_state(depth=depth, width=planes)

    def manipulate_depth(self, depth):                                      # :149-152
        assert depth >= 1
        self.depth_state = depth

    def manipulate_width(self, width):                                      # :154-157
        for m in self:
            m.manipulate_width(width)

    def forward(self, x):                                                   # :159-172
        if getattr(self, '_deploying', False):
            del self[self.depth_state:]
        for i in range(self.depth_state):
            x = self[i](x)
        return x


class DynamicResNet(nn.Module, DynamicMixin):
    """gaiaseg/models/backbones/dynamic_resnet.py:25-421."""
    search_space = {'stem', 'body'}

    def __init__(self, in_channels, stem_width, body_width, body_depth, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style='pytorch', deep_stem=False, avg_down=False,
                 conv_cfg=None, norm_cfg=dict(type='DynSyncBN'), norm_eval=False, zero_init_residual=True,
                 contract_dilation=False, **_ignored):
        super().__init__()
        self.deep_stem, self.out_indices, self.norm_eval = deep_stem, out_indices, norm_eval
        self.zero_init_residual = zero_init_residual
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        inplanes = stem_width[-1] if deep_stem else stem_width             # :139
        self._make_stem_layer(in_channels, stem_width)
        self.res_layers = []
        for i, num_blocks in enumerate(body_depth[:num_stages]):            # :147-173
            layer = DynamicResLayer(block=DynamicBottleneck, inplanes=inplanes, planes=body_width[i], depth=num_blocks,
                                    stride=strides[i], dilation=dilations[i], style=style, avg_down=avg_down,
                                    conv_cfg=conv_cfg, norm_cfg=norm_cfg, contract_dilation=contract_dilation)
            inplanes = body_width[i] * 4
            self.add_module(f'layer{i + 1}', layer)
            self.res_layers.append(f'layer{i + 1}')

    def _make_stem_layer(self, in_channels, stem_width):                    # :255-302
        if self.deep_stem:
            assert isinstance(stem_width, Sequence)
            ch = [in_channels] + list(stem_width)
            mods = []
            for i in range(3):
                mods += [build_conv_layer(self.conv_cfg, ch[i], ch[i + 1], kernel_size=3, stride=2 if i == 0 else 1,
                                          padding=1, bias=False),
                         build_norm_layer(self.norm_cfg, ch[i + 1])[1], nn.ReLU(inplace=True)]
            self.stem = nn.Sequential(*mods)
        else:
            self.conv1 = build_conv_layer(self.conv_cfg, in_channels, stem_width, kernel_size=7, stride=2, padding=3,
                                          bias=False)
            self.norm1_name, norm1 = build_norm_layer(self.norm_cfg, stem_width, postfix=1)
            self.add_module(self.norm1_name, norm1)
            self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

    norm1 = property(lambda self: getattr(self, self.norm1_name))

    def init_weights(self, pretrained=None):                                # :336-367
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, _BatchNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, DynamicBottleneck):
                    nn.init.constant_(m.norm3.weight, 0)

    def manipulate_stem(self, arch_meta):                                   # :381-395
        if self.deep_stem:
            sliced = [dict(zip(arch_meta, t)) for t in zip(*arch_meta.values())]
            for i, j in enumerate((0, 3, 6)):
                self.stem[j].manipulate_arch(sliced[i])
        else:
            self.conv1.manipulate_arch(arch_meta)

    def manipulate_body(self, arch_meta):                                   # :397-403
        sliced = [dict(zip(arch_meta, t)) for t in zip(*arch_meta.values())]
        for i, name in enumerate(self.res_layers):
            getattr(self, name).manipulate_arch(sliced[i])

    def forward(self, x):                                                   # :405-421
        if self.deep_stem:
            x = self.stem(x)
        else:
            x = self.relu(self.norm1(self.conv1(x)))
        x = self.maxpool(x)
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)


def cross_entropy(pred, label, ignore_index=255):
    """gaiaseg/models/losses/cross_entropy_loss.py:67-94 with weight=None, class_weight=None, reduction='mean':
    per-pixel CE (0 at ignored pixels) averaged over ALL pixels (utils.py:6-23, 45-47)."""
    loss = F.cross_entropy(pred, label, weight=None, reduction='none', ignore_index=ignore_index)
    return loss.mean()


def accuracy(pred, target):
    """gaiaseg/models/losses/accuracy.py:4-49 for topk=1, thresh=None."""
    _, pred_label = pred.topk(1, dim=1)
    pred_label = pred_label.transpose(0, 1)
    correct = pred_label.eq(target.unsqueeze(0).expand_as(pred_label))
    return correct[:1].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.numel())


class DynamicFCNHead(nn.Module, DynamicMixin):
    """gaiaseg/models/decode_heads/dynamic_fcn_head.py:23-159 (+ fcn_head.py:139-253)."""

    def __init__(self, in_channels, channels, num_classes, num_convs=2, kernel_size=3, concat_input=True,
                 dropout_ratio=0.1, conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0), ignore_index=255,
                 align_corners=False, **_ignored):
        super().__init__()
        self.in_channels, self.channels, self.num_classes, self.in_index = in_channels, channels, num_classes, in_index
        self.loss_weight = loss_decode.get('loss_weight', 1.0)
        self.ignore_index, self.align_corners = ignore_index, align_corners
        self.conv_seg = DynamicConv2d(channels, num_classes, kernel_size=1, padding=0)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.num_convs, self.concat_input = num_convs, concat_input
        convs = [DynamicConvModule(in_channels if i == 0 else channels, channels, kernel_size=kernel_size,
                                   padding=kernel_size // 2, conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg)
                 for i in range(num_convs)]
        self.convs = nn.Identity() if num_convs == 0 else nn.Sequential(*convs)
        if concat_input:
            self.conv_cat = DynamicConvModule(in_channels + channels, channels, kernel_size=kernel_size,
                                              padding=kernel_size // 2, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                              act_cfg=act_cfg)

    def init_weights(self):                                                  # fcn_head.py:175-177
        nn.init.normal_(self.conv_seg.weight, 0, 0.01)
        nn.init.constant_(self.conv_seg.bias, 0)

    def forward(self, inputs):                                               # dynamic_fcn_head.py:128-135
        x = inputs[self.in_index]
        output = self.convs(x)
        if self.concat_input:
            output = self.conv_cat(torch.cat([x, output], dim=1))
        if self.dropout is not None:                                         # fcn_head.py:248-253
            output = self.dropout(output)
        return self.conv_seg(output)

    def losses(self, seg_logit, seg_label):                                  # dynamic_fcn_head.py:137-159
        seg_logit = F.interpolate(seg_logit, size=seg_label.shape[2:], mode='bilinear',
                                  align_corners=self.align_corners)
        seg_label = seg_label.squeeze(1)
        return dict(loss_seg=self.loss_weight * cross_entropy(seg_logit, seg_label, self.ignore_index),
                    acc_seg=accuracy(seg_logit, seg_label))

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg=None):
        return self.losses(self.forward(inputs), gt_semantic_seg)


class DynamicPPM(nn.ModuleList):
    """gaiaseg/models/decode_heads/dynamic_psp_head.py:25-73."""

    def __init__(self, pool_scales, in_channels, channels, conv_cfg, norm_cfg, act_cfg, align_corners):
        super().__init__()
        self.align_corners = align_corners
        for s in pool_scales:
            self.append(nn.Sequential(nn.AdaptiveAvgPool2d(s),
                                      DynamicConvModule(in_channels, channels, 1, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                                        act_cfg=act_cfg)))

    def forward(self, x):
        return [F.interpolate(ppm(x), size=x.size()[2:], mode='bilinear', align_corners=self.align_corners)
                for ppm in self]


class DynamicPSPHead(DynamicFCNHead):
    """gaiaseg/models/decode_heads/dynamic_psp_head.py:75-173 + psp_head.py:228-241.  `channel_record` is read as a
    SEGMENTED input slice: segment k of the concatenated input uses the weight columns starting at the
    MAX-width offset of segment k; 'prefix' = plain prefix slice of the concatenation."""

    def __init__(self, in_channels, channels, num_classes, pool_scales=(1, 2, 3, 6), dropout_ratio=0.1, conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0), ignore_index=255,
                 align_corners=False, channel_record_mode='segmented', **_ignored):
        nn.Module.__init__(self)
        self.in_channels, self.channels, self.num_classes, self.in_index = in_channels, channels, num_classes, in_index
        self.loss_weight = loss_decode.get('loss_weight', 1.0)
        self.ignore_index, self.align_corners = ignore_index, align_corners
        self.channel_record_mode = channel_record_mode
        self.conv_seg = DynamicConv2d(channels, num_classes, kernel_size=1, padding=0)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.psp_modules = DynamicPPM(pool_scales, in_channels, channels, conv_cfg, norm_cfg, act_cfg, align_corners)
        self.bottleneck = DynamicConvModule(in_channels + len(pool_scales) * channels, channels, 3, padding=1,
                                            conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg)

    def forward(self, inputs):                                               # psp_head.py:228-241
        x = inputs[self.in_index]
        psp_outs = [x] + self.psp_modules(x)
        channel_record = [t.size(1) for t in psp_outs]
        cat = torch.cat(psp_outs, dim=1)
        bt = self.bottleneck
        if self.channel_record_mode == 'prefix':
            out = bt(cat)
        else:
            conv = bt.conv
            max_off = [0, self.in_channels] + [self.in_channels + self.channels * i for i in range(1, len(channel_record))]
            cols = torch.cat([torch.arange(max_off[k], max_off[k] + c) for k, c in enumerate(channel_record)])
            out = bt.activate(bt.norm(conv(cat, cols)))
        if self.dropout is not None:
            out = self.dropout(out)
        return self.conv_seg(out)


class DynamicASPPHead(DynamicFCNHead):
    """BASELINE config 3's DeepLabV3 head.  NOT in the reference tree: [EXT] mmseg ASPPHead /
    ASPPModule (mmseg/models/decode_heads/aspp_head.py) restated with the reference's DynamicConvModule, the way
    dynamic_psp_head.py:75-147 restates PSPHead:  cat[resize(image_pool(x)), aspp_d(x) for d in dilations] ->
    3x3 bottleneck -> dropout -> conv_seg; 1x1 conv for dilation 1, else 3x3 with padding = dilation."""

    def __init__(self, in_channels, channels, num_classes, dilations=(1, 6, 12, 18), dropout_ratio=0.1, conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0), ignore_index=255,
                 align_corners=False, **_ignored):
        nn.Module.__init__(self)
        self.in_channels, self.channels, self.num_classes, self.in_index = in_channels, channels, num_classes, in_index
        self.loss_weight = loss_decode.get('loss_weight', 1.0)
        self.ignore_index, self.align_corners = ignore_index, align_corners
        self.conv_seg = DynamicConv2d(channels, num_classes, kernel_size=1, padding=0)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.image_pool = nn.Sequential(nn.AdaptiveAvgPool2d(1),
                                        DynamicConvModule(in_channels, channels, 1, conv_cfg=conv_cfg, norm_cfg=norm_cfg,
                                                          act_cfg=act_cfg))
        self.aspp_modules = nn.ModuleList(
            DynamicConvModule(in_channels, channels, 1 if d == 1 else 3, dilation=d, padding=0 if d == 1 else d,
                              conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg) for d in dilations)
        self.bottleneck = DynamicConvModule((len(dilations) + 1) * channels, channels, 3, padding=1, conv_cfg=conv_cfg,
                                            norm_cfg=norm_cfg, act_cfg=act_cfg)

    def forward(self, inputs):
        x = inputs[self.in_index]
        outs = [F.interpolate(self.image_pool(x), size=x.size()[2:], mode='bilinear', align_corners=self.align_corners)]
        outs.extend(m(x) for m in self.aspp_modules)
        out = self.bottleneck(torch.cat(outs, dim=1))
        if self.dropout is not None:
            out = self.dropout(out)
        return self.conv_seg(out)


class DynamicEncoderDecoder(nn.Module, DynamicMixin):
    """gaiaseg/models/segmentors/dynamic_encoder_decoder.py:8-42 + [EXT] mmseg EncoderDecoder (inference restated
    in-tree at gaiaseg/models/segmentors/dynamic_distiller.py:252-262, 461-521)."""
    search_space = {'backbone', 'decode_head', 'neck', 'auxiliary_head'}

    def __init__(self, backbone, decode_head, auxiliary_head=None, train_cfg=None, test_cfg=None, **_ignored):
        super().__init__()
        self.backbone = _build(backbone)
        self.decode_head = _build(decode_head)
        self.auxiliary_head = _build(auxiliary_head) if auxiliary_head is not None else None
        self.backbone.init_weights()
        self.decode_head.init_weights()
        if self.auxiliary_head is not None:
            self.auxiliary_head.init_weights()

    def manipulate_backbone(self, arch_meta):
        self.backbone.manipulate_arch(arch_meta)

    def manipulate_decode_head(self, arch_meta):
        pass

    def manipulate_neck(self, arch_meta):
        pass

    def manipulate_auxiliary_head(self, arch_meta):
        pass

    def forward_train(self, img, img_metas, gt_semantic_seg):
        x = self.backbone(img)
        losses = {f'decode.{k}': v for k, v in self.decode_head.forward_train(x, img_metas, gt_semantic_seg).items()}
        if self.auxiliary_head is not None:
            losses.update({f'aux.{k}': v
                           for k, v in self.auxiliary_head.forward_train(x, img_metas, gt_semantic_seg).items()})
        return losses

    @staticmethod
    def parse_losses(losses):
        return sum(v.mean() for k, v in losses.items() if 'loss' in k)

    def encode_decode(self, img):                                             # dynamic_distiller.py:252-262
        out = self.decode_head.forward(self.backbone(img))
        return F.interpolate(out, size=img.shape[2:], mode='bilinear', align_corners=False)

    def simple_test(self, img, ori_shape=None):                               # dynamic_distiller.py:461-521
        seg_logit = self.encode_decode(img)
        if ori_shape is not None and tuple(ori_shape) != tuple(img.shape[2:]):
            seg_logit = F.interpolate(seg_logit, size=tuple(ori_shape), mode='bilinear', align_corners=False)
        return F.softmax(seg_logit, dim=1).argmax(dim=1)


_TYPES = dict(DynamicResNet=DynamicResNet, DynamicFCNHead=DynamicFCNHead, DynamicPSPHead=DynamicPSPHead,
              DynamicASPPHead=DynamicASPPHead,
              DynamicEncoderDecoder=DynamicEncoderDecoder)


def _build(cfg):
    cfg = dict(cfg)
    return _TYPES[cfg.pop('type')](**cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    cfg = dict(cfg)
    cfg.pop('type', None)
    return DynamicEncoderDecoder(**cfg, train_cfg=train_cfg, test_cfg=test_cfg)
