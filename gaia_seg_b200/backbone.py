"""DynamicResNet / DynamicResLayer -- elastic-width / elastic-depth ResNet backbone.

Mirrors gaiaseg/models/backbones/dynamic_resnet.py:25-421 and gaiaseg/models/utils/dynamic_res_layer.py:16-172
(same constructor arguments, module / parameter names, `manipulate_stem` / `manipulate_body` / `manipulate_depth`
/ `manipulate_width`, `deploy`), with every tensor op running in the sm_100a kernels:
  stem conv (im2col of the fp32 image + tcgen05 GEMM) -> DynBN -> ReLU -> maxpool -> 4 stages of fused
  DynamicBottleneck nodes.  Activations between modules are bf16 NHWC; `out_fp32=True` converts the returned
  feature maps to the reference's fp32 NCHW.
"""
import warnings
from collections.abc import Sequence

import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import functional as F_gs
from .core import (DynamicBottleneck, DynamicConv2d, DynamicMixin, Registry, build_conv_layer, build_norm_layer)

BACKBONES = Registry('backbone')


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0):
    nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


class DynamicResLayer(nn.ModuleList, DynamicMixin):
    """One stage: ModuleList of max-depth bottlenecks; forward runs the first `depth_state` blocks
    (dynamic_res_layer.py:166-172); block 0 owns the downsample branch (1x1 conv stride s -> norm, :70-94);
    `contract_dilation` halves the dilation of block 0 (:98-102)."""
    search_space = {'depth', 'width'}

    def init_state(self, depth=None, width=None, **kwargs):
        if depth is not None:
            self.depth_state = depth
        if width is not None:
            self.width_state = width
        for k, v in kwargs.items():
            setattr(self, f'{k}_state', v)

    def __init__(self, block, inplanes, planes, depth, stride=1, dilation=1, avg_down=False, conv_cfg=None,
                 norm_cfg=None, downsample_first=True, contract_dilation=False, **kwargs):
        if conv_cfg is None or conv_cfg.get('type') != 'DynConv2d':
            warnings.warn('Non-dynamic-conv detected in dynamic block.')
        if norm_cfg is None or 'Dyn' not in norm_cfg.get('type', ''):
            warnings.warn('Non-dynamic-bn detected in dynamic block.')
        assert downsample_first, 'downsample_first=False (Hourglass) is not supported (dynamic_res_layer.py:128)'
        if avg_down:
            raise NotImplementedError('avg_down=True is unused by every GAIA-seg config (SURVEY K7)')
        self.block = block
        self.avg_down = avg_down
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                build_conv_layer(conv_cfg, inplanes, planes * block.expansion, kernel_size=1, padding=0, stride=stride,
                                 bias=False),
                build_norm_layer(norm_cfg, planes * block.expansion)[1])
        first_dilation = dilation // 2 if (dilation > 1 and contract_dilation) else dilation
        layers = [block(inplanes=inplanes, planes=planes, stride=stride, dilation=first_dilation, downsample=downsample,
                        conv_cfg=conv_cfg, norm_cfg=norm_cfg, **kwargs)]
        inplanes = planes * block.expansion
        for _ in range(1, depth):
            layers.append(block(inplanes=inplanes, planes=planes, stride=1, dilation=dilation, conv_cfg=conv_cfg,
                                norm_cfg=norm_cfg, **kwargs))
        super().__init__(layers)
        self.init_state(depth=depth, width=planes)

    def manipulate_depth(self, depth):
        assert depth >= 1, 'Depth must be greater than 0, skipping stage is not supported yet.'
        assert depth <= len(self), f'depth {depth} exceeds the max depth {len(self)} of this stage'
        self.depth_state = depth

    def manipulate_width(self, width):
        self.width_state = width
        for m in self:
            m.manipulate_width(width)

    def deploy_forward(self, x):
        del self[self.depth_state:]
        for i in range(self.depth_state):
            x = self[i](x)
        return x

    def forward(self, x):
        if getattr(self, '_deploying', False):
            return self.deploy_forward(x)
        blocks = [self[i] for i in range(self.depth_state)]
        if all(type(b) is DynamicBottleneck and (b.downsample is None or len(b.downsample) == 2) for b in blocks):
            return F_gs.res_stage(x, blocks)      # one autograd node per stage (fused BN-backward reductions)
        for b in blocks:
            x = b(x)
        return x


@BACKBONES.register_module()
class DynamicResNet(nn.Module, DynamicMixin):
    search_space = {'stem', 'body'}

    def init_state(self, stem=None, body=None, **kwargs):
        if stem is not None:
            self.stem_state = stem
        if body is not None:
            self.body_state = body
        for k, v in kwargs.items():
            setattr(self, f'{k}_state', v)

    def __init__(self, in_channels, stem_width, body_width, body_depth, num_stages=4, strides=(1, 2, 2, 2),
                 dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style='pytorch', deep_stem=False, avg_down=False,
                 frozen_stages=-1, frozen_layers=None, conv_cfg=None, norm_cfg=dict(type='DynSyncBN'),
                 act_cfg=dict(type='ReLU'), norm_eval=False, dcn=None, stage_with_dcn=(False, False, False, False),
                 plugins=None, with_cp=False, zero_init_residual=True, contract_dilation=False, out_fp32=False):
        super().__init__()
        assert 1 <= num_stages <= 4
        assert len(strides) == len(dilations) == num_stages
        assert max(out_indices) < num_stages
        if dcn is not None or plugins is not None:
            raise NotImplementedError('dcn / plugins are not on the GAIA-seg hot path')
        self.stem_width, self.body_width = stem_width, body_width
        self.num_stages, self.strides, self.dilations, self.out_indices = num_stages, strides, dilations, out_indices
        self.style, self.deep_stem, self.avg_down = style, deep_stem, avg_down
        self.frozen_stages, self.frozen_layers = frozen_stages, frozen_layers
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.with_cp, self.norm_eval = with_cp, norm_eval
        self.contract_dilation, self.zero_init_residual = contract_dilation, zero_init_residual
        self.out_fp32 = out_fp32
        self.block = DynamicBottleneck
        self.body_depth = body_depth[:num_stages]
        self.inplanes = stem_width[-1] if deep_stem else stem_width
        self.init_state(stem={'width': stem_width}, body={'depth': body_depth, 'width': body_width})
        self._make_stem_layer(in_channels, stem_width)
        self.res_layers = []
        for i, num_blocks in enumerate(self.body_depth):
            planes = body_width[i]
            res_layer = self.make_res_layer(block=self.block, inplanes=self.inplanes, planes=planes, depth=num_blocks,
                                            stride=strides[i], dilation=dilations[i], style=self.style,
                                            avg_down=self.avg_down, with_cp=with_cp, conv_cfg=conv_cfg,
                                            norm_cfg=norm_cfg, contract_dilation=contract_dilation)
            self.inplanes = planes * self.block.expansion
            layer_name = f'layer{i + 1}'
            self.add_module(layer_name, res_layer)
            self.res_layers.append(layer_name)
        self._freeze_stages()
        self._freeze_layers()
        self.feat_dim = self.block.expansion * body_width[0] * 2 ** (len(self.body_depth) - 1)
        self.active_feat_dim = self.feat_dim

    def make_res_layer(self, **kwargs):
        return DynamicResLayer(**kwargs)

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    def _make_stem_layer(self, in_channels, stem_width):
        if self.deep_stem:
            assert isinstance(stem_width, Sequence)
            chans = [in_channels] + list(stem_width)
            mods = []
            for i in range(3):
                mods += [build_conv_layer(self.conv_cfg, chans[i], chans[i + 1], kernel_size=3,
                                          stride=2 if i == 0 else 1, padding=1, bias=False),
                         build_norm_layer(self.norm_cfg, chans[i + 1])[1], nn.ReLU(inplace=True)]
            self.stem = nn.Sequential(*mods)
        else:
            self.conv1 = build_conv_layer(self.conv_cfg, in_channels, stem_width, kernel_size=7, stride=2, padding=3,
                                          bias=False)
            self.norm1_name, norm1 = build_norm_layer(self.norm_cfg, stem_width, postfix=1)
            self.add_module(self.norm1_name, norm1)
            self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            if self.deep_stem:
                self.stem.eval()
                for p in self.stem.parameters():
                    p.requires_grad = False
            else:
                self.norm1.eval()
                for m in (self.conv1, self.norm1):
                    for p in m.parameters():
                        p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f'layer{i}')
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def _freeze_layers(self):
        if self.frozen_layers is not None:
            for i, layer_name in enumerate(self.res_layers):
                res_layer = getattr(self, layer_name)
                n = self.frozen_layers[i]
                assert n <= len(res_layer)
                for j in range(n):
                    res_layer[j].eval()
                    for p in res_layer[j].parameters():
                        p.requires_grad = False

    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):
            from .runner import load_checkpoint
            load_checkpoint(self, pretrained, strict=False)
        elif pretrained is None:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    kaiming_init(m)
                elif isinstance(m, (_BatchNorm, nn.GroupNorm)):
                    constant_init(m, 1)
            if self.zero_init_residual:
                for m in self.modules():
                    if isinstance(m, DynamicBottleneck):
                        constant_init(m.norm3, 0)
        else:
            raise TypeError('pretrained must be a str or None')

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        self._freeze_layers()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, _BatchNorm):
                    m.eval()
        return self

    def manipulate_stem(self, arch_meta):
        """arch_meta = {'width': 32} or {'width': [16, 16, 32]} for deep_stem (dynamic_resnet.py:381-395)."""
        self.stem_state = arch_meta
        if self.deep_stem:
            sliced = [dict(zip(arch_meta, t)) for t in zip(*arch_meta.values())]
            self.stem[0].manipulate_arch(sliced[0])
            self.stem[3].manipulate_arch(sliced[1])
            self.stem[6].manipulate_arch(sliced[2])
        else:
            self.conv1.manipulate_arch(arch_meta)

    def manipulate_body(self, arch_meta):
        self.body_state = arch_meta
        sliced = [dict(zip(arch_meta, t)) for t in zip(*arch_meta.values())]
        for i, layer_name in enumerate(self.res_layers):
            getattr(self, layer_name).manipulate_arch(sliced[i])

    def _stem_forward(self, x):
        deploying = getattr(self, '_deploying', False)
        if self.deep_stem:
            for i in (0, 3, 6):
                conv, bn = self.stem[i], self.stem[i + 1]
                if deploying:
                    conv.deploy_slice(x.size(1))
                    bn.deploy_slice(conv.width_state)
                x = F_gs.conv_bn_act(x, conv, bn, relu=True)
        else:
            if deploying:
                self.conv1.deploy_slice(x.size(1))
                self.norm1.deploy_slice(self.conv1.width_state)
            x = F_gs.conv_bn_act(x, self.conv1, self.norm1, relu=True)
        return x

    def forward(self, x):
        x = self._stem_forward(x)
        x = F_gs.maxpool3x3s2(x)
        outs = []
        for i, layer_name in enumerate(self.res_layers):
            x = getattr(self, layer_name)(x)
            if i in self.out_indices:
                outs.append(F_gs.to_nchw_f32(x) if self.out_fp32 else x)
        return tuple(outs)
