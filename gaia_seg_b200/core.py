"""Host-side mirror of the gaiavision operator API the GAIA-seg hot path is written against
(`gaiavision.core`, `gaiavision.core.bricks` -- not vendored by the reference, contracts inferred from the
call sites listed in SURVEY.md 2.1):

    DynamicMixin, DynamicConv2d ('DynConv2d'), DynamicBatchNorm2d ('DynBN'), DynamicSyncBatchNorm ('DynSyncBN'),
    DynamicBottleneck, DynamicConvModule, build_conv_layer, build_norm_layer, Registry / build_from_cfg.

Same class names, constructor arguments, `manipulate_*` / `deploy` methods and state_dict layout
(max-width OIHW fp32).  The arithmetic is NOT here: `forward` hands the tensors to the sm_100a kernels
through gaia_seg_b200.functional; a CPU tensor raises (no fallback).
"""
import inspect

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import functional as F_gs


# ------------------------------------------------------------------------------------------------
# registry (mmcv.utils.Registry / build_from_cfg semantics, minimal)
# ------------------------------------------------------------------------------------------------
class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def _register(self, cls, name=None, force=False):
        if not inspect.isclass(cls) and not callable(cls):
            raise TypeError(f'module must be a class or callable, got {type(cls)}')
        names = [cls.__name__] if name is None else ([name] if isinstance(name, str) else list(name))
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f'{n} is already registered in {self._name}')
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _dec(cls):
            self._register(cls, name, force)
            return cls
        return _dec


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg and not (default_args and 'type' in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif inspect.isclass(obj_type) or callable(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    return obj_cls(**args)


CONV_LAYERS = Registry('conv layer')
NORM_LAYERS = Registry('norm layer')


# ------------------------------------------------------------------------------------------------
# DynamicMixin
# ------------------------------------------------------------------------------------------------
class DynamicMixin:
    """search_space names the keys `manipulate_arch` accepts; each key k is routed to
    `self.manipulate_<k>(value)` (evidence: gaiaseg/models/backbones/dynamic_resnet.py:381-403,
    gaiaseg/models/segmentors/dynamic_encoder_decoder.py:31-42)."""
    search_space = set()

    def init_state(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, f'{k}_state', v)

    def manipulate_arch(self, arch_meta):
        if arch_meta is None:
            return
        for k, v in arch_meta.items():
            fn = getattr(self, f'manipulate_{k}', None)
            if fn is None:
                raise KeyError(f'{type(self).__name__}: "{k}" is not in the search space {sorted(self.search_space)}')
            fn(v)

    def deploy(self, mode=True):
        """Switch every dynamic sub-module to deploy mode: the next forward physically slices the
        parameters to the active sub-net (tools/extract_subnet.py:94,130)."""
        self._deploying = mode
        if isinstance(self, nn.Module):
            for m in self.children():
                _deploy_recursive(m, mode)

    def arch_state(self):
        return {k: getattr(self, f'{k}_state', None) for k in self.search_space}


def _deploy_recursive(m, mode):
    if isinstance(m, DynamicMixin):
        m.deploy(mode)
    else:
        for c in m.children():
            _deploy_recursive(c, mode)


# ------------------------------------------------------------------------------------------------
# DynamicConv2d
# ------------------------------------------------------------------------------------------------
@CONV_LAYERS.register_module(name='DynConv2d')
class DynamicConv2d(nn.Conv2d, DynamicMixin):
    """nn.Conv2d holding the MAX-width weight [Co_max, Ci_max, kh, kw]; forward uses the channel-prefix
    slice W[:width_state, :x.size(1)] -- input channels follow the input, they are never configured.

    The fp32 master is stored channels_last ([Co][kh][kw][Ci] in memory) so the bf16 shadow the
    tcgen05 kernel reads through TMA is an element-wise cast and the weight gradient is written in
    place; `state_dict()` still shows logical OIHW tensors (the reference checkpoint format)."""
    search_space = {'width'}

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros'):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                         groups=groups, bias=bias, padding_mode=padding_mode)
        if padding_mode != 'zeros':
            raise NotImplementedError('DynamicConv2d: only zero padding')
        self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)
        self.init_state(width=out_channels)

    def manipulate_width(self, width):
        assert 0 < width <= self.out_channels, f'width {width} exceeds max width {self.out_channels}'
        self.width_state = width

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._gs_key = None  # device / dtype moved: shadows are stale
        return out

    def deploy_slice(self, ci):
        """Overwrite the parameters with the active slice [:width_state, :ci] (extract_subnet)."""
        co = self.width_state
        if self.weight.size(0) != co or self.weight.size(1) != ci:
            w = self.weight.data[:co, :ci].contiguous(memory_format=torch.channels_last)
            self.weight = nn.Parameter(w, requires_grad=self.weight.requires_grad)
            if self.bias is not None:
                self.bias = nn.Parameter(self.bias.data[:co].clone(), requires_grad=self.bias.requires_grad)
            self.out_channels, self.in_channels = co, ci
            self._gs_key = None

    def deploy_forward(self, x):
        self.deploy_slice(x.size(1))
        return self._run(x)

    def _run(self, x):
        return F_gs.conv_bn_act(x, self, None, relu=False, Co=self.width_state)

    def forward(self, x):
        if getattr(self, '_deploying', False):
            return self.deploy_forward(x)
        return self._run(x)


# ------------------------------------------------------------------------------------------------
# DynamicBatchNorm2d / DynamicSyncBatchNorm
# ------------------------------------------------------------------------------------------------
@NORM_LAYERS.register_module(name='DynBN')
class DynamicBatchNorm2d(_BatchNorm, DynamicMixin):
    """_BatchNorm with max-width weight / bias / running stats; forward slices all four to [:x.size(1)].
    Train mode normalises with mini-batch statistics and updates the running-stat prefix (momentum 0.1,
    unbiased variance); eval mode uses running_*[:C]."""
    sync = False
    search_space = set()

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self._register_state_dict_hook(_flush_nbt)

    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError(f'expected 4D input (got {input.dim()}D input)')

    def deploy_slice(self, c):
        if self.num_features != c:
            if self.affine:
                self.weight = nn.Parameter(self.weight.data[:c].clone(), requires_grad=self.weight.requires_grad)
                self.bias = nn.Parameter(self.bias.data[:c].clone(), requires_grad=self.bias.requires_grad)
            if self.running_mean is not None:
                self.running_mean = self.running_mean[:c].clone()
                self.running_var = self.running_var[:c].clone()
            self.num_features = c

    def deploy_forward(self, x):
        self.deploy_slice(x.size(1))
        return self._run(x)

    def _run(self, x):
        return DynBNFn.apply(x, self.weight, self)

    def forward(self, x):
        self._check_input_dim(x)
        if x.size(1) > self.num_features:
            raise ValueError(f'input has {x.size(1)} channels, DynBN max width is {self.num_features}')
        if getattr(self, '_deploying', False):
            return self.deploy_forward(x)
        return self._run(x)


def _flush_nbt(module, state_dict, prefix, local_metadata):
    pend = getattr(module, '_gs_nbt_pending', 0)
    if pend and module.num_batches_tracked is not None:
        module.num_batches_tracked += pend
        module._gs_nbt_pending = 0
        state_dict[prefix + 'num_batches_tracked'] = module.num_batches_tracked
    return state_dict


@NORM_LAYERS.register_module(name='DynSyncBN')
class DynamicSyncBatchNorm(DynamicBatchNorm2d):
    """DynBN whose statistics are summed over the data-parallel group: one packed all-reduce of
    (sum, sum of squares) per layer forward and of (sum dy, sum dy*xhat) backward.  `group_size` is
    accepted as in the reference config (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23);
    the group is the whole world unless `process_group` is set."""
    sync = True

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, group_size=None,
                 process_group=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.group_size = group_size
        self.process_group = process_group


@NORM_LAYERS.register_module(name='SyncBN')
class SyncBatchNorm2d(DynamicSyncBatchNorm):
    """Static-width SyncBN used by the heads (pspnet_ar50to101v2_gsync.py:34,48); same kernels."""


@NORM_LAYERS.register_module(name='BN')
class BatchNorm2d(DynamicBatchNorm2d):
    """Static-width BN; same kernels."""


class DynBNFn(torch.autograd.Function):
    """Stand-alone DynBN (not preceded by one of our convs): stats kernel + apply kernel."""

    @staticmethod
    def forward(ctx, x, weight, bn):
        from . import _lib
        _lib.require_device()
        x = F_gs.as_act(x)
        C = x.shape[1]
        if F_gs.bn_batch_mode(bn):
            stats = F_gs.bn_stats(x)
            z, aff, count = F_gs.bn_train_apply(bn, x, stats, C)
            ctx.mode = 'batch'
        else:
            aff = F_gs.bn_eval_affine(bn, C)
            z = F_gs.bn_apply(x, aff[0], aff[1])
            count = 0.0
            ctx.mode = 'eval'
        ctx.x, ctx.aff, ctx.count, ctx.bn = x, aff, count, bn
        return z

    @staticmethod
    def backward(ctx, dz):
        from ._lib import call
        dz = F_gs.as_act(dz)
        x, aff, bn = ctx.x, ctx.aff, ctx.bn
        N, C, H, W = x.shape
        if ctx.mode == 'batch':
            dx, _ = F_gs.bn_backward(bn, dz, x, aff, ctx.count, None, False, False)
        else:
            dx = F_gs.new_act(N, C, H, W, x.device)
            call('gs_affine_bwd', dz.data_ptr(), F_gs.act_ld(dz), None, 0, aff[0].data_ptr(), N * H * W, C,
                 dx.data_ptr(), C, None, 0, F_gs._stream())
        ctx.x = None
        return dx, None, None


def build_conv_layer(cfg, *args, **kwargs):
    """mmcv.cnn.build_conv_layer: cfg None -> plain width conv (still our kernels), 'DynConv2d' -> dynamic."""
    if cfg is None:
        cfg_ = dict(type='DynConv2d')
    else:
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise KeyError('the cfg dict must contain the key "type"')
        cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    if layer_type in ('Conv2d', 'Conv'):
        layer_type = 'DynConv2d'
    cls = CONV_LAYERS.get(layer_type)
    if cls is None:
        raise KeyError(f'Unrecognized conv type {layer_type}')
    return cls(*args, **kwargs, **cfg_)


_NORM_ABBR = {'DynBN': 'bn', 'DynSyncBN': 'bn', 'SyncBN': 'bn', 'BN': 'bn'}


def build_norm_layer(cfg, num_features, postfix=''):
    """gaiavision / mmcv build_norm_layer: returns (name, layer); name = abbreviation + postfix ('bn1')."""
    if not isinstance(cfg, dict) or 'type' not in cfg:
        raise KeyError('the cfg dict must contain the key "type"')
    cfg_ = dict(cfg)
    layer_type = cfg_.pop('type')
    cls = NORM_LAYERS.get(layer_type)
    if cls is None:
        raise KeyError(f'Unrecognized norm type {layer_type}')
    requires_grad = cfg_.pop('requires_grad', True)
    cfg_.setdefault('eps', 1e-5)
    if not issubclass(cls, DynamicSyncBatchNorm):
        cfg_.pop('group_size', None)
    layer = cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return _NORM_ABBR[layer_type] + str(postfix), layer


# ------------------------------------------------------------------------------------------------
# bricks
# ------------------------------------------------------------------------------------------------
class DynamicBottleneck(nn.Module, DynamicMixin):
    """Post-activation bottleneck, expansion 4, stride on conv2 (`style='pytorch'`):
        relu(norm3(conv3(relu(norm2(conv2(relu(norm1(conv1 x))))))) + (downsample(x) or x))
    manipulate_width(w): conv1, conv2 -> w ; conv3, downsample conv -> 4w
    (constructed at gaiaseg/models/utils/dynamic_res_layer.py:106-125; width rule restated in-tree at
    gaiaseg/models/backbones/elastic_convformer.py:334-341)."""
    expansion = 4
    search_space = {'width'}

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch', with_cp=False,
                 conv_cfg=None, norm_cfg=dict(type='DynBN'), dcn=None, plugins=None):
        super().__init__()
        assert style in ('pytorch', 'caffe')
        if dcn is not None or plugins:
            raise NotImplementedError('DynamicBottleneck: dcn / plugins are not on the GAIA-seg hot path')
        self.inplanes, self.planes, self.stride, self.dilation = inplanes, planes, stride, dilation
        self.style, self.with_cp = style, with_cp
        self.conv1_stride, self.conv2_stride = (1, stride) if style == 'pytorch' else (stride, 1)
        self.norm1_name, norm1 = build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = build_norm_layer(norm_cfg, planes, postfix=2)
        self.norm3_name, norm3 = build_norm_layer(norm_cfg, planes * self.expansion, postfix=3)
        self.conv1 = build_conv_layer(conv_cfg, inplanes, planes, kernel_size=1, stride=self.conv1_stride, bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = build_conv_layer(conv_cfg, planes, planes, kernel_size=3, stride=self.conv2_stride,
                                      padding=dilation, dilation=dilation, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.conv3 = build_conv_layer(conv_cfg, planes, planes * self.expansion, kernel_size=1, bias=False)
        self.add_module(self.norm3_name, norm3)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.init_state(width=planes)

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    @property
    def norm2(self):
        return getattr(self, self.norm2_name)

    @property
    def norm3(self):
        return getattr(self, self.norm3_name)

    def manipulate_width(self, width):
        self.width_state = width
        self.conv1.manipulate_width(width)
        self.conv2.manipulate_width(width)
        self.conv3.manipulate_width(width * self.expansion)
        if self.downsample is not None:
            for m in self.downsample:
                if isinstance(m, DynamicConv2d):
                    m.manipulate_width(width * self.expansion)

    def forward(self, x):
        if getattr(self, '_deploying', False):
            return self.deploy_forward(x)
        if self.downsample is not None and len(self.downsample) != 2:
            raise NotImplementedError('avg_down downsample branches are not on the GAIA-seg hot path')
        return F_gs.bottleneck(x, self)

    def deploy_forward(self, x):
        """Slice every conv / norm of the block to the active sub-net, then run the fused block."""
        ci, w = x.size(1), self.width_state
        self.conv1.deploy_slice(ci)
        self.norm1.deploy_slice(w)
        self.conv2.deploy_slice(w)
        self.norm2.deploy_slice(w)
        self.conv3.deploy_slice(w)
        self.norm3.deploy_slice(w * self.expansion)
        if self.downsample is not None:
            self.downsample[0].deploy_slice(ci)
            self.downsample[1].deploy_slice(w * self.expansion)
        return F_gs.bottleneck(x, self)


class DynamicConvModule(nn.Module, DynamicMixin):
    """mmcv ConvModule analogue: conv (bias iff no norm) -> norm -> act, from conv_cfg / norm_cfg / act_cfg
    (gaiaseg/models/decode_heads/dynamic_fcn_head.py:94-126).  The three steps run as one fused call."""
    search_space = {'width'}

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias='auto',
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True, order=('conv', 'norm', 'act')):
        super().__init__()
        assert tuple(order) == ('conv', 'norm', 'act'), 'only conv -> norm -> act is on the hot path'
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if self.with_activation and act_cfg.get('type', 'ReLU') != 'ReLU':
            raise NotImplementedError('DynamicConvModule: only ReLU')
        if bias == 'auto':
            bias = not self.with_norm
        self.with_bias = bias
        self.conv = build_conv_layer(conv_cfg if conv_cfg is not None else dict(type='DynConv2d'), in_channels,
                                     out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                                     groups=groups, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        else:
            self.norm_name = None
        if self.with_activation:
            self.activate = nn.ReLU(inplace=inplace)
        self.init_state(width=out_channels)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.norm_name else None

    def manipulate_width(self, width):
        self.width_state = width
        self.conv.manipulate_width(width)

    def forward(self, x, channel_record=None, grad_carrier=None):
        if channel_record is not None:
            raise NotImplementedError('channel_record (segmented input slice) is handled by DynamicPSPHead')
        if getattr(self, '_deploying', False):
            self.conv.deploy_slice(x.size(1))
            if self.with_norm:
                self.norm.deploy_slice(self.conv.width_state)
        return F_gs.conv_bn_act(x, self.conv, self.norm, relu=self.with_activation, Co=self.conv.width_state,
                                grad_carrier=grad_carrier)
