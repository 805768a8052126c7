"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN CODE in this container.

Run here (needs /root/reference, which does not exist on the GPU box):  python tests/golden/make_golden.py

The reference (pure Python) cannot be imported as a package: gaiaseg/__init__.py pulls mmcv / mmseg / gaiavision,
none of which is installed or installable (no network).  Two things can still be pinned against its code:

 A. loss / accuracy arithmetic -- `cross_entropy`, `reduce_loss`, `weight_reduce_loss`, `accuracy` are pure-torch
    functions (gaiaseg/models/losses/{cross_entropy_loss,utils,accuracy}.py); their source is extracted with `ast`
    (so the dead `..builder` import and the `pdb` debris around them are not executed) and run on seeded inputs.
 B. the in-tree model WIRING -- gaiaseg/models/backbones/dynamic_resnet.py, gaiaseg/models/utils/dynamic_res_layer.py,
    gaiaseg/models/decode_heads/{fcn_head,dynamic_fcn_head}.py are imported unmodified from /root/reference with stub
    `mmcv` / `mmseg` / `gaiavision` packages.  The stubs' dynamic operators (DynamicConv2d, DynBN, DynamicBottleneck,
    DynamicConvModule -- gaiavision, not vendored) are the ORACLE's restatements, so what gets pinned is: module tree,
    parameter names, stem / stage / downsample / dilation wiring, manipulate_stem / manipulate_body fan-out, depth
    truncation, FCN head forward + losses().  Outputs for a few sub-nets are stored as fixtures.

Fixtures: tests/golden/loss_golden.npz, tests/golden/wiring_golden.npz (a few hundred KB).
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from oracle import ref_model as O  # noqa: E402


def extract_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'torch': torch, 'nn': nn, 'F': F, 'np': np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, 'exec')
            exec(code, ns)
    return ns


def make_loss_golden():
    ns = extract_functions(f'{REF}/gaiaseg/models/losses/utils.py', {'reduce_loss', 'weight_reduce_loss'})
    ns.update({k: v for k, v in extract_functions(f'{REF}/gaiaseg/models/losses/cross_entropy_loss.py',
                                                  {'cross_entropy'}).items() if k == 'cross_entropy'})
    ns['cross_entropy'].__globals__.update(weight_reduce_loss=ns['weight_reduce_loss'], F=F)
    acc_ns = extract_functions(f'{REF}/gaiaseg/models/losses/accuracy.py', {'accuracy'})
    out = {}
    for i, (N, K, H, W, ign) in enumerate([(2, 19, 24, 40, 0.1), (1, 150, 16, 16, 0.3), (2, 5, 8, 8, 1.0), (3, 19, 9, 7, 0.0)]):
        g = torch.Generator().manual_seed(100 + i)
        pred = torch.randn(N, K, H, W, generator=g) * 2
        lab = torch.randint(0, K, (N, H, W), generator=g)
        lab[torch.rand(N, H, W, generator=g) < ign] = 255
        loss = ns['cross_entropy'](pred, lab, weight=None, class_weight=None, reduction='mean', avg_factor=None,
                                   ignore_index=255)
        acc = acc_ns['accuracy'](pred, lab)
        out[f'pred{i}'], out[f'lab{i}'] = pred.numpy(), lab.numpy()
        out[f'loss{i}'], out[f'acc{i}'] = loss.numpy(), acc.numpy()
    # ties: the first class index must win (topk) -- integer-valued scores
    g = torch.Generator().manual_seed(7)
    pred = torch.randint(0, 3, (1, 6, 10, 10), generator=g).float()
    lab = torch.randint(0, 6, (1, 10, 10), generator=g)
    out['pred_tie'], out['lab_tie'] = pred.numpy(), lab.numpy()
    out['acc_tie'] = acc_ns['accuracy'](pred, lab).numpy()
    np.savez_compressed(os.path.join(HERE, 'loss_golden.npz'), **out)
    print('loss_golden.npz:', {k: v.shape for k, v in out.items() if k.startswith(('loss', 'acc'))})


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    def kaiming_init(m, a=0, mode='fan_out', nonlinearity='relu', bias=0):
        nn.init.kaiming_normal_(m.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        if getattr(m, 'bias', None) is not None:
            nn.init.constant_(m.bias, bias)

    def constant_init(m, val, bias=0):
        nn.init.constant_(m.weight, val)
        if getattr(m, 'bias', None) is not None:
            nn.init.constant_(m.bias, bias)

    def normal_init(m, mean=0, std=1, bias=0):
        nn.init.normal_(m.weight, mean, std)
        if getattr(m, 'bias', None) is not None:
            nn.init.constant_(m.bias, bias)

    ident = lambda *a, **k: (lambda f: f)

    class _CE(nn.Module):
        def __init__(self, use_sigmoid=False, loss_weight=1.0, **kw):
            super().__init__()
            self.loss_weight = loss_weight

        def forward(self, cls_score, label, weight=None, ignore_index=255, **kw):
            return self.loss_weight * O.cross_entropy(cls_score, label, ignore_index)

    def build_loss(cfg):
        cfg = dict(cfg)
        assert cfg.pop('type') == 'CrossEntropyLoss'
        return _CE(**cfg)

    def resize(input, size=None, scale_factor=None, mode='nearest', align_corners=None, warning=True):
        return F.interpolate(input, size, scale_factor, mode, align_corners)

    mod('mmcv')
    mod('mmcv.cnn', build_plugin_layer=None, constant_init=constant_init, kaiming_init=kaiming_init,
        normal_init=normal_init, build_conv_layer=O.build_conv_layer, build_activation_layer=None, ConvModule=None)
    mod('mmcv.runner', load_checkpoint=None, auto_fp16=ident, force_fp32=ident)
    mod('mmseg')
    mod('mmseg.utils', get_root_logger=lambda *a, **k: None)
    mod('mmseg.models')
    mod('mmseg.models.builder', BACKBONES=_Registry(), HEADS=_Registry(), build_loss=build_loss)
    mod('mmseg.models.utils', ResLayer=None)
    mod('mmseg.models.losses', accuracy=O.accuracy)
    mod('mmseg.core', build_pixel_sampler=None)
    mod('mmseg.ops', resize=resize)
    mod('gaiavision')
    mod('gaiavision.core', DynamicMixin=O.DynamicMixin, DynamicConv2d=O.DynamicConv2d)
    mod('gaiavision.core.bricks', build_norm_layer=O.build_norm_layer, DynamicBottleneck=O.DynamicBottleneck,
        DynamicConvModule=O.DynamicConvModule)
    for pkg in ('gaiaseg', 'gaiaseg.models', 'gaiaseg.models.utils', 'gaiaseg.models.backbones',
                'gaiaseg.models.decode_heads'):
        mod(pkg)


def load_ref(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


WIRING_CASES = {
    # name: (backbone kwargs, list of arch metas)
    'os32': (dict(in_channels=3, stem_width=16, body_width=[16, 24, 32, 48], body_depth=[2, 2, 3, 2]),
             [{'stem': {'width': 16}, 'body': {'width': [16, 24, 32, 48], 'depth': [2, 2, 3, 2]}},
              {'stem': {'width': 8}, 'body': {'width': [8, 16, 24, 32], 'depth': [1, 2, 1, 1]}}]),
    'os8_v1c': (dict(in_channels=3, stem_width=[8, 8, 16], body_width=[16, 24, 32, 48], body_depth=[2, 2, 3, 2],
                     deep_stem=True, strides=(1, 2, 1, 1), dilations=(1, 1, 2, 4), contract_dilation=True),
                [{'stem': {'width': [8, 8, 16]}, 'body': {'width': [16, 24, 32, 48], 'depth': [2, 2, 3, 2]}},
                 {'stem': {'width': [4, 8, 8]}, 'body': {'width': [8, 24, 24, 40], 'depth': [2, 1, 2, 2]}}]),
}
HEAD_KW = dict(in_channels=192, channels=32, num_classes=7, num_convs=2, concat_input=True, dropout_ratio=0.0,
               in_index=3, loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.4))
CFG_COMMON = dict(conv_cfg=dict(type='DynConv2d'), norm_cfg=dict(type='DynBN', requires_grad=True))


def seeded_params(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in sorted(module.named_parameters()):
            if p.dim() == 4:
                p.copy_(torch.randn(p.shape, generator=g) / (p.shape[1] * p.shape[2] * p.shape[3]) ** 0.5)
            elif n.endswith('weight'):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def make_wiring_golden():
    install_stubs()
    utils = load_ref('gaiaseg.models.utils.dynamic_res_layer', 'gaiaseg/models/utils/dynamic_res_layer.py')
    sys.modules['gaiaseg.models.utils'].DynamicResLayer = utils.DynamicResLayer
    bb_mod = load_ref('gaiaseg.models.backbones.dynamic_resnet', 'gaiaseg/models/backbones/dynamic_resnet.py')
    load_ref('gaiaseg.models.decode_heads.fcn_head', 'gaiaseg/models/decode_heads/fcn_head.py')
    head_mod = load_ref('gaiaseg.models.decode_heads.dynamic_fcn_head', 'gaiaseg/models/decode_heads/dynamic_fcn_head.py')
    out = {}
    for cname, (bb_kw, archs) in WIRING_CASES.items():
        torch.manual_seed(0)
        ref_bb = bb_mod.DynamicResNet(**bb_kw, **CFG_COMMON)
        ref_head = head_mod.DynamicFCNHead(**HEAD_KW, **CFG_COMMON)
        seeded_params(ref_bb, 1)
        seeded_params(ref_head, 2)
        out[f'{cname}.backbone_keys'] = np.array(sorted(ref_bb.state_dict().keys()))
        out[f'{cname}.head_keys'] = np.array(sorted(ref_head.state_dict().keys()))
        g = torch.Generator().manual_seed(11)
        img = torch.randn(2, 3, 64, 64, generator=g)
        lab = torch.randint(0, 7, (2, 1, 64, 64), generator=g)
        lab[torch.rand(2, 1, 64, 64, generator=g) < 0.1] = 255
        for ai, arch in enumerate(archs):
            ref_bb.manipulate_arch(arch)
            ref_bb.train(); ref_head.train()
            feats = ref_bb(img)
            losses = ref_head.forward_train(feats, None, lab, None)
            for fi, f in enumerate(feats):
                out[f'{cname}.arch{ai}.feat{fi}'] = f.detach().numpy()
            out[f'{cname}.arch{ai}.loss_seg'] = losses['loss_seg'].detach().numpy()
            out[f'{cname}.arch{ai}.acc_seg'] = losses['acc_seg'].detach().numpy()
            ref_bb.eval(); ref_head.eval()
            with torch.no_grad():
                out[f'{cname}.arch{ai}.eval_logits'] = ref_head(ref_bb(img)).numpy()
    np.savez_compressed(os.path.join(HERE, 'wiring_golden.npz'), **out)
    print('wiring_golden.npz:', len(out), 'arrays,', os.path.getsize(os.path.join(HERE, 'wiring_golden.npz')) // 1024, 'KB')


if __name__ == '__main__':
    make_loss_golden()
    make_wiring_golden()
