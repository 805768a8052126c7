"""The C-ABI library loads and exports every symbol include/gaiaseg_b200.h declares, with matching arity
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gaiaseg_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'typedef struct.*?}\s*\w+;', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b([A-Za-z_][\w \*]*?)\b(gs_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[m.group(2)] = n
    return out


def test_header_declares_hot_path_entry_points():
    fns = declared_functions()
    for must in ('gs_conv2d_fwd', 'gs_conv2d_dgrad', 'gs_conv2d_wgrad', 'gs_bn_apply_train', 'gs_bn_bwd_reduce',
                 'gs_bn_bwd_apply', 'gs_upsample_ce_fwd', 'gs_upsample_ce_bwd', 'gs_upsample_argmax',
                 'gs_maxpool3x3s2_fwd', 'gs_sgd_flat', 'gs_device_check', 'gs_last_error'):
        assert must in fns, must
    assert len(fns) >= 35


def test_library_exports_every_declared_symbol(gs):
    lib = gs._lib.load()
    fns = declared_functions()
    missing = [f for f in fns if not hasattr(lib, f)]
    assert not missing, missing


def test_python_prototypes_match_header(gs):
    fns = declared_functions()
    protos = gs._lib.PROTOTYPES
    assert set(protos) == set(fns), (set(protos) ^ set(fns))
    bad = {f: (len(protos[f][1]), n) for f, n in fns.items() if len(protos[f][1]) != n}
    assert not bad, bad


def test_version_and_error_string_without_gpu(gs):
    lib = gs._lib.load()
    assert lib.gs_version() == 1
    assert isinstance(gs._lib.last_error(), str)
    assert ctypes.sizeof(gs._lib.ConvGeom) == 16 * 4


def test_product_path_fails_loudly_without_gpu(gs):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    conv = gs.DynamicConv2d(16, 16, 1)
    with pytest.raises(gs.GsError):
        conv(torch.randn(1, 16, 8, 8))
    bn = gs.DynamicBatchNorm2d(16)
    with pytest.raises(gs.GsError):
        bn(torch.randn(1, 16, 8, 8))
    with pytest.raises(gs.GsError):
        gs._lib.require_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'gaia_seg_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('CPU oracle', '').replace('the oracle', '').lower() or f == 'x', \
                    f'{f} mentions the oracle package'
