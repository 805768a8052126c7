"""`-m gpu` parity tests: CUDA path (through the C ABI of libgaiaseg_b200.so) vs the CPU oracle on the same
seeded inputs.  Tolerances are stated in tests/gs_checks.py next to each comparison."""
import pytest
import torch

import gs_checks as C

pytestmark = pytest.mark.gpu


def _assert_all(results):
    bad = [{k: v for k, v in r.items() if k != 'tb'} for r in results if not r['ok']]
    assert not bad, bad


@pytest.fixture(scope='module', autouse=True)
def _device(gs):
    gs._lib.require_device()


@pytest.mark.parametrize('case', C.CONV_CASES, ids=[c[0] for c in C.CONV_CASES])
def test_conv_tcgen05_fwd_dgrad_wgrad(gs, case):
    _assert_all(C.conv_case_checks(case, gs, 'tc'))


@pytest.mark.parametrize('case', C.CONV_CASES[:4], ids=[c[0] for c in C.CONV_CASES[:4]])
def test_conv_simt_twin(gs, case):
    _assert_all(C.conv_case_checks(case, gs, 'simt'))


def test_conv_epilogue(gs):
    _assert_all(C.conv_epilogue_checks(gs))


def test_image_conv(gs):
    _assert_all(C.image_conv_checks(gs))


def test_conv_bn_act_train_eval(gs):
    _assert_all(C.bn_checks(gs))


def test_standalone_dynbn(gs):
    _assert_all(C.standalone_bn_checks(gs))


def test_res_stage_fwd_bwd(gs):
    _assert_all(C.stage_checks(gs))


def test_maxpool(gs):
    _assert_all(C.maxpool_checks(gs))


def test_fused_upsample_ce(gs):
    _assert_all(C.loss_checks(gs))


def test_fused_upsample_argmax(gs):
    _assert_all(C.argmax_checks(gs))


def test_segmentor_train_and_eval_parity(gs):
    _assert_all(C.model_checks(gs))


def test_no_cpu_fallback(gs):
    conv = gs.DynamicConv2d(16, 16, 1)
    with pytest.raises(gs.GsError):
        conv(torch.randn(1, 16, 8, 8))


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
