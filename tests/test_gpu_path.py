"""`-m gpu` parity tests, part 2: the steps either side of the conv / BN kernels (fused SGD, BN re-calibration modes,
Dropout2d, rescale) and VALUE parity at the full-size shapes of the benchmark.  Checks live in tests/gs_checks_path.py;
everything goes through the C ABI of libgaiaseg_b200.so."""
import pytest

import gs_checks_path as P

pytestmark = pytest.mark.gpu


def _assert_all(results):
    bad = [{k: v for k, v in r.items() if k != 'tb'} for r in results if not r['ok']]
    assert not bad, bad


@pytest.fixture(scope='module', autouse=True)
def _device(gs):
    gs._lib.require_device()


def test_fused_sgd_matches_torch_sgd(gs):
    _assert_all(P.sgd_checks(gs))


def test_bn_recalibration_modes(gs):
    _assert_all(P.bn_calibration_checks(gs))


def test_dropout2d_injected_mask(gs):
    _assert_all(P.dropout_checks(gs))


def test_fused_loss_full_size_fwd_bwd(gs):
    _assert_all(P.loss_full_size_checks(gs))


def test_rescale_two_step_resize(gs):
    _assert_all(P.rescale_checks(gs))


@pytest.mark.parametrize('case', P.BIG_CONV_CASES, ids=[c[0] for c in P.BIG_CONV_CASES])
def test_conv_values_at_benchmark_shapes(gs, case):
    _assert_all(P.big_conv_case_checks(case, gs))


def test_bn_backward_one_pass_and_two_kernels(gs):
    _assert_all(P.bn_bwd_one_pass_checks(gs))


def test_fcn_head_skip_gradient_in_dgrad_epilogue(gs):
    _assert_all(P.fcn_head_skip_gradient_checks(gs))


def test_conv_bn_one_launch_equals_two_kernels(gs):
    _assert_all(P.conv_bn_fused_launch_checks(gs))


def test_wide_tile_epilogue(gs):
    _assert_all(P.wide_tile_epilogue_checks(gs))


def test_full_depth_stage(gs):
    _assert_all(P.deep_stage_checks(gs))


def test_config3_r101_aspp_full_size(gs):
    _assert_all(P.config3_full_size_checks(gs))
