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


def test_psp_pool_and_concat_ops(gs):
    _assert_all(C.psp_op_checks(gs))


def test_full_size_properties(gs):
    _assert_all(C.full_size_checks(gs))


def test_maxpool(gs):
    _assert_all(C.maxpool_checks(gs))


def test_fused_upsample_ce(gs):
    _assert_all(C.loss_checks(gs))


def test_fused_upsample_argmax(gs):
    _assert_all(C.argmax_checks(gs))


def test_segmentor_train_and_eval_parity(gs):
    _assert_all(C.model_checks(gs))


def test_cuda_graph_replay_matches_eager(gs):
    """GraphedTrainStep (capture on the 3rd occurrence of a sub-net, replay afterwards) must train like the eager
    loop: same loss sequence (fp32 atomics make the two runs differ only at rounding level)."""
    import json
    cfg = C.small_cfg(aux=True)
    losses = {}
    for mode in ('eager', 'graph'):
        om, gm, _ = C.build_pair(gs, cfg, seed=1)
        opt = gs.GsSGD(gm, lr=0.05, momentum=0.9, weight_decay=5e-4)
        stepper = gs.GraphedTrainStep(gm, opt, graph_after=2, max_graphs=2 if mode == 'graph' else 0)
        if mode == 'eager':
            stepper.graph_after = 10 ** 9
        gm.train()
        seq = []
        for it in range(10):
            name = ('max', 'min')[it % 2]
            gm.manipulate_arch(C.SMALL_ARCHS[name])
            g = torch.Generator().manual_seed(100 + it)
            img = C.bf16r(torch.randn(2, 3, 64, 96, generator=g)).cuda()
            lab = C._labels(g, 2, 19, 64, 96).cuda()
            opt.param_groups[0]['lr'] = 0.05 * (1 - it / 10) ** 0.9          # LR schedule must reach the replay
            out = stepper(json.dumps(C.SMALL_ARCHS[name], sort_keys=True),
                          dict(img=img, img_metas=[{}, {}], gt_semantic_seg=lab))
            seq.append(float(out['log_vars']['loss']))
        losses[mode] = seq
        if mode == 'graph':
            assert len(stepper.graphs) == 2
    for a, b in zip(losses['eager'], losses['graph']):
        assert abs(a - b) <= 5e-3 * abs(a), (losses['eager'], losses['graph'])
    assert losses['eager'][-1] < losses['eager'][0]          # and it actually trains


def test_side_stream_weight_gradients_match_the_inline_order(gs):
    """Weight gradients run on a side stream during backward (one join per backward pass).  Same parameters, same batch:
    every gradient must equal the single-stream result up to the order of the fp32 split-K atomics."""
    from gaia_seg_b200 import functional as Fg
    cfg = C.small_cfg(aux=True, deep_stem=True, os8=True)
    grads = {}
    for mode in (True, False):
        om, gm, _ = C.build_pair(gs, cfg, seed=3)
        gm.manipulate_arch({'backbone': dict(C.SMALL_ARCHS['mid']['backbone'], stem={'width': [16, 16, 32]})})
        opt = gs.GsSGD(gm, lr=0.01, momentum=0.9)
        gm.train()
        g = torch.Generator().manual_seed(5)
        img = C.bf16r(torch.randn(2, 3, 64, 96, generator=g)).cuda()
        lab = C._labels(g, 2, 19, 64, 96).cuda()
        old, Fg.WGRAD_STREAM = Fg.WGRAD_STREAM, mode
        try:
            out = gm.train_step(dict(img=img, img_metas=[{}, {}], gt_semantic_seg=lab), opt)
            opt.zero_grad()
            out['loss'].backward()
            # no explicit synchronisation of the side stream here: the join installed by the backward pass must make
            # the CURRENT stream see every weight gradient
            grads[mode] = opt.flat.flat_g.clone().cpu()
        finally:
            Fg.WGRAD_STREAM = old
    a, b = grads[True], grads[False]
    assert float(a.abs().sum()) > 0
    err = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)
    assert err <= 1e-4, err


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_syncbn_and_gradient_allreduce_match_the_global_batch_oracle():
    """2 ranks x 2 images == the oracle on the 4-image batch (loss, gradients, running statistics), buffers
    and parameters bit-identical across ranks -- exercised through the peer-memory exchange / all-reduce kernels."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29571', os.path.join(root, 'tools', 'gpu_multi_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res['ok'] and res['buffers_identical_across_ranks'] and res['params_identical_after_step'], res
    v = res['vs_1rank']
    assert v['syncbn_allreduce_bit_exact'] and v['grad_allreduce_bit_exact'] and v['buffers_identical'], v
    assert v['loss_rel'] <= 1e-6 and v['layer_fwd_bit_exact'] and v['layer_dw_max_rel'] <= 1e-4, v


def test_no_cpu_fallback(gs):
    conv = gs.DynamicConv2d(16, 16, 1)
    with pytest.raises(gs.GsError):
        conv(torch.randn(1, 16, 8, 8))


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
