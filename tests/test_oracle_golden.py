"""The CPU oracle against the golden vectors produced by EXECUTING THE REFERENCE'S OWN CODE
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_model as O

HERE = os.path.dirname(os.path.abspath(__file__))
LOSS = np.load(os.path.join(HERE, 'golden', 'loss_golden.npz'))
WIRE = np.load(os.path.join(HERE, 'golden', 'wiring_golden.npz'))


def test_cross_entropy_and_accuracy_match_reference_code():
    for i in range(4):
        pred, lab = torch.from_numpy(LOSS[f'pred{i}']), torch.from_numpy(LOSS[f'lab{i}'])
        loss = O.cross_entropy(pred, lab, 255)
        acc = O.accuracy(pred, lab)
        assert np.array_equal(loss.numpy(), LOSS[f'loss{i}']), i          # bit-exact: same torch calls
        assert np.array_equal(acc.numpy(), LOSS[f'acc{i}']), i


def test_all_ignored_pixels_give_zero_loss():
    pred, lab = torch.from_numpy(LOSS['pred2']), torch.from_numpy(LOSS['lab2'])
    assert (lab == 255).all()
    assert float(LOSS['loss2']) == 0.0 and float(O.cross_entropy(pred, lab)) == 0.0
    assert float(LOSS['acc2'][0]) == 0.0


def test_accuracy_on_exact_ties_is_topk_defined():
    """accuracy() uses torch.topk(1): on EXACT ties its pick is implementation-defined (not the lowest index, and
    different between torch's CPU and CUDA kernels), so the reference itself has no portable answer there.  The
    oracle reproduces the reference (same call); the CUDA path and inference use arg-max = lowest index.  The two
    conventions can only differ on pixels whose top score is tied."""
    pred, lab = torch.from_numpy(LOSS['pred_tie']), torch.from_numpy(LOSS['lab_tie'])
    assert np.array_equal(O.accuracy(pred, lab).numpy(), LOSS['acc_tie'])
    top2 = pred.topk(2, dim=1).values
    tied = int((top2[:, 0] == top2[:, 1]).sum())
    hits_argmax = int((pred.argmax(1) == lab).sum())
    hits_topk = int(round(float(LOSS['acc_tie'][0]) * lab.numel() / 100.0))
    assert abs(hits_argmax - hits_topk) <= tied


def _build_oracle(case):
    import sys
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import make_golden as MG
    bb_kw, archs = MG.WIRING_CASES[case]
    bb = O.DynamicResNet(**bb_kw, **MG.CFG_COMMON)
    head = O.DynamicFCNHead(**MG.HEAD_KW, **MG.CFG_COMMON)
    MG.seeded_params(bb, 1)
    MG.seeded_params(head, 2)
    return MG, bb, head, archs


def test_wiring_parameter_names_match_reference_modules():
    for case in ('os32', 'os8_v1c'):
        _, bb, head, _ = _build_oracle(case)
        assert sorted(bb.state_dict().keys()) == list(WIRE[f'{case}.backbone_keys'])
        assert sorted(head.state_dict().keys()) == list(WIRE[f'{case}.head_keys'])


def test_wiring_outputs_match_reference_modules():
    for case in ('os32', 'os8_v1c'):
        MG, bb, head, archs = _build_oracle(case)
        g = torch.Generator().manual_seed(11)
        img = torch.randn(2, 3, 64, 64, generator=g)
        lab = torch.randint(0, 7, (2, 1, 64, 64), generator=g)
        lab[torch.rand(2, 1, 64, 64, generator=g) < 0.1] = 255
        for ai, arch in enumerate(archs):
            bb.manipulate_arch(arch)
            bb.train(); head.train()
            feats = bb(img)
            losses = head.forward_train(feats, None, lab)
            for fi, f in enumerate(feats):
                ref = WIRE[f'{case}.arch{ai}.feat{fi}']
                assert f.shape == ref.shape, (case, ai, fi)
                np.testing.assert_allclose(f.detach().numpy(), ref, rtol=0, atol=0)
            np.testing.assert_allclose(losses['loss_seg'].detach().numpy(), WIRE[f'{case}.arch{ai}.loss_seg'], rtol=0, atol=0)
            np.testing.assert_allclose(losses['acc_seg'].detach().numpy(), WIRE[f'{case}.arch{ai}.acc_seg'], rtol=0, atol=0)
            bb.eval(); head.eval()
            with torch.no_grad():
                np.testing.assert_allclose(head(bb(img)).numpy(), WIRE[f'{case}.arch{ai}.eval_logits'], rtol=0, atol=0)


def test_oracle_conv_slice_is_pure_prefix_and_bit_exact():
    conv = O.DynamicConv2d(12, 10, 3, padding=1)
    x = torch.randn(2, 7, 9, 9)
    conv.manipulate_width(6)
    ref = F.conv2d(x, conv.weight[:6, :7], conv.bias[:6], 1, 1)
    assert torch.equal(conv(x), ref)


def test_oracle_bn_updates_only_the_prefix_with_unbiased_variance():
    bn = O.DynamicBatchNorm2d(8)
    x = torch.randn(4, 5, 6, 6) * 3 + 1
    bn.train()
    y = bn(x)
    m = x.mean((0, 2, 3))
    v = x.var((0, 2, 3), unbiased=True)
    np.testing.assert_allclose(bn.running_mean[:5].numpy(), 0.1 * m.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bn.running_var[:5].numpy(), 0.9 + 0.1 * v.numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(bn.running_mean[5:], torch.zeros(3)) and torch.equal(bn.running_var[5:], torch.ones(3))
    np.testing.assert_allclose(y.mean((0, 2, 3)).detach().numpy(), np.zeros(5), atol=1e-5)


def test_syncbn_equals_global_batch_bn():
    """SyncBN over R ranks == the same BN on the concatenated batch (the multi-GPU oracle)."""
    bn = O.DynamicBatchNorm2d(6)
    xs = [torch.randn(2, 6, 5, 5) for _ in range(3)]
    full = bn(torch.cat(xs))
    s = sum(x.sum((0, 2, 3)) for x in xs)
    q = sum((x * x).sum((0, 2, 3)) for x in xs)
    n = sum(x.numel() // 6 for x in xs)
    mean, var = s / n, q / n - (s / n) ** 2
    ref0 = (xs[0] - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5)
    np.testing.assert_allclose(full[:2].detach().numpy(), ref0.numpy(), rtol=1e-4, atol=1e-5)


def test_extracted_subnet_equals_manipulated_supernet():
    """config 4: deploy() slices the parameters physically; logits must be bit-identical."""
    import copy
    cfg = dict(in_channels=3, stem_width=16, body_width=[16, 24, 32, 48], body_depth=[2, 2, 3, 2],
               conv_cfg=dict(type='DynConv2d'), norm_cfg=dict(type='DynBN', requires_grad=True))
    bb = O.DynamicResNet(**cfg)
    bb.init_weights()
    bb.eval()
    arch = {'stem': {'width': 8}, 'body': {'width': [8, 16, 24, 32], 'depth': [1, 2, 2, 1]}}
    bb.manipulate_arch(arch)
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        ref = bb(x)
        sub = copy.deepcopy(bb)
        sub.deploy()
        got = sub(x)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    assert sub.layer3[0].conv1.weight.shape[0] == 24 and len(sub.layer3) == 2
    assert sum(p.numel() for p in sub.parameters()) < sum(p.numel() for p in bb.parameters())


def test_aspp_head_equals_static_torch_modules_on_the_prefix_slice():
    """BASELINE config 3's DeepLabV3 head is not in the reference tree; the oracle's DynamicASPPHead is pinned against
    stock torch modules (nn.Conv2d / nn.BatchNorm2d holding the PREFIX-SLICED weights) wired like mmseg's ASPPHead."""
    torch.manual_seed(3)
    dil, C_in, C_x, ch, K = (1, 2, 3), 24, 16, 8, 5
    head = O.DynamicASPPHead(C_in, ch, K, dilations=dil, dropout_ratio=0, conv_cfg=dict(type='DynConv2d'),
                             norm_cfg=dict(type='DynBN', requires_grad=True), in_index=0)
    for p in head.parameters():
        torch.nn.init.normal_(p, 0, 0.3)
    head.eval()
    x = torch.randn(2, C_x, 9, 11)                                       # narrower than in_channels: prefix slice

    def cba(mod, inp, k, d):
        w = mod.conv.weight[:, :inp.size(1)]
        y = F.conv2d(inp, w, None, 1, 0 if k == 1 else d, d)
        n = mod.norm
        return F.relu(F.batch_norm(y, n.running_mean, n.running_var, n.weight, n.bias, False, 0.1, n.eps))

    outs = [F.interpolate(cba(head.image_pool[1], F.adaptive_avg_pool2d(x, 1), 1, 1), size=(9, 11), mode='bilinear',
                          align_corners=False)]
    outs += [cba(m, x, 1 if d == 1 else 3, d) for m, d in zip(head.aspp_modules, dil)]
    ref = F.conv2d(cba(head.bottleneck, torch.cat(outs, 1), 3, 1), head.conv_seg.weight, head.conv_seg.bias)
    with torch.no_grad():
        got = head([x])
    np.testing.assert_allclose(got.numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert list(dict(head.named_parameters()))[:3] == ['conv_seg.weight', 'conv_seg.bias', 'image_pool.1.conv.weight']
