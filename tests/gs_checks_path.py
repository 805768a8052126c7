"""Parity checks for the steps either side of the conv/BN kernels (optimizer, BN re-calibration modes, dropout,
two-step rescale) and for the FULL-SIZE shapes of the benchmark (the layers that cost the time), CUDA path vs the CPU
oracle.  Same conventions as gs_checks.py: every check returns dict(name, ok, err, tol, ...).

Reference semantics pinned here:
  * SGD(0.01, 0.9, 5e-4) + poly LR            configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:175-178
  * caliberate_bn.use_minibatch_stats          tools/test_supernet.py:190-198 (running stats dropped -> batch stats in eval)
  * caliberate_bn.reset_stats                  gaiaseg/apis/train.py:177-184 (running_mean = 0, running_var = 1)
  * multi_gpu_test leaves the model in train   gaiaseg/apis/test.py:85
  * Dropout2d before conv_seg                  gaiaseg/models/decode_heads/fcn_head.py:248-253
  * rescale to ori_shape (second resize)       gaiaseg/models/segmentors/dynamic_distiller.py:461-521
"""
import math
import time

import torch
import torch.nn.functional as F

import gs_checks as C
from gs_checks import O, bf16r, check_bf16, check_f32, rel_err


# ------------------------------------------------------------------------------------------------
# fused SGD over the flat buffer vs torch.optim.SGD
# ------------------------------------------------------------------------------------------------
def sgd_checks(gs, steps=6):
    """gs_sgd_flat == torch.optim.SGD(momentum 0.9, wd 5e-4, dampening 0) with the poly LR schedule, fed with the SAME
    gradients (read back from the CUDA flat gradient buffer), over `steps` iterations that sample different sub-nets:
    step 0 is MAX (every block visited), later steps leave a whole block (layer3.2, layer1.1 ...) and channel slices
    unused.  The reference semantics (SURVEY 8c hazard 5): after the first visit `zero_grad()` leaves ZERO gradients
    (torch 1.9 set_to_none=False), so unused slices AND unused blocks keep decaying and keep their momentum -- the
    torch optimizer below therefore receives zero tensors for them, and the test asserts that they really moved.
    Also: first-step momentum (buf = d), momentum buffers, and flat_shadow == bf16(flat_p) bit-exactly."""
    out = []
    cfg = C.small_cfg(aux=True)
    om, gm, _ = C.build_pair(gs, cfg, seed=4)
    lr0, mom, wd, power = 0.01, 0.9, 5e-4, 0.9
    opt = gs.GsSGD(gm, lr=lr0, momentum=mom, weight_decay=wd)
    names = [n for n, p in gm.named_parameters() if p.requires_grad]
    ref = {n: torch.nn.Parameter(p.detach().cpu().clone()) for n, p in gm.named_parameters() if p.requires_grad}
    ropt = torch.optim.SGD([ref[n] for n in names], lr=lr0, momentum=mom, weight_decay=wd)
    gm.train()
    seq = ['max', 'min', 'mid', 'min', 'max', 'mid'][:steps]
    unused_moved, worst_p, worst_m = True, 0.0, 0.0
    shadow_exact = True
    for it, aname in enumerate(seq):
        lr = lr0 * (1 - it / steps) ** power
        opt.param_groups[0]['lr'] = lr
        ropt.param_groups[0]['lr'] = lr
        gm.manipulate_arch(C.SMALL_ARCHS[aname])
        g = torch.Generator().manual_seed(40 + it)
        img = bf16r(torch.randn(2, 3, 64, 96, generator=g)).cuda()
        lab = C._labels(g, 2, 19, 64, 96).cuda()
        res = gm.train_step(dict(img=img, img_metas=[{}, {}], gt_semantic_seg=lab), opt)
        opt.zero_grad()
        res['loss'].backward()
        gm_params = dict(gm.named_parameters())
        before = {n: gm_params[n].detach().cpu().clone() for n in names}
        zero_grad_names = []
        for n in names:
            gq = gm_params[n].grad.detach().cpu().clone()
            ref[n].grad = gq
            if float(gq.abs().max()) == 0.0:
                zero_grad_names.append(n)
        opt.step()
        ropt.step()
        torch.cuda.synchronize()
        for n in names:
            a, b = gm_params[n].detach().cpu(), ref[n].detach()
            e = float((a - b).abs().max() / (b.abs().max() + 1e-12))
            worst_p = max(worst_p, e)
        # momentum buffers, through the flat offsets
        for p_, off, n in zip(opt.flat.params, opt.flat.offsets, names):
            mb = opt.momentum_buf[off:off + p_.numel()]
            if p_.dim() == 4:
                Co, Ci, kh, kw = p_.shape
                mb = mb.view(Co, kh, kw, Ci).permute(0, 3, 1, 2)
            else:
                mb = mb.view(p_.shape)
            rb = ropt.state[ref[n]]['momentum_buffer']
            worst_m = max(worst_m, float((mb.cpu() - rb).abs().max() / (rb.abs().max() + 1e-12)))
        if it > 0 and zero_grad_names:
            # parameters with an all-zero gradient (blocks / layers outside this sub-net) still decay and coast on momentum
            moved = [float((gm_params[n].detach().cpu() - before[n]).abs().max()) > 0 for n in zero_grad_names
                     if float(before[n].abs().max()) > 0]
            unused_moved &= bool(moved) and all(moved)
        shadow_exact &= bool(torch.equal(opt.flat.flat_shadow, opt.flat.flat_p.to(torch.bfloat16)))
        if it == 0:
            # first step: momentum buffer == d = g + wd * p  (torch: buf = clone(d))
            n0 = names[0]
            d0 = ref[n0].grad + wd * before[n0]
            rb = ropt.state[ref[n0]]['momentum_buffer']
            out.append(dict(name='sgd.first_step_momentum_is_d', ok=bool(torch.allclose(rb, d0, rtol=1e-6, atol=1e-9)),
                            err=float((rb - d0).abs().max()), tol=1e-6))
    out.append(dict(name=f'sgd.params_vs_torch_SGD[{steps} steps, poly lr, MAX/MIN/MID]', ok=worst_p <= 2e-6, err=worst_p,
                    tol=2e-6))
    out.append(dict(name='sgd.momentum_buffers_vs_torch_SGD', ok=worst_m <= 2e-6, err=worst_m, tol=2e-6))
    out.append(dict(name='sgd.zero_grad_params_still_decay', ok=bool(unused_moved), err=0.0, tol=0))
    out.append(dict(name='sgd.flat_shadow_equals_bf16_of_master_bit_exact', ok=bool(shadow_exact), err=0.0, tol=0))
    # torch-format optimizer state round trip (reference checkpoints: state[i]['momentum_buffer'] in OIHW + param_groups)
    sd = opt.state_dict()
    ok_fmt = isinstance(sd.get('state'), dict) and 'param_groups' in sd and 'params' in sd['param_groups'][0]
    if ok_fmt:
        i0 = sd['param_groups'][0]['params'][0]
        ok_fmt = tuple(sd['state'][i0]['momentum_buffer'].shape) == tuple(opt.flat.params[0].shape)
        rsd = ropt.state_dict()
        e = max(float((sd['state'][i]['momentum_buffer'].cpu() - rsd['state'][i]['momentum_buffer']).abs().max())
                for i in rsd['state'])
        ok_fmt = ok_fmt and e <= 1e-6
        mb0 = opt.momentum_buf.clone()
        opt.momentum_buf.zero_()
        opt.load_state_dict(rsd)            # a torch.optim.SGD state dict (what a reference checkpoint carries)
        ok_fmt = ok_fmt and float((opt.momentum_buf - mb0).abs().max()) <= 1e-6
    out.append(dict(name='sgd.state_dict_is_torch_SGD_format_and_loads_reference_state', ok=bool(ok_fmt), err=0.0, tol=0))
    return out


# ------------------------------------------------------------------------------------------------
# BN re-calibration modes of the evaluation path
# ------------------------------------------------------------------------------------------------
def _drop_running_stats(model):
    from torch.nn.modules.batchnorm import _BatchNorm
    for m in model.modules():
        if isinstance(m, _BatchNorm):
            m.running_mean = None
            m.running_var = None
            m.track_running_stats = False


def _reset_running_stats(model):
    from torch.nn.modules.batchnorm import _BatchNorm
    for m in model.modules():
        if isinstance(m, _BatchNorm):
            m.running_mean.zero_()
            m.running_var.fill_(1)


def rel_l2(got, ref, tol, name):
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    err = float((got - ref).norm() / (ref.norm() + 1e-300)) if torch.isfinite(got).all() else float('inf')
    return dict(name=name, ok=err <= tol, err=err, tol=tol)


def bn_calibration_checks(gs):
    """(a) use_minibatch_stats: eval() with the running statistics dropped normalises with the statistics of the test
    batch; (b) reset_stats followed by train-mode forward passes (multi_gpu_test does not call eval()) re-estimates the
    running statistics; (c) eval() afterwards uses them.  Low-res logits vs the fp32 oracle as a relative L2 error over
    the whole logit tensor <= 5e-2 (bf16 activation storage through ~35 layers; an ELEMENT-wise bound is meaningless here:
    batch statistics on 8 x 12 maps with random weights leave channels with variance ~1e-6 whose 1 / std amplifies one bf16
    rounding to O(1) in a few logits) and label-map agreement >= 97 %."""
    import numpy as np
    out = []
    cfg = C.small_cfg(aux=False, deep_stem=True, os8=True)
    arch = {'backbone': {'stem': {'width': [8, 8, 16]}, 'body': {'width': [16, 48, 48, 80], 'depth': [2, 1, 3, 1]}}}
    g = torch.Generator().manual_seed(77)
    img = bf16r(torch.randn(2, 3, 64, 96, generator=g))
    metas = [[dict(ori_shape=(64, 96, 3), flip=False)] * 2]

    def lowres(om, gm):
        with torch.no_grad():
            lo = om.decode_head(om.backbone(img))
            lg = gm.encode_decode_lowres(img.cuda(), metas[0]).float().cpu()
        return lo, lg

    # (a) batch statistics on 8 x 12 maps amplify bf16 STORAGE rounding (channels with variance ~1e-6 get 1 / std ~ 1e3), so
    # the reference here is the oracle with bf16 storage emulated at the CUDA path's rounding points (fp64 arithmetic);
    # the plain fp32 oracle's distance is reported next to it
    om, gm, _ = C.build_pair(gs, cfg, seed=6)
    om.manipulate_arch(arch); gm.manipulate_arch(arch)
    _drop_running_stats(om); _drop_running_stats(gm)
    om.eval(); gm.eval()
    lo32, lg = lowres(om, gm)
    def storage_oracle(dtype):
        ome = O.build_segmentor(cfg)
        ome.load_state_dict(om.state_dict(), strict=False)
        ome.manipulate_arch(arch)
        _drop_running_stats(ome)
        C.emulate_bf16_storage(ome)
        ome = ome.to(dtype).eval()
        with torch.no_grad():
            return ome.decode_head(ome.backbone(img.to(dtype))).double(), ome.simple_test(img.to(dtype))

    lo, po = storage_oracle(torch.float64)
    lo_s32, po_s32 = storage_oracle(torch.float32)
    # noise floor of this ill-conditioned case: the SAME oracle (same bf16 rounding points) in fp32 vs fp64 arithmetic
    floor = rel_l2(lo_s32, lo, 1.0, '')['err']
    floor_agree = float((po_s32 == po).float().mean())
    tol = 4 * floor + 2e-2
    r = rel_l2(lg, lo, tol, 'bn_calib.use_minibatch_stats.eval_logits_rel_l2_vs_bf16_storage_oracle')
    r.update(oracle_fp32_vs_fp64_same_rounding_points=floor, rel_l2_vs_plain_fp32_oracle=rel_l2(lg, lo32, 1.0, '')['err'],
             plain_fp32_vs_bf16_storage_oracle=rel_l2(lo32, lo, 1.0, '')['err'])
    out.append(r)
    with torch.no_grad():
        pg = gm(return_loss=False, img=[img.cuda()], img_metas=metas)
    agree = float((torch.from_numpy(np.stack(pg)) == po).float().mean())
    need = min(0.97, floor_agree - 0.03)
    out.append(dict(name='bn_calib.use_minibatch_stats.labelmap_agreement', ok=agree >= need, err=1 - agree, tol=1 - need,
                    oracle_fp32_vs_fp64_agreement=floor_agree))
    # the batch-stat eval result must differ from the running-stat eval result (otherwise the mode is not exercised)
    om2, gm2, _ = C.build_pair(gs, cfg, seed=6)
    om2.manipulate_arch(arch); gm2.manipulate_arch(arch)
    om2.eval(); gm2.eval()
    _, lg_run = lowres(om2, gm2)
    out.append(dict(name='bn_calib.use_minibatch_stats.differs_from_running_stat_eval',
                    ok=float((lg_run - lg).abs().max()) > 1e-2, err=float((lg_run - lg).abs().max()), tol=1e-2))
    # (b) + (c)
    om, gm, _ = C.build_pair(gs, cfg, seed=6)
    om.manipulate_arch(arch); gm.manipulate_arch(arch)
    with torch.no_grad():
        for m in list(om.modules()) + list(gm.modules()):
            if hasattr(m, 'running_mean') and m.running_mean is not None:
                m.running_mean.fill_(3.0)
                m.running_var.fill_(7.0)
    _reset_running_stats(om); _reset_running_stats(gm)
    rm0 = [b for n, b in gm.named_buffers() if n.endswith('running_mean')]
    rv0 = [b for n, b in gm.named_buffers() if n.endswith('running_var')]
    ok_reset = all(float(b.abs().max()) == 0.0 for b in rm0) and all(float((b - 1).abs().max()) == 0.0 for b in rv0)
    out.append(dict(name='bn_calib.reset_stats.zero_mean_unit_var', ok=bool(ok_reset), err=0.0, tol=0))
    om.train(); gm.train()            # multi_gpu_test: model left in train mode -> forward passes update the statistics
    for k in range(3):
        gk = torch.Generator().manual_seed(200 + k)
        imk = bf16r(torch.randn(2, 3, 64, 96, generator=gk))
        with torch.no_grad():
            om.decode_head(om.backbone(imk))
            gm(return_loss=False, img=[imk.cuda()], img_metas=metas)
    torch.cuda.synchronize()
    bo = dict(om.named_buffers())
    e_rm = max(float((b.cpu() - bo[n]).abs().max()) for n, b in gm.named_buffers() if n.endswith('running_mean'))
    e_rv = max(float(((b.cpu() - bo[n]).abs() / (bo[n].abs() + 1e-3)).max()) for n, b in gm.named_buffers()
               if n.endswith('running_var'))
    out.append(dict(name='bn_calib.recalibrated_running_mean_vs_oracle', ok=e_rm <= 2e-2, err=e_rm, tol=2e-2))
    out.append(dict(name='bn_calib.recalibrated_running_var_vs_oracle', ok=e_rv <= 5e-2, err=e_rv, tol=5e-2))
    nbt = [int(b) for n, b in gm.state_dict().items() if n.endswith('num_batches_tracked')]
    out.append(dict(name='bn_calib.num_batches_tracked_counts_the_passes', ok=bool(nbt) and all(v == 3 for v in nbt
                    if v > 0) and any(v == 3 for v in nbt), err=0.0, tol=0))
    om.eval(); gm.eval()
    lo, lg = lowres(om, gm)
    out.append(rel_l2(lg, lo, 5e-2, 'bn_calib.eval_after_recalibration_logits_rel_l2_vs_fp32_oracle'))
    with torch.no_grad():
        po = om.simple_test(img)
        pg = gm(return_loss=False, img=[img.cuda()], img_metas=metas)
    agree = float((torch.from_numpy(np.stack(pg)) == po).float().mean())
    out.append(dict(name='bn_calib.eval_after_recalibration.labelmap_agreement', ok=agree >= 0.97, err=1 - agree, tol=0.03))
    return out


# ------------------------------------------------------------------------------------------------
# Dropout2d with an injected mask
# ------------------------------------------------------------------------------------------------
def dropout_checks(gs):
    """(a) Dropout2dFn forward / backward with an injected [N, C] mask: y = bf16(x * m), dx = bf16(dy * m), bit-exact;
    whole channels are zeroed.  (b) the product mask (torch generator): values in {0, 1/keep}, per-(n, c) constant,
    drop rate near p.  (c) FCN head in train mode with dropout 0.1 and the SAME mask injected on both sides vs the
    oracle head: loss 2e-2 (whole-net bf16 tolerance)."""
    Fg = gs.functional
    dev = torch.device('cuda')
    out = []
    g = torch.Generator().manual_seed(12)
    N, Cc, H, W, p = 2, 48, 10, 14, 0.25
    keep = 1 - p
    mask = (torch.rand(N, Cc, generator=g) < keep).float() / keep
    x = bf16r(torch.randn(N, Cc, H, W, generator=g))
    dy = bf16r(torch.randn(N, Cc, H, W, generator=g))
    Fg.DROPOUT_MASK_FN = lambda n, c, k, d: mask.to(d)
    try:
        xg = Fg.as_act(x.to(dev)).requires_grad_(True)
        yg = Fg.dropout2d(xg, p, True)
        yg.backward(Fg.as_act(dy.to(dev)))
        torch.cuda.synchronize()
    finally:
        Fg.DROPOUT_MASK_FN = None
    y_ref = bf16r(x * mask[:, :, None, None])
    dx_ref = bf16r(dy * mask[:, :, None, None])
    out.append(dict(name='dropout2d.fwd_injected_mask_bit_exact', ok=bool(torch.equal(yg.float().cpu(), y_ref)),
                    err=float((yg.float().cpu() - y_ref).abs().max()), tol=0))
    out.append(dict(name='dropout2d.bwd_injected_mask_bit_exact', ok=bool(torch.equal(xg.grad.float().cpu(), dx_ref)),
                    err=float((xg.grad.float().cpu() - dx_ref).abs().max()), tol=0))
    # (b) product mask
    torch.manual_seed(3)
    xb = Fg.as_act(torch.ones(8, 256, 4, 4).to(dev))
    yb = Fg.dropout2d(xb, 0.1, True).float().cpu()
    per = yb.flatten(2)
    const = bool((per.max(-1).values == per.min(-1).values).all())
    vals = set(round(v, 3) for v in per[:, :, 0].flatten().tolist())
    rate = float((per[:, :, 0] == 0).float().mean())
    out.append(dict(name='dropout2d.product_mask_channelwise_values_rate', ok=const and vals <= {0.0, round(float(bf16r(torch.tensor(1 / 0.9))), 3)}
                    and 0.05 <= rate <= 0.16, err=rate, tol=0.1, values=sorted(vals)))
    # eval mode: identity (same tensor)
    out.append(dict(name='dropout2d.eval_is_identity', ok=Fg.dropout2d(xb, 0.1, False) is xb, err=0.0, tol=0))
    # (c) head level
    cfg = C.small_cfg(dropout=0.1)
    om, gm, _ = C.build_pair(gs, cfg, seed=8)
    arch = C.SMALL_ARCHS['mid']
    om.manipulate_arch(arch); gm.manipulate_arch(arch)
    img = bf16r(torch.randn(2, 3, 64, 96, generator=g))
    lab = C._labels(g, 2, 19, 64, 96)
    hm = (torch.rand(2, 64, generator=g) < 0.9).float() / 0.9

    class _InjectedDropout(torch.nn.Module):
        def forward(self, t):
            return t * hm[:, :, None, None].to(t.dtype) if self.training else t

    om.decode_head.dropout = _InjectedDropout()
    om.train(); gm.train()
    Fg.DROPOUT_MASK_FN = lambda n, c, k, d: hm.to(d)
    try:
        with torch.no_grad():
            lo = om.parse_losses(om.forward_train(img, None, lab))
            res = gm.train_step(dict(img=img.cuda(), img_metas=[{}, {}], gt_semantic_seg=lab.cuda()), None)
        # and the mask must matter: same model without dropout gives another loss
        Fg.DROPOUT_MASK_FN = lambda n, c, k, d: torch.ones(n, c, device=d)
        with torch.no_grad():
            res1 = gm.train_step(dict(img=img.cuda(), img_metas=[{}, {}], gt_semantic_seg=lab.cuda()), None)
    finally:
        Fg.DROPOUT_MASK_FN = None
    out.append(check_f32(res['loss'].reshape(1).cpu(), lo.reshape(1), 'dropout2d.fcn_head_train_loss_same_mask_vs_oracle', 2e-2))
    d = abs(float(res['loss']) - float(res1['loss']))
    out.append(dict(name='dropout2d.mask_changes_the_loss', ok=d > 1e-4, err=d, tol=1e-4))
    return out


# ------------------------------------------------------------------------------------------------
# rescale (two-step resize) of the inference path
# ------------------------------------------------------------------------------------------------
def rescale_checks(gs):
    """simple_test(rescale=True) with ori_shape != input size: logits -> bilinear to the input size -> bilinear to
    ori_shape -> argmax (the reference resizes twice).  Exact outside top-2 margins < 1e-4 of the oracle's logits
    (the oracle runs the net in fp32, the CUDA path in bf16 -> compare on the SAME low-res logits)."""
    import numpy as np
    Fg = gs.functional
    out = []
    g = torch.Generator().manual_seed(31)
    for (K, h, w, Hin, Win, Ho, Wo) in ((19, 8, 12, 64, 96, 50, 75), (150, 8, 8, 64, 64, 97, 131)):
        logits = torch.randn(1, K, h, w, generator=g) * 2
        up1 = F.interpolate(logits, size=(Hin, Win), mode='bilinear', align_corners=False)
        up2 = F.interpolate(up1, size=(Ho, Wo), mode='bilinear', align_corners=False)
        ref = F.softmax(up2, dim=1).argmax(dim=1)
        lg = logits.cuda().contiguous(memory_format=torch.channels_last)
        mid = Fg.upsample_bilinear_f32(lg, (Hin, Win))
        got = Fg.upsample_argmax(mid, (Ho, Wo)).cpu()
        out.append(check_f32(mid.cpu(), up1, f'rescale[K{K}].first_resize_f32', 1e-5))
        top2 = up2.topk(2, dim=1).values
        fragile = (top2[:, 0] - top2[:, 1]).abs() < 1e-4
        bad = (got != ref) & ~fragile
        out.append(dict(name=f'rescale[K{K},{h}x{w}->{Hin}x{Win}->{Ho}x{Wo}].labels', ok=int(bad.sum()) == 0,
                        err=int(bad.sum()), tol=0, at_ties=int(((got != ref) & fragile).sum())))
    # through the public API: ori_shape in the metas drives the second resize, output shape == ori_shape
    cfg = C.small_cfg()
    om, gm, _ = C.build_pair(gs, cfg, seed=2)
    om.eval(); gm.eval()
    img = bf16r(torch.randn(1, 3, 64, 96, generator=g))
    with torch.no_grad():
        pg = gm(return_loss=False, img=[img.cuda()], img_metas=[[dict(ori_shape=(50, 75, 3), flip=False)]])
        po = om.simple_test(img, ori_shape=(50, 75))
    agree = float((torch.from_numpy(np.stack(pg)) == po).float().mean())
    out.append(dict(name='rescale.simple_test_ori_shape', ok=pg[0].shape == (50, 75) and agree >= 0.97, err=1 - agree,
                    tol=0.03, shape=list(pg[0].shape)))
    return out


# ------------------------------------------------------------------------------------------------
# the conv shapes that cost the time (full size) -- VALUES, not only properties
# ------------------------------------------------------------------------------------------------
BIG_CONV_CASES = [
    # name,                N, H,   W,   Ci,   Co,  k, s, p, d, Ci_max, Co_max   (P = 16384 unless noted)
    ('s3_1x1_320_1280',    2, 64, 128, 320, 1280, 1, 1, 0, 1, 320, 1280),
    ('s3_1x1_1280_320',    2, 64, 128, 1280, 320, 1, 1, 0, 1, 1280, 320),
    ('s3_3x3_320_d2',      2, 64, 128, 320, 320, 3, 1, 2, 2, 320, 320),
    ('s3_1x1_256_1024',    2, 64, 128, 256, 1024, 1, 1, 0, 1, 320, 1280),      # prefix slice of the supernet weight
    ('s3_1x1_1024_256',    2, 64, 128, 1024, 256, 1, 1, 0, 1, 1280, 320),
    ('s3_3x3_256_d2',      2, 64, 128, 256, 256, 3, 1, 2, 2, 320, 320),
    ('s4_3x3_640_d4',      2, 64, 128, 640, 640, 3, 1, 4, 4, 640, 640),
    ('s4_1x1_640_2560',    2, 64, 128, 640, 2560, 1, 1, 0, 1, 640, 2560),
    ('s4_1x1_2560_640',    2, 64, 128, 2560, 640, 1, 1, 0, 1, 2560, 640),
    ('head_3x3_2560_512',  2, 64, 128, 2560, 512, 3, 1, 1, 1, 2560, 512),
    ('head_cat_3072_512',  2, 64, 128, 3072, 512, 3, 1, 1, 1, 3072, 512),
    ('s1_1x1_320_80',      2, 128, 256, 320, 80, 1, 1, 0, 1, 320, 80),          # P = 65536
    ('s1_3x3_80',          2, 128, 256, 80, 80, 3, 1, 1, 1, 80, 80),
    ('s1_1x1_80_320',      2, 128, 256, 80, 320, 1, 1, 0, 1, 80, 320),
    ('s2_3x3_160_s2',      2, 128, 256, 160, 160, 3, 2, 1, 1, 160, 160),        # stride-2 (zero-inserted dgrad)
    ('s2_ds_1x1_320_640',  2, 128, 256, 320, 640, 1, 2, 0, 1, 320, 640),
    ('stem_3x3_32_64',     2, 256, 512, 32, 64, 3, 1, 1, 1, 32, 64),            # P = 262144
    # wide n-tiles (one 320 / 384-column accumulator, two MMAs per K step): 384 = stage-4 width 384, prefix slices, ragged
    ('s4_3x3_384_d4',      2, 64, 128, 384, 384, 3, 1, 4, 4, 640, 640),
    ('s4_1x1_1536_384',    2, 64, 128, 1536, 384, 1, 1, 0, 1, 2560, 640),
    ('s4_1x1_384_1536',    2, 64, 128, 384, 1536, 1, 1, 0, 1, 640, 2560),
    ('wide_ragged_640',    3, 37, 53, 192, 640, 3, 1, 1, 1, 192, 640),
]


def wide_tile_epilogue_checks(gs):
    """The wide n-tile path (Cout = 320 in ONE accumulator, second drain round of every epilogue group) with the FULL
    epilogue: per-channel scale / shift, residual (TMA-loaded into the staging slot), ReLU, and the DynBN statistics."""
    Fg = gs.functional
    dev = torch.device('cuda')
    out = []
    for Co in (320, 384):
        case = (f'wide_epi{Co}', 2, 24, 40, 128, Co, 3, 1, 1, 1, 160, Co)
        conv, x, g = C._mk_conv(case, dev, gs)
        scale = torch.rand(Co, generator=g) + 0.5
        shift = torch.randn(Co, generator=g)
        res = bf16r(torch.randn(2, Co, 24, 40, generator=g))
        ref = F.conv2d(x.double(), conv.weight.detach().cpu().double()[:Co, :128], None, 1, 1, 1)
        ref = torch.relu(ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1) + res.double())
        z, _, _, _ = Fg.conv_forward(Fg.as_act(x.to(dev)), conv, Co, scale=scale.to(dev), shift=shift.to(dev),
                                     residual=Fg.as_act(res.to(dev)), relu=True)
        torch.cuda.synchronize()
        out.append(check_bf16(z.float(), ref, f'wide_tile[{Co}].scale_shift_res_relu'))
        y, stats, _, _ = Fg.conv_forward(Fg.as_act(x.to(dev)), conv, Co, want_stats=True)
        torch.cuda.synchronize()
        yr = y.float().double().cpu()
        out.append(check_f32(stats, torch.cat([yr.sum((0, 2, 3)), (yr * yr).sum((0, 2, 3))]), f'wide_tile[{Co}].stats', 1e-4))
    return out



def big_conv_case_checks(case, gs):
    """fwd (+ epilogue statistics) / dgrad / wgrad of one FULL-SIZE geometry against fp32 F.conv2d on the host
    (fp64 would take minutes per case; fp32 accumulation error ~1e-6 is far below the bf16 output tolerance)."""
    Fg = gs.functional
    name, N, H, W, Ci, Co, k, s, p, d, Ci_max, Co_max = case
    tag = f'bigconv[{name}]'
    dev = torch.device('cuda')
    t0 = time.time()
    conv, x, g = C._mk_conv(case, dev, gs)
    out = []
    xd = x.clone().requires_grad_(True)
    wd = conv.weight.detach().cpu()[:Co, :Ci].clone().contiguous().requires_grad_(True)
    ref = F.conv2d(xd, wd, None, s, p, d)
    dy = bf16r(torch.randn(ref.shape, generator=g))
    ref.backward(dy)
    xa = Fg.as_act(x.to(dev))
    y, stats, a, geom = Fg.conv_forward(xa, conv, Co, want_stats=True)
    torch.cuda.synchronize()
    out.append(check_bf16(y.float(), ref.detach(), tag + '.fwd'))
    yr = y.float().double()
    st_ref = torch.cat([yr.sum((0, 2, 3)), (yr * yr).sum((0, 2, 3))]).cpu()
    out.append(check_f32(stats, st_ref, tag + '.stats', 1e-4))
    dya = Fg.as_act(dy.to(dev))
    dx = Fg.conv_dgrad(conv, dya, geom, tuple(x.shape))
    torch.cuda.synchronize()
    out.append(check_bf16(dx.float(), xd.grad, tag + '.dgrad'))
    conv.weight.grad = None
    Fg.conv_wgrad(conv, a, dya, geom)
    torch.cuda.synchronize()
    gw = conv.weight.grad.detach().cpu()
    out.append(check_f32(gw[:Co, :Ci], wd.grad, tag + '.wgrad', 2e-3))
    rest = gw.clone()
    rest[:Co, :Ci] = 0
    out.append(dict(name=tag + '.wgrad_outside_slice_zero', ok=bool((rest == 0).all()), err=float(rest.abs().max()), tol=0))
    out[-1]['seconds'] = round(time.time() - t0, 1)
    return out


# ------------------------------------------------------------------------------------------------
# full-depth stage (29 blocks) -- between the 2-block stage test and the whole net
# ------------------------------------------------------------------------------------------------
def deep_stage_checks(gs, depth=29):
    """Stage 3 of the supernet at FULL DEPTH (29 bottlenecks, dilation 2, contract_dilation) on a small map and a
    channel-prefix slice, forward / input gradient / every parameter gradient against the storage-emulating fp64 oracle.
    The chain is 87 conv+BN layers deep, so a single bf16 rounding flip early on moves later values by several ulp: the
    tolerances are calibrated on the oracle's own fp32-vs-fp64 disagreement (same rounding points, 1e-7 differences):
    CUDA-vs-oracle64 must stay within 4x that noise floor + a tight absolute term."""
    Fg = gs.functional
    dev = torch.device('cuda')
    out = []
    g = torch.Generator().manual_seed(29)
    kw = dict(inplanes=96, planes=48, depth=depth, stride=1, dilation=2, contract_dilation=True,
              conv_cfg=dict(type='DynConv2d'), norm_cfg=dict(type='DynBN', requires_grad=True), style='pytorch')
    w_act, cin = 32, 64

    def oracle(dtype):
        ol = O.DynamicResLayer(block=O.DynamicBottleneck, **kw)
        C.randomize(ol, 9)
        sd = {k_: v.clone() for k_, v in ol.state_dict().items()}
        ol.manipulate_arch({'width': w_act, 'depth': depth})
        ol.train()
        C.emulate_bf16_storage(ol)
        ol = ol.to(dtype)
        xo = x.clone().to(dtype).requires_grad_(True)
        zo = ol(xo)
        zo.backward(dz.to(dtype))
        return sd, zo.detach().double(), xo.grad.double(), {n: p.grad.detach().double() for n, p in ol.named_parameters()
                                                           if p.grad is not None}

    x = bf16r(torch.randn(2, cin, 12, 16, generator=g))
    dz = bf16r(torch.randn(2, w_act * 4, 12, 16, generator=g))
    sd, z64, dx64, g64 = oracle(torch.float64)
    _, z32, dx32, g32 = oracle(torch.float32)
    gl = gs.DynamicResLayer(block=gs.DynamicBottleneck, **kw)
    gl.load_state_dict(sd)
    gl = gl.to(dev)
    gl.manipulate_arch({'width': w_act, 'depth': depth})
    gl.train()
    xg = Fg.as_act(x.to(dev)).requires_grad_(True)
    zg = gl(xg)
    zg.backward(Fg.as_act(dz.to(dev)))
    torch.cuda.synchronize()
    name = f'deep_stage[{depth} blocks,w{w_act}]'
    l2 = lambda a, b: float((a.double().cpu().flatten() - b.flatten()).norm() / (b.flatten().norm() + 1e-300))
    e_f, n_f = l2(zg.float(), z64), l2(z32, z64)
    out.append(dict(name=name + '.fwd_rel_l2', ok=e_f <= 4 * n_f + 2 * C.BF16_EPS, err=e_f, tol=4 * n_f + 2 * C.BF16_EPS,
                    oracle_fp32_vs_fp64=n_f))
    cos = lambda a, b: 1.0 - float(torch.dot(a.double().cpu().flatten(), b.flatten()) /
                                   (a.double().cpu().flatten().norm() * b.flatten().norm() + 1e-300))
    e_d, n_d = cos(xg.grad.float(), dx64), cos(dx32, dx64)
    out.append(dict(name=name + '.dx_1mcos', ok=e_d <= 4 * n_d + 5e-3, err=e_d, tol=4 * n_d + 5e-3, oracle_fp32_vs_fp64=n_d))
    gq = {n: p.grad.detach().cpu() for n, p in gl.named_parameters() if p.grad is not None}
    d_c, d_s = C._grad_cos(gq, g64), C._grad_cos(g32, g64)
    m_c, m_s = sum(d_c.values()) / len(d_c), sum(d_s.values()) / len(d_s)
    worst = max(d_c, key=d_c.get)
    out.append(dict(name=name + '.param_grads_mean_1mcos', ok=m_c <= 4 * m_s + 5e-3 and len(d_c) == len(d_s), err=m_c,
                    tol=4 * m_s + 5e-3, oracle_fp32_vs_fp64=m_s, worst_param=worst, worst_1mcos=d_c[worst],
                    n_params=len(d_c)))
    out += deep_stage_teacher_forced(gs, kw, w_act, depth, x, dz)
    return out


def deep_stage_teacher_forced(gs, kw, w_act, depth, x, dz):
    """TIGHT full-depth check without the chaos: the storage-emulating fp64 oracle runs the whole stage once; then EVERY
    one of its `depth` blocks is run alone on the CUDA path with the oracle's own (bf16-stored) input activation and
    output gradient of that block.  Each block's forward, input gradient and parameter gradients must match at the
    single-block tolerances of stage_checks (fwd rel-L2 2 * 2^-8, 1 - cos(dx) <= 2e-3, 1 - cos(param grad) <= 5e-3) --
    so depth indexing, per-block dilation / downsample wiring and every block's weights are pinned at full depth."""
    Fg = gs.functional
    dev = torch.device('cuda')
    ol = O.DynamicResLayer(block=O.DynamicBottleneck, **kw)
    C.randomize(ol, 9)
    sd = {k_: v.clone() for k_, v in ol.state_dict().items()}
    ol.manipulate_arch({'width': w_act, 'depth': depth})
    ol.train()
    C.emulate_bf16_storage(ol)
    ol = ol.double()
    xin, zout, gin, gout = {}, {}, {}, {}

    def mk(b):
        def hook(mod, inp, outp):
            xin[b], zout[b] = inp[0], outp
            outp.register_hook(lambda g_, b=b: gout.__setitem__(b, g_.detach().clone()))
            if inp[0].requires_grad:
                inp[0].register_hook(lambda g_, b=b: gin.__setitem__(b, g_.detach().clone()))
        return hook

    for b in range(depth):
        ol[b].register_forward_hook(mk(b))       # registered AFTER the storage-rounding hooks: sees the stored values
    xo = x.clone().double().requires_grad_(True)
    zo = ol(xo)
    zo.backward(dz.double())
    g_or = {n: p.grad.detach() for n, p in ol.named_parameters() if p.grad is not None}
    gl = gs.DynamicResLayer(block=gs.DynamicBottleneck, **kw)
    gl.load_state_dict(sd)
    gl = gl.to(dev)
    gl.manipulate_arch({'width': w_act, 'depth': depth})
    gl.train()
    worst = dict(fwd=(0.0, -1), dx=(0.0, -1), dp=(0.0, ''))
    for b in range(depth):
        xb = Fg.as_act(xin[b].detach().float().to(dev)).requires_grad_(True)
        zb = gl[b](xb)
        zb.backward(Fg.as_act(gout[b].float().to(dev)))
        torch.cuda.synchronize()
        ref = zout[b].detach()
        e = float((zb.float().double().cpu() - ref).norm() / (ref.norm() + 1e-300))
        worst['fwd'] = max(worst['fwd'], (e, b))
        d = C._grad_cos({'x': xb.grad.float().cpu()}, {'x': gin[b]})['x'] if b in gin else 0.0
        worst['dx'] = max(worst['dx'], (d, b))
        gq = {f'{b}.{n}': p.grad.detach().cpu() for n, p in gl[b].named_parameters() if p.grad is not None}
        dc = C._grad_cos(gq, {n: v for n, v in g_or.items() if n.startswith(f'{b}.')})
        for n, v in dc.items():
            worst['dp'] = max(worst['dp'], (v, n))
    name = f'deep_stage_teacher_forced[{depth} blocks]'
    return [dict(name=name + '.every_block_fwd_rel_l2', ok=worst['fwd'][0] <= 2 * C.BF16_EPS, err=worst['fwd'][0],
                 tol=2 * C.BF16_EPS, worst_block=worst['fwd'][1]),
            dict(name=name + '.every_block_dx_1mcos', ok=worst['dx'][0] <= 2e-3, err=worst['dx'][0], tol=2e-3,
                 worst_block=worst['dx'][1]),
            dict(name=name + '.every_block_param_grads_1mcos', ok=worst['dp'][0] <= 5e-3, err=worst['dp'][0], tol=5e-3,
                 worst_param=worst['dp'][1])]


# ------------------------------------------------------------------------------------------------
# BASELINE config 3 at full size: R101 anchor + ASPP head, fwd + loss vs the fp32 oracle
# ------------------------------------------------------------------------------------------------
R101 = {'backbone': {'stem': {'width': [32, 32, 64]}, 'body': {'width': [64, 128, 256, 512], 'depth': [3, 4, 23, 3]}}}


def config3_cfg():
    import bench as B
    cfg = B.supernet_cfg('os8')
    cfg['decode_head'] = dict(type='DynamicASPPHead', conv_cfg=dict(type='DynConv2d'), in_channels=2560, in_index=3,
                              channels=512, dilations=(1, 12, 24, 36), dropout_ratio=0.0, num_classes=19,
                              norm_cfg=dict(type='SyncBN', requires_grad=True), align_corners=False,
                              loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0))
    return cfg


def config3_full_size_checks(gs):
    """BASELINE configs[2]: R101 sub-net of the OS8 supernet + DeepLabV3 ASPP head, 2x3x512x1024, 19 classes: train-mode
    forward + loss against the fp32 CPU oracle (loss 2e-2 relative, acc_seg within 1 point, exact ignored-pixel count
    is covered by loss_checks), eval-mode label maps >= 97 % agreement.  (~1 minute of host time for the oracle.)"""
    import numpy as np
    out = []
    t0 = time.time()
    cfg = config3_cfg()
    om = O.build_segmentor(cfg)
    C.randomize(om, 13)
    gm = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    gm.load_state_dict(om.state_dict(), strict=True)
    gm = gm.cuda()
    om.manipulate_arch(R101); gm.manipulate_arch(R101)
    g = torch.Generator().manual_seed(33)
    img = bf16r(torch.randn(2, 3, 512, 1024, generator=g))
    lab = C._labels(g, 2, 19, 512, 1024)
    om.train(); gm.train()
    with torch.no_grad():
        lo_d = om.forward_train(img, None, lab)
        lo = om.parse_losses(lo_d)
        res = gm.train_step(dict(img=img.cuda(), img_metas=[{}, {}], gt_semantic_seg=lab.cuda()), None)
    out.append(check_f32(res['loss'].reshape(1).cpu(), lo.reshape(1), 'config3.R101_ASPP.train_loss_vs_fp32_oracle', 2e-2))
    acc_o, acc_g = float(lo_d['decode.acc_seg']), float(res['log_vars']['decode.acc_seg'])
    out.append(dict(name='config3.R101_ASPP.acc_seg', ok=abs(acc_o - acc_g) < 1.0, err=abs(acc_o - acc_g), tol=1.0))
    om.eval(); gm.eval()
    metas = [[dict(ori_shape=(512, 1024, 3), flip=False)] * 2]
    with torch.no_grad():
        po = om.simple_test(img)
        pg = gm(return_loss=False, img=[img.cuda()], img_metas=metas)
    agree = float((torch.from_numpy(np.stack(pg)) == po).float().mean())
    out.append(dict(name='config3.R101_ASPP.labelmap_agreement', ok=agree >= 0.97, err=1 - agree, tol=0.03,
                    seconds=round(time.time() - t0, 1)))
    return out


# ------------------------------------------------------------------------------------------------
# fused upsample + CE at the benchmark size, forward AND backward values
# ------------------------------------------------------------------------------------------------
def loss_full_size_checks(gs):
    """2 x 19 x 64 x 128 -> 512 x 1024 (the benchmark's loss call) and 2 x 150 x 64 x 64 -> 512 x 512 (config 4): loss,
    exact ignored-pixel count and dlogits against F.interpolate -> cross_entropy autograd on the host."""
    Fg = gs.functional
    out = []
    for (N, K, h, w, H, W) in ((2, 19, 64, 128, 512, 1024), (2, 150, 64, 64, 512, 512)):
        g = torch.Generator().manual_seed(K + H)
        tag = f'loss_full[N{N},K{K},{h}x{w}->{H}x{W}]'
        logits = torch.randn(N, K, h, w, generator=g) * 2
        lab = C._labels(g, N, K, H, W)
        lo = logits.clone().requires_grad_(True)
        up = F.interpolate(lo, size=(H, W), mode='bilinear', align_corners=False)
        loss_o = O.cross_entropy(up, lab.squeeze(1), 255)
        loss_o.backward()
        lg = logits.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        loss_g, acc_g, counts = Fg.upsample_ce(lg, lab.cuda(), 255, 1.0)
        loss_g.backward()
        torch.cuda.synchronize()
        out.append(check_f32(loss_g.reshape(1).cpu(), loss_o.detach().reshape(1), tag + '.loss', 1e-4))
        n_ign = int((lab == 255).sum())
        out.append(dict(name=tag + '.ignored_count_exact', ok=int(counts[0]) == n_ign, err=abs(int(counts[0]) - n_ign), tol=0))
        out.append(check_f32(lg.grad.cpu(), lo.grad, tag + '.dlogits', 2e-3))
        # deterministic: a second backward gives bit-identical gradients (single writer per cell, fixed order)
        lg2 = logits.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        Fg.upsample_ce(lg2, lab.cuda(), 255, 1.0)[0].backward()
        torch.cuda.synchronize()
        out.append(dict(name=tag + '.dlogits_bitwise_reproducible', ok=bool(torch.equal(lg.grad, lg2.grad)), err=0.0, tol=0))
    return out


def bn_bwd_one_pass_checks(gs):
    """DynBN backward: the ONE-launch channel-partitioned cluster kernel (gs_bn_bwd) and the reduce + apply pair against an
    fp64 restatement of F.batch_norm's backward on the channel prefix -- every mask mode (none / recomputed from y /
    from the stored output with a residual gradient), ragged channel counts (C8 not a power of two), a pixel count that
    is no multiple of anything, and the benchmark's stage-1 / stage-3 sizes."""
    Fg = gs.functional
    dev = torch.device('cuda')
    out = []
    cases = [  # N, C, Cmax, H, W, relu, residual
        (3, 40, 64, 9, 14, True, False), (3, 40, 64, 9, 14, False, False), (2, 72, 80, 17, 23, True, True),
        (1, 8, 8, 5, 7, True, False), (2, 320, 320, 64, 128, True, False), (2, 64, 80, 128, 256, True, False),
        (2, 1280, 1280, 32, 32, True, True), (2, 160, 160, 64, 128, False, False)]
    saved = Fg.FUSED_BN_BWD
    try:
        for (N, C, Cmax, H, W, relu, has_res) in cases:
            g = torch.Generator().manual_seed(C * 7 + H)
            bn = gs.DynamicBatchNorm2d(Cmax).to(dev).train()
            with torch.no_grad():
                bn.weight.copy_(torch.rand(Cmax, generator=g) + 0.5)
                bn.bias.copy_(torch.randn(Cmax, generator=g) * 0.1)
            y = Fg.as_act((torch.randn(N, C, H, W, generator=g) * 2 + 0.5).to(dev))
            res = Fg.as_act(torch.randn(N, C, H, W, generator=g).to(dev)) if has_res else None
            dz = Fg.as_act(torch.randn(N, C, H, W, generator=g).to(dev))
            z, aff, count = Fg.bn_train_apply(bn, y, Fg.bn_stats(y), C, res, relu)
            torch.cuda.synchronize()
            y64, dz64, z64 = y.double().cpu(), dz.double().cpu(), z.double().cpu()
            mean, invstd = aff[0].double().cpu().view(1, C, 1, 1), aff[1].double().cpu().view(1, C, 1, 1)
            gam = bn.weight[:C].double().cpu().view(1, C, 1, 1)
            gg = dz64 * (z64 > 0) if relu else dz64
            xh = (y64 - mean) * invstd
            m = float(N * H * W)
            sg, sx = gg.sum((0, 2, 3), keepdim=True), (gg * xh).sum((0, 2, 3), keepdim=True)
            dy_ref = gam * invstd * (gg - sg / m - xh * sx / m)
            tag = f'bn_bwd[N{N},C{C},{H}x{W},relu={int(relu)},res={int(has_res)}]'
            got = {}
            for mode in ('0', '1'):
                Fg.FUSED_BN_BWD = mode
                bn.weight.grad = None
                bn.bias.grad = None
                dy, dres = Fg.bn_backward(bn, dz, y, aff, count, z if has_res else None, relu, has_res)
                torch.cuda.synchronize()
                name = tag + ('.one_pass' if mode == '1' else '.two_kernels')
                out.append(check_bf16(dy.float().cpu(), dy_ref.float(), name + '.dy', 4.0))
                out.append(check_f32(bn.weight.grad[:C].cpu(), sx.view(C).float(), name + '.dgamma', 2e-3))
                out.append(check_f32(bn.bias.grad[:C].cpu(), sg.view(C).float(), name + '.dbeta', 2e-3))
                if Cmax > C:
                    out.append(dict(name=name + '.inactive_grads_zero', tol=0,
                                    ok=bool((bn.weight.grad[C:] == 0).all() and (bn.bias.grad[C:] == 0).all()), err=0.0))
                if has_res:
                    out.append(dict(name=name + '.dres_exact', ok=bool((dres.float().cpu() == gg.float()).all()), err=0.0, tol=0))
                got[mode] = dy.float().cpu()
            # the two paths differ only in the summation order of the per-channel sums
            out.append(check_bf16(got['1'], got['0'], tag + '.one_pass_vs_two_kernels', 2.0))
    finally:
        Fg.FUSED_BN_BWD = saved
    return out


def fcn_head_skip_gradient_checks(gs):
    """FCN head with concat_input (fcn_head.py:68-81): x feeds convs[0] AND the concat.  The product path adds the concat's
    share of dL/dx inside convs[0]'s dgrad epilogue (functional.GradCarrier); it must equal autograd's own accumulation of
    the two gradients (same kernels, GS_SKIP_CARRIER=0 behaviour) and the fp32 oracle head."""
    Fg = gs.functional
    dev = torch.device('cuda')
    out = []
    for num_convs in (1, 2):
        torch.manual_seed(11 + num_convs)
        kw = dict(in_channels=48, channels=32, num_classes=19, num_convs=num_convs, concat_input=True, dropout_ratio=0.0,
                  conv_cfg=dict(type='DynConv2d'), norm_cfg=dict(type='DynBN', requires_grad=True), in_index=0)
        oh = O.DynamicFCNHead(**kw)
        C.randomize(oh, seed=3)
        head = gs.DynamicFCNHead(**kw)
        head.load_state_dict(oh.state_dict(), strict=True)
        head = head.to(dev).train()
        x = bf16r(torch.randn(2, 48, 12, 20))
        dz = bf16r(torch.randn(2, 19, 12, 20))
        grads = {}
        saved = Fg.SKIP_GRAD_CARRIER
        try:
            for mode in (True, False):
                Fg.SKIP_GRAD_CARRIER = mode
                head.zero_grad(set_to_none=True)
                xg = Fg.as_act(x.to(dev)).requires_grad_(True)
                y = head.forward([xg])
                y.backward(dz.to(dev).to(y.dtype).contiguous(memory_format=torch.channels_last))
                Fg.wgrad_join()
                torch.cuda.synchronize()
                grads[mode] = (xg.grad.float().cpu(), {n: p.grad.float().cpu().clone() for n, p in head.named_parameters()
                                                       if p.grad is not None})
        finally:
            Fg.SKIP_GRAD_CARRIER = saved
        tag = f'fcn_head_skip[num_convs={num_convs}]'
        out.append(check_bf16(grads[True][0], grads[False][0], tag + '.dx_carrier_vs_autograd_sum', 3.0))
        # same kernels on the same inputs in both modes; the only freedom is the ORDER in which the stream-K units of the weight
        # gradient add their fp32 partial sums (an item can be split over three units -> (a + b) + c vs (a + c) + b), so the
        # parameter gradients agree to an fp32 rounding, not always bit for bit (seen once in ~6 full-suite runs)
        worst = max(float((grads[True][1][n] - g).abs().max()) / (float(g.abs().max()) + 1e-30)
                    for n, g in grads[False][1].items())
        out.append(dict(name=tag + '.param_grads_equal_up_to_summation_order', ok=worst <= 1e-6, err=worst, tol=1e-6))
        # fp32 oracle head on the same parameters
        oh.train()
        xo = x.clone().requires_grad_(True)
        oh.forward([xo]).backward(dz)
        # (fp32 oracle without bf16 storage between the 2-3 conv + train-mode-BN layers: judged in relative L2 -- 0.08 is
        #  1 - cos of 3e-3, the stage tests' bound; a lost skip gradient would show as > 0.5)
        out.append(rel_l2(grads[True][0], xo.grad, 8e-2, tag + '.dx_vs_oracle_rel_l2'))
    return out


def conv_bn_fused_launch_checks(gs):
    """gs_conv2d_fwd_bn (conv + DynBN apply in ONE launch: statistic flush -> grid barrier -> every CTA normalises its own
    tiles) against the two-kernel path (gs_conv2d_fwd + gs_bn_apply_train) on the same inputs: conv output y identical,
    z within 1 bf16 ulp (the fp64 statistic atomics arrive in a different order), aff / running statistics 1e-6 --
    single-CTA and CTA-pair kernels, wide n-tiles (Cout 320), several tiles per CTA, a phantom tile (odd tile count),
    ragged maps, channel-prefix slices of a wider BN, residual + ReLU, no ReLU, a conv bias."""
    Fg = gs.functional
    dev = torch.device('cuda')
    out = []
    cases = [  # name, N, H, W, Ci, Co, Co_max, k, dil, residual, relu, bias
        ('1x1_short_k', 2, 32, 32, 64, 96, 160, 1, 1, False, True, False),
        ('1x1_multi_tile', 2, 64, 128, 96, 384, 384, 1, 1, False, True, False),
        ('1x1_res_1280', 2, 64, 128, 320, 1280, 1280, 1, 1, True, True, False),
        ('3x3_pair_wide320', 2, 64, 128, 320, 320, 320, 3, 2, False, True, False),
        ('3x3_pair_oddtiles', 3, 8, 16, 320, 320, 320, 3, 1, True, False, False),
        ('3x3_ragged_bias', 1, 17, 23, 64, 80, 80, 3, 1, False, True, True),
        ('1x1_pair_1280_320', 2, 64, 128, 1280, 320, 320, 1, 1, False, True, False),
    ]
    saved = Fg.CONV_BN_FUSE
    try:
        for (name, N, H, W, Ci, Co, Co_max, k, dil, has_res, relu, bias) in cases:
            g = torch.Generator().manual_seed(Ci + 3 * Co + H)
            conv = gs.DynamicConv2d(Ci, Co_max, k, padding=dil * (k // 2), dilation=dil, bias=bias).to(dev)
            with torch.no_grad():
                conv.weight.copy_(bf16r(torch.randn(conv.weight.shape, generator=g) * (2.0 / (Ci * k * k)) ** 0.5))
                if bias:
                    conv.bias.copy_(torch.randn(Co_max, generator=g) * 0.1)
            x = Fg.as_act(torch.randn(N, Ci, H, W, generator=g).to(dev))
            res = Fg.as_act(torch.randn(N, Co, H, W, generator=g).to(dev)) if has_res else None
            Fg.conv_shadows(conv)                      # (the bf16 weight shadow is cast once, outside the launch count)
            got = {}
            for fuse in (False, True):
                Fg.CONV_BN_FUSE = fuse
                bn = gs.DynamicBatchNorm2d(Co_max).to(dev).train()
                with torch.no_grad():
                    gg = torch.Generator().manual_seed(Co)
                    bn.weight.copy_(torch.rand(Co_max, generator=gg) + 0.5)
                    bn.bias.copy_(torch.randn(Co_max, generator=gg) * 0.1)
                n0 = gs._lib.launch_count()
                z, rec = Fg.cba_forward(x, conv, bn, relu=relu, residual=res, Co=Co)
                torch.cuda.synchronize()
                got[fuse] = dict(z=z.float().cpu(), y=rec.y.float().cpu(), aff=rec.aff.cpu(), rm=bn.running_mean.cpu().clone(),
                                 rv=bn.running_var.cpu().clone(), launches=gs._lib.launch_count() - n0)
            a, b = got[True], got[False]
            out.append(dict(name=f'conv_bn_fused[{name}].one_launch', ok=a['launches'] == 1 and b['launches'] == 2,
                            err=float(a['launches']), tol=1))
            out.append(dict(name=f'conv_bn_fused[{name}].y_identical', ok=bool(torch.equal(a['y'], b['y'])), err=0.0, tol=0))
            out.append(check_bf16(a['z'], b['z'], f'conv_bn_fused[{name}].z', 1.0))
            out.append(check_f32(a['aff'], b['aff'], f'conv_bn_fused[{name}].aff', 1e-6))
            out.append(check_f32(a['rm'][:Co], b['rm'][:Co], f'conv_bn_fused[{name}].running_mean', 1e-6))
            out.append(check_f32(a['rv'][:Co], b['rv'][:Co], f'conv_bn_fused[{name}].running_var', 1e-6))
            out.append(dict(name=f'conv_bn_fused[{name}].stats_outside_prefix_untouched',
                            ok=bool(torch.equal(a['rm'][Co:], b['rm'][Co:]) and torch.equal(a['rv'][Co:], b['rv'][Co:])),
                            err=0.0, tol=0))
    finally:
        Fg.CONV_BN_FUSE = saved
    return out
