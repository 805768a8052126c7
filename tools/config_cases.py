#!/usr/bin/env python
"""BASELINE.json configs 1, 3, 4, 5 at their full sizes on one GPU (synthetic data, random-init weights).  These are the
parity-test configurations, not bench lines: the numbers are reported for context next to bench.py's configs[1] line.

  config 1: MIN sub-net + FCN head, forward + loss, 2x3x512x512: CPU oracle (host cores) next to the CUDA path
  config 3: dynamic ResNet-101 anchor (and MAX) + DeepLabV3 ASPP head, bf16 fwd/bwd/SGD, 2x3x512x1024, 19 classes
  config 4: extract_subnet(R50) -> fixed-arch finetune, 2x3x512x512, 150 classes; extracted == manipulated, bit-exact
  config 5: test_supernet sweep: 50 sub-nets ~ random.Random(0), whole-image inference 1x3x1024x2048, int64 label maps

    python tools/config_cases.py [3 4 5]  ->  one JSON object per config on stdout + gpurun_out/config_cases.json
"""
import copy
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (supernet / sampler definitions of the headline workload)
import gaia_seg_b200 as gs  # noqa: E402

R101 = {'backbone': {'stem': {'width': [32, 32, 64]}, 'body': {'width': [64, 128, 256, 512], 'depth': [3, 4, 23, 3]}}}
R50 = {'backbone': {'stem': {'width': [32, 32, 64]}, 'body': {'width': [64, 128, 256, 512], 'depth': [3, 4, 6, 3]}}}
MAX = {'backbone': {'stem': {'width': [32, 32, 64]}, 'body': {'width': [80, 160, 320, 640], 'depth': [4, 6, 29, 4]}}}


def timed(fn, reps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def batch(N, H, W, K, dev, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(N, 3, H, W, generator=g)
    lab = torch.randint(0, K, (N, 1, H, W), generator=g)
    lab[torch.rand(N, 1, H, W, generator=g) < 0.1] = 255
    return dict(img=img.to(dev), img_metas=[dict(ori_shape=(H, W, 3), flip=False)] * N, gt_semantic_seg=lab.to(dev))


def train_fn(model, opt, data, key='fixed', graph=True):
    """One training iteration the way IterBasedRunner runs it: through GraphedTrainStep (a fixed architecture is
    captured on its 3rd occurrence and replayed afterwards); graph=False keeps every iteration eager."""
    stepper = gs.GraphedTrainStep(model, opt, graph_after=2 if graph else 10 ** 9, max_graphs=2, pool_gb=16)

    def step():
        return stepper(key, data)
    return step


def config1(dev):
    """BASELINE configs[0] (the reference's CPU-runnable case): MIN sub-net + FCN head, forward + loss, 2x3x512x512,
    19 classes -- the CPU oracle on the host cores next to the CUDA path on the same weights and inputs."""
    import time
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from oracle import ref_model as O
    import gs_checks as C
    cfg = bench.supernet_cfg('os32')
    cfg['backbone']['norm_cfg'] = dict(type='DynBN', requires_grad=True)
    cfg['decode_head'].update(dropout_ratio=0.0, norm_cfg=dict(type='BN', requires_grad=True))
    om = O.build_segmentor(cfg)
    C.randomize(om, seed=0)
    gm = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    gm.load_state_dict(om.state_dict())
    gm = gm.to(dev)
    arch = {'backbone': {'stem': {'width': 32}, 'body': {'width': [48, 96, 192, 384], 'depth': [2, 2, 5, 2]}}}
    om.manipulate_arch(arch)
    gm.manipulate_arch(arch)
    torch.manual_seed(0)
    img = torch.randn(2, 3, 512, 512)
    lab = torch.randint(0, 19, (2, 1, 512, 512))
    lab[torch.rand(2, 1, 512, 512) < 0.1] = 255
    om.train()
    gm.train()
    torch.set_num_threads(os.cpu_count())
    best, loss_o = 1e9, None
    with torch.no_grad():
        for i in range(4):
            t0 = time.perf_counter()
            lo = om.forward_train(img, None, lab)
            dt = time.perf_counter() - t0
            if i > 0:
                best = min(best, dt)
            loss_o = float(lo['decode.loss_seg'])
            acc_o = float(lo['decode.acc_seg'])
    data = dict(img=img.to(dev), img_metas=[{}, {}], gt_semantic_seg=lab.to(dev))

    def fwd():
        with torch.no_grad():
            return gm.forward_train(data['img'], data['img_metas'], data['gt_semantic_seg'])
    ms = timed(fwd, 10, 2)
    lg = fwd()
    loss_g, acc_g = float(lg['decode.loss_seg']), float(lg['decode.acc_seg'])
    return {'config': 1, 'workload': 'MIN sub-net (OS32 supernet) + FCN head, forward + loss, 2x3x512x512, 19 classes, '
                                     'train-mode BN', 'cpu_oracle': dict(seconds=round(best, 3), imgs_per_s=round(2 / best, 2),
                                                                         cores=os.cpu_count(), loss=loss_o, acc_seg=acc_o),
            'cuda': dict(ms=round(ms, 3), imgs_per_s=round(2e3 / ms, 1), loss=loss_g, acc_seg=acc_g),
            'loss_rel_err': abs(loss_g - loss_o) / abs(loss_o), 'ok': abs(loss_g - loss_o) <= 2e-2 * abs(loss_o)}


def config3(dev):
    cfg = bench.supernet_cfg('os8')
    cfg['decode_head'] = dict(type='DynamicASPPHead', conv_cfg=dict(type='DynConv2d'), in_channels=2560, in_index=3,
                              channels=512, dilations=(1, 12, 24, 36), dropout_ratio=0.1, num_classes=19,
                              norm_cfg=dict(type='SyncBN', requires_grad=True), align_corners=False,
                              loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0))
    gs.set_random_seed(0)
    model = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole')).to(dev).train()
    opt = gs.GsSGD(model, lr=0.01, momentum=0.9, weight_decay=5e-4)
    data = batch(2, 512, 1024, 19, dev, 3)
    res = {'config': 3, 'workload': 'dynamic ResNet-101 (OS8 V1c supernet) + DeepLabV3 ASPP head, bf16 fwd+bwd+SGD, '
                                    '2x3x512x1024, 19 classes'}
    # ONE GraphedTrainStep (= one stream) per model, keyed by the sub-net, as IterBasedRunner drives it.  (Round 1 built a
    # second stepper -- a second stream -- for the second sub-net of the SAME parameters; autograd's gradient accumulators
    # remember the stream they were created on, so the captured backward then depended on uncaptured work of the first
    # stepper's stream: cudaErrorStreamCaptureIsolation.  The product path never does that.)
    stepper = gs.GraphedTrainStep(model, opt, graph_after=2 if os.environ.get('CONFIG3_GRAPH', '1') == '1' else 10 ** 9,
                                  max_graphs=2, pool_gb=16)
    for name, arch in (('R101', R101), ('MAX', MAX)):
        model.manipulate_arch(arch)
        step = lambda name=name: stepper(name, data)
        ms = timed(step, 5, 3)
        out = step()
        res[name] = dict(ms_per_step=round(ms, 2), imgs_per_s=round(2e3 / ms, 1), loss=float(out['loss']),
                         acc_seg=float(out['log_vars']['decode.acc_seg']), finite=bool(torch.isfinite(out['loss'])))
    return res


def config4(dev):
    cfg = bench.supernet_cfg('os8')
    cfg['decode_head'].update(num_classes=150)

    def swap(d):   # tools/extract_subnet.py: every norm becomes a plain DynBN before deploy()
        if isinstance(d, dict):
            if d.get('type') in ('DynSyncBN', 'SyncBN'):
                d['type'] = 'DynBN'
                d.pop('group_size', None)
            for v in d.values():
                swap(v)
    swap(cfg)
    gs.set_random_seed(0)
    model = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole')).to(dev).eval()
    for m in model.modules():   # non-trivial running statistics
        if hasattr(m, 'running_mean') and m.running_mean is not None:
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    model.manipulate_arch(R50)
    img = torch.randn(2, 3, 512, 512, device=dev)
    with torch.no_grad():
        ref = model.encode_decode_lowres(img, None).float().clone()
    model.deploy()
    sub = copy.deepcopy(model)
    with torch.no_grad():
        got = sub.encode_decode_lowres(img, None).float().clone()     # first forward slices every tensor physically
        again = sub.encode_decode_lowres(img, None).float().clone()
    n_sup = sum(p.numel() for p in model.parameters())
    n_sub = sum(p.numel() for p in sub.parameters())
    sub.train()
    opt = gs.GsSGD(sub, lr=0.01, momentum=0.9, weight_decay=5e-4)
    data = batch(2, 512, 512, 150, dev, 4)
    step = train_fn(sub, opt, data)
    l0 = float(step()['loss'])
    ms = timed(step, 10, 3)
    l1 = float(step()['loss'])
    return {'config': 4, 'workload': 'extract_subnet(R50) -> finetune, FCN head, 2x3x512x512, 150 classes',
            'extracted_logits_bit_identical': bool(torch.equal(ref, got) and torch.equal(got, again)),
            'params_M': {'supernet': round(n_sup / 1e6, 1), 'extracted': round(n_sub / 1e6, 1)},
            'finetune': dict(ms_per_step=round(ms, 2), imgs_per_s=round(2e3 / ms, 1), loss_first=l0, loss_after_13_steps=l1)}


def config5(dev, n_subnets=50):
    cfg = bench.supernet_cfg('os8')
    gs.set_random_seed(0)
    model = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole')).to(dev).eval()
    rng = random.Random(0)
    stems = [[16, 16, 32], [24, 24, 48], [32, 32, 64]]
    widths = [[48, 64, 80], [96, 128, 160], [192, 256, 320], [384, 512, 640]]
    depths = [[2, 3, 4], [2, 4, 6], list(range(5, 30, 2)), [2, 3, 4]]
    H, W = 1024, 2048
    img = torch.randn(1, 3, H, W, device=dev)
    metas = [[dict(ori_shape=(H, W, 3), flip=False)]]
    rows = []
    for i in range(n_subnets):
        arch = {'backbone': {'stem': {'width': rng.choice(stems)},
                             'body': {'width': [rng.choice(w) for w in widths], 'depth': [rng.choice(d) for d in depths]}}}
        model.manipulate_arch(arch)

        def infer():
            with torch.no_grad():
                return model(return_loss=False, rescale=True, img=[img], img_metas=metas)
        ms = timed(infer, 2, 1)          # includes the device -> host copy of the 1024x2048 int64 label map
        seg = infer()[0]
        assert seg.shape == (H, W) and seg.dtype.name == 'int64' and 0 <= seg.min() and seg.max() < 19
        rows.append(dict(arch=arch['backbone'], ms=round(ms, 2), imgs_per_s=round(1e3 / ms, 2)))
    tot = sum(r['ms'] for r in rows)
    return {'config': 5, 'workload': f'test_supernet sweep, {n_subnets} sub-nets ~ random.Random(0), eval-mode running '
                                     'stats, whole-image 1x3x1024x2048 -> int64 label map on the host',
            'aggregate_imgs_per_s': round(n_subnets * 1e3 / tot, 2), 'min_imgs_per_s': min(r['imgs_per_s'] for r in rows),
            'max_imgs_per_s': max(r['imgs_per_s'] for r in rows), 'subnets': rows}


def main():
    which = [int(a) for a in sys.argv[1:]] or [1, 3, 4, 5]
    gs._lib.require_device()
    dev = torch.device('cuda', 0)
    out = []
    for c in which:
        r = {1: config1, 3: config3, 4: config4, 5: config5}[c](dev)
        print(json.dumps({k: v for k, v in r.items() if k != 'subnets'}), flush=True)
        out.append(r)
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'config_cases.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
