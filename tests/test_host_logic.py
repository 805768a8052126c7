"""Host-side mirror of the reference interface: registry, arch manipulation, samplers, config, deploy slicing,
state_dict compatibility.  Integer / indexing logic is bit-exact by construction.  No GPU."""
import copy
import os
import textwrap

import pytest
import torch

import gs_checks as C
from oracle import ref_model as O

MAX = {'backbone': {'stem': {'width': 64}, 'body': {'width': [80, 160, 320, 640], 'depth': [4, 6, 29, 4]}}}
MIN = {'backbone': {'stem': {'width': 32}, 'body': {'width': [48, 96, 192, 384], 'depth': [2, 2, 5, 2]}}}


def test_fold_unfold_roundtrip(gs):
    flat = {'name': 'R50', 'arch.backbone.stem.width': 64, 'arch.backbone.body.width': [64, 128, 256, 512],
            'arch.backbone.body.depth': [3, 4, 6, 3]}
    nested = gs.fold_dict(flat)
    assert nested['arch']['backbone']['body']['depth'] == [3, 4, 6, 3] and nested['name'] == 'R50'
    assert gs.unfold_dict(nested) == flat
    with pytest.raises(ValueError):
        gs.fold_dict({'a': 1, 'a.b': 2})


def test_registry_and_build_from_cfg(gs):
    assert gs.SEGMENTORS.get('DynamicEncoderDecoder') is gs.DynamicEncoderDecoder
    assert gs.BACKBONES.get('DynamicResNet') is gs.DynamicResNet
    assert gs.HEADS.get('DynamicFCNHead') is gs.DynamicFCNHead
    assert gs.CONV_LAYERS.get('DynConv2d') is gs.DynamicConv2d
    assert gs.NORM_LAYERS.get('DynSyncBN') is gs.DynamicSyncBatchNorm and gs.NORM_LAYERS.get('DynBN') is gs.DynamicBatchNorm2d
    with pytest.raises(KeyError):
        gs.build_from_cfg(dict(type='Nope'), gs.BACKBONES)
    with pytest.raises(KeyError):
        gs.build_from_cfg(dict(), gs.BACKBONES)
    name, layer = gs.build_norm_layer(dict(type='DynSyncBN', requires_grad=False, group_size=1), 32, postfix=1)
    assert name == 'bn1' and isinstance(layer, torch.nn.modules.batchnorm._BatchNorm) and not layer.weight.requires_grad


def test_manipulate_arch_fans_out_like_the_reference(gs):
    cfg = C.small_cfg(aux=True)
    m = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    o = O.build_segmentor(cfg)
    for arch in C.SMALL_ARCHS.values():
        m.manipulate_arch(arch)
        o.manipulate_arch(arch)
        got = {n: mod.width_state for n, mod in m.named_modules() if isinstance(mod, gs.DynamicConv2d)}
        ref = {n: mod.width_state for n, mod in o.named_modules() if isinstance(mod, O.DynamicConv2d)}
        assert got == ref
        for i in range(4):
            assert getattr(m.backbone, f'layer{i + 1}').depth_state == arch['backbone']['body']['depth'][i]
    w = arch['backbone']['body']['width']
    blk = m.backbone.layer3[0]
    assert (blk.conv1.width_state, blk.conv2.width_state, blk.conv3.width_state) == (w[2], w[2], 4 * w[2])
    assert blk.downsample[0].width_state == 4 * w[2]
    with pytest.raises(KeyError):
        m.manipulate_arch({'nonexistent': {}})
    with pytest.raises(AssertionError):
        m.backbone.layer1.manipulate_depth(0)
    with pytest.raises(AssertionError):
        m.backbone.conv1.manipulate_width(10 ** 6)


@pytest.mark.parametrize('kw', [dict(aux=True, deep_stem=True, os8=True), dict(psp=True, aux=True, os8=True),
                                dict(aspp=True, os8=True)], ids=['fcn_aux', 'psp_aux', 'aspp'])
def test_state_dict_is_reference_format(gs, kw):
    cfg = C.small_cfg(**kw)
    m = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    o = O.build_segmentor(cfg)
    sd_m, sd_o = m.state_dict(), o.state_dict()
    assert list(sd_m.keys()) == list(sd_o.keys())
    for k in sd_m:
        assert tuple(sd_m[k].shape) == tuple(sd_o[k].shape), k          # logical OIHW max-width tensors
    m.load_state_dict(sd_o, strict=True)
    w = m.backbone.layer1[0].conv2.weight
    assert torch.equal(w.detach(), sd_o['backbone.layer1.0.conv2.weight'])
    assert w.permute(0, 2, 3, 1).is_contiguous()                           # KRSC memory kept after loading


def test_zero_init_residual_and_head_init(gs):
    cfg = C.small_cfg()
    m = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    for blk in m.backbone.layer2:
        assert float(blk.norm3.weight.abs().sum()) == 0.0 and float(blk.norm1.weight.min()) == 1.0
    assert float(m.decode_head.conv_seg.bias.abs().sum()) == 0.0
    assert 0.005 < float(m.decode_head.conv_seg.weight.std()) < 0.02


def test_sampler_cycle_matches_shipped_config(gs):
    rng = dict(type='composite', model_samplers=[
        dict(type='range', key='arch.backbone.stem.width', start=32, end=64, step=16),
        dict(type='range', key='arch.backbone.body.width', start=[48, 96, 192, 384], end=[80, 160, 320, 640],
             step=[16, 32, 64, 128], ascending=True),
        dict(type='range', key='arch.backbone.body.depth', start=[2, 2, 5, 2], end=[4, 6, 29, 4], step=[1, 2, 2, 1])])
    anchors = [dict(name=n, **{'arch.backbone.stem.width': 64}) for n in ('MAX', 'MIN', 'R101', 'R77', 'R50')]
    s = gs.build_model_sampler(dict(type='concat', seed=0, model_samplers=[
        dict(type='anchor', anchors=anchors), dict(type='repeat', times=3, model_sampler=rng)]))
    assert s.period() == 8
    names = [s.sample().get('name') for _ in range(16)]
    assert names[:5] == ['MAX', 'MIN', 'R101', 'R77', 'R50'] and names[5:8] == [None] * 3 and names[8:13] == names[:5]
    assert [s.anchor_name(i) for i in range(5)] == ['MAX', 'MIN', 'R101', 'R77', 'R50']
    for _ in range(200):
        meta = gs.fold_dict(gs.build_model_sampler(dict(rng, seed=_)).sample())['arch']['backbone']
        w, d = meta['body']['width'], meta['body']['depth']
        assert meta['stem']['width'] in (32, 48, 64)
        assert all(a in g for a, g in zip(w, ([48, 64, 80], [96, 128, 160], [192, 256, 320], [384, 512, 640])))
        assert w == sorted(w)
        assert d[0] in (2, 3, 4) and d[1] in (2, 4, 6) and d[2] in range(5, 30, 2) and d[3] in (2, 3, 4)
    n = sum(1 for _ in gs.build_model_sampler(rng).traverse())
    assert n == 3 * 3 ** 4 * (3 * 3 * 13 * 3)          # 85 293 sub-nets


def test_sampler_is_deterministic_under_a_shared_seed(gs):
    MAXa, MINa = dict(name='MAX', **{'arch.x': 1}), dict(name='MIN', **{'arch.x': 0})
    rnd = dict(type='range', key='arch.x', start=0, end=100, step=1)
    a = gs.build_model_sampler(gs.sandwich_sampler_cfg(MAXa, MINa, rnd, num_random=2, seed=3))
    b = gs.build_model_sampler(gs.sandwich_sampler_cfg(MAXa, MINa, rnd, num_random=2, seed=3))
    sa, sb = [a.sample() for _ in range(12)], [b.sample() for _ in range(12)]
    assert sa == sb and sa[0]['name'] == 'MAX' and sa[1]['name'] == 'MIN' and sa[4]['name'] == 'MAX'
    c = gs.build_model_sampler(gs.sandwich_sampler_cfg(MAXa, MINa, rnd, num_random=2, seed=4))
    assert [c.sample() for _ in range(12)] != sa


def test_deploy_slices_parameters_physically(gs):
    cfg = C.small_cfg()
    m = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    arch = C.SMALL_ARCHS['min']
    m.manipulate_arch(arch)
    m.deploy()
    bb = m.backbone
    # slicing itself is host-side tensor indexing; emulate what the first deploy forward does per module
    bb.conv1.deploy_slice(3)
    assert tuple(bb.conv1.weight.shape) == (16, 3, 7, 7)
    blk = bb.layer1[0]
    blk.conv1.deploy_slice(16); blk.norm1.deploy_slice(16)
    assert tuple(blk.conv1.weight.shape) == (16, 16, 1, 1) and blk.norm1.num_features == 16
    assert blk.norm1.running_mean.shape == (16,)
    assert all(getattr(mod, '_deploying', False) for mod in m.modules() if isinstance(mod, gs.DynamicMixin))


def test_config_loader_base_merge_and_overrides(gs, tmp_path):
    (tmp_path / 'base.py').write_text(textwrap.dedent('''
        model = dict(type='DynamicEncoderDecoder', backbone=dict(type='DynamicResNet', stem_width=64))
        optimizer = dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0005)
        runner = dict(type='IterBasedRunner', max_iters=80000)
    '''))
    (tmp_path / 'child.py').write_text(textwrap.dedent('''
        _base_ = ['base.py']
        model = dict(backbone=dict(stem_width=32))
        lr_config = dict(policy='poly', power=0.9, min_lr=1e-4, by_epoch=False)
    '''))
    cfg = gs.Config.fromfile(str(tmp_path / 'child.py'))
    assert cfg.model.backbone.stem_width == 32 and cfg.model.backbone.type == 'DynamicResNet'
    assert cfg.optimizer.lr == 0.01 and cfg.runner.max_iters == 80000 and cfg.get('nothing') is None
    cfg.merge_from_dict({'optimizer.lr': 0.02, 'data.samples_per_gpu': 2})
    assert cfg.optimizer.lr == 0.02 and cfg.data.samples_per_gpu == 2


def test_poly_lr_schedule(gs):
    from gaia_seg_b200.runner import PolyLrUpdaterHook

    class R:
        pass
    r = R()
    r.optimizer = type('O', (), {'param_groups': [dict(lr=0.01)]})()
    r.max_iters, r.iter = 80000, 0
    h = PolyLrUpdaterHook(power=0.9, min_lr=1e-4)
    h.before_run(r)
    h.before_train_iter(r)
    assert abs(r.optimizer.param_groups[0]['lr'] - 0.01) < 1e-12
    r.iter = 40000
    h.before_train_iter(r)
    assert abs(r.optimizer.param_groups[0]['lr'] - ((0.01 - 1e-4) * 0.5 ** 0.9 + 1e-4)) < 1e-12


def test_synthetic_dataset_is_seeded_and_shaped(gs):
    ds = gs.SyntheticSegDataset(size=(32, 48), num_classes=19, length=4, seed=5)
    a, b = ds[1], ds[1]
    assert torch.equal(a['img'], b['img']) and torch.equal(a['gt_semantic_seg'], b['gt_semantic_seg'])
    assert a['img'].shape == (3, 32, 48) and a['gt_semantic_seg'].shape == (1, 32, 48)
    lab = a['gt_semantic_seg']
    assert set(lab.unique().tolist()) <= set(range(19)) | {255} and 0.02 < float((lab == 255).float().mean()) < 0.25
    dl = gs.build_dataloader(ds, 2, 0, shuffle=False)
    batch = next(iter(dl))
    assert batch['img'].shape == (2, 3, 32, 48) and len(batch['img_metas']) == 2
    perfect = [ds.labels(i) for i in range(4)]
    assert ds.evaluate(perfect)['aAcc'] == 1.0


def test_model_space_manager_roundtrip(gs, tmp_path):
    metas = [{'overhead': {'flops': 1e9 * i, 'params': 1e6}, 'arch': MIN, 'data': {}} for i in range(5)]
    p = tmp_path / 'flops.json'
    gs.ModelSpaceManager(metas).dump(str(p))
    ms = gs.ModelSpaceManager.load(str(p))
    assert len(ms.pack()) == 5 and ms.pack()[2]['overhead']['flops'] == 2e9
    from gaia_seg_b200.model_space import eval_rule
    ms.ms_manager.apply_rule(eval_rule("lambda m: m['overhead.flops'] >= 2e9"))
    assert len(ms.pack()) == 3


@pytest.mark.parametrize('kw', [dict(aux=True, deep_stem=True, os8=True), dict(psp=True, aux=True, os8=True),
                                dict(aspp=True, os8=True), dict()], ids=['fcn_os8', 'psp', 'aspp', 'fcn_os32'])
def test_analytic_complexity_equals_hooked_conv_macs_of_the_oracle(gs, kw):
    """tools/count_flops.py replacement: shapes propagated analytically must give the conv MACs a hook-based counter
    measures on the oracle's real forward pass, for every head type and several sub-nets."""
    from gaia_seg_b200.complexity import conv_macs, get_model_complexity_info
    cfg = C.small_cfg(**kw)
    m = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole'))
    o = O.build_segmentor(cfg)
    o.eval()
    H, W = 64, 96
    for name in ('max', 'min', 'mid'):
        arch = {'backbone': dict(C.SMALL_ARCHS[name]['backbone'])}
        if kw.get('deep_stem'):
            w = arch['backbone']['stem']['width']
            arch['backbone'] = dict(arch['backbone'], stem={'width': [w // 2, w // 2, w]})
        m.manipulate_arch(arch)
        o.manipulate_arch(arch)
        macs = [0]

        def hook(mod, inp, out):
            kh, kw_ = mod.kernel_size
            macs[0] += out.numel() // out.shape[0] * inp[0].shape[1] * kh * kw_

        hs = [c.register_forward_hook(hook) for c in o.modules() if isinstance(c, torch.nn.Conv2d)]
        with torch.no_grad():
            feats = o.backbone(torch.zeros(1, 3, H, W))
            bb_macs = macs[0]
            o.decode_head(feats)
        for h in hs:
            h.remove()
        assert conv_macs(m, (3, H, W), only_backbone=True) == bb_macs
        assert conv_macs(m, (3, H, W)) == macs[0]
        flops, params = get_model_complexity_info(m, (3, H, W))
        assert flops > macs[0] and 0 < params <= sum(p.numel() for p in m.parameters())
    # the MAX sub-net uses every backbone / decode-head parameter
    m.manipulate_arch(C.SMALL_ARCHS['max'] if not kw.get('deep_stem') else
                      {'backbone': dict(C.SMALL_ARCHS['max']['backbone'], stem={'width': [16, 16, 32]})})
    _, params = get_model_complexity_info(m, (3, H, W))
    full = sum(p.numel() for n, p in m.named_parameters() if not n.startswith('auxiliary_head'))
    assert params == full


def test_analytic_macs_reproduce_the_survey_figures(gs):
    """conv MACs per 512x1024 image of the BASELINE supernet (backbone + FCN head): OS32 MAX 176.6 G /
    MIN 27.1 G (plain 7x7 stem); OS8 V1c MAX 938.3 G / MIN 237.1 G plus the deep stem (+2.5 G at full width)."""
    import bench
    from gaia_seg_b200.complexity import conv_macs
    want = {('os32', 'MAX'): 176.6, ('os32', 'MIN'): 27.1, ('os8', 'MAX'): 938.3 + 2.5, ('os8', 'MIN'): 237.1 + 0.4}
    for v in ('os32', 'os8'):
        m = gs.build_segmentor(bench.supernet_cfg(v), train_cfg=dict(), test_cfg=dict(mode='whole'))
        MAX, MIN, _ = bench.sampler_cfg(v)
        for a in (MAX, MIN):
            m.manipulate_arch(gs.fold_dict(a)['arch'])
            got = conv_macs(m, (3, 512, 1024)) / 1e9
            assert abs(got - want[(v, a['name'])]) <= 0.15, (v, a['name'], got)


def test_gradient_exchange_plan_follows_data_dependencies(gs):
    """The overlapped gradient all-reduce is planned from the model structure (runner.FlatParams._plan): the active
    blocks of res stage k and the heads reading that stage's feature are exchanged when StageFn.backward of stage k has
    been enqueued (final by data dependency); blocks beyond the sampled depth are never exchanged (zero on every rank);
    the rest goes after the backward pass.  Host logic only -- the flat offsets are computed as FlatParams does."""
    import gs_checks as C
    from gaia_seg_b200 import functional as Fg
    from gaia_seg_b200.runner import FlatParams, _ALIGN
    model = gs.build_segmentor(C.small_cfg(aux=True), train_cfg=dict(), test_cfg=dict(mode='whole'))
    fp = FlatParams.__new__(FlatParams)
    params = [p for p in model.parameters() if p.requires_grad]
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
    fp.total = total
    fp._sizes = {o: (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN for p, o in zip(params, offs)}
    fp._off_of = {id(p): o for p, o in zip(params, offs)}
    fp._stage_blocks, fp._stage_extra, fp._done, fp._snapshots = {}, {}, [], []
    fp._overlap, fp._check = True, False
    fp._plan(model)
    assert sorted(fp._stage_blocks) == [1, 2, 3, 4]
    # blocks of a stage are contiguous and ordered; stages ascend in the flat buffer
    prev_hi = 0
    for k in (1, 2, 3, 4):
        blocks = fp._stage_blocks[k][1]
        assert len(blocks) == len(getattr(model.backbone, f'layer{k}'))
        assert blocks[0][0] >= prev_hi
        for a, b in zip(blocks, blocks[1:]):
            assert a[1] == b[0]
        prev_hi = blocks[-1][1]
    # decode head reads feature 3 (stage 4), the auxiliary head feature 2 (stage 3)
    assert fp._stage_extra[4] == fp._ranges_of(model.decode_head)
    assert fp._stage_extra[3] == fp._ranges_of(model.auxiliary_head)
    # MIN sub-net: depths [1, 1, 2, 1] of [2, 2, 3, 2] -> the tail blocks are inactive
    model.manipulate_arch(C.SMALL_ARCHS['min'])
    inactive = fp._inactive_ranges()
    want = []
    for k, d in zip((1, 2, 3, 4), (1, 1, 2, 1)):
        blocks = fp._stage_blocks[k][1]
        want.append((blocks[d][0], blocks[-1][1]))
    assert sorted(inactive) == sorted(want)
    # StageFn.backward -> _stage_grads_done -> owner._reduce_stage(k): record what would be exchanged
    calls = []
    Fg_side, fp.peer_grad = Fg.side_stream_run, None
    fp.flat_g = torch.zeros(1)

    class _PG:
        world = 2

        def all_reduce(self, lo, n):
            calls.append((lo, lo + n))

    fp.peer_grad = _PG()
    Fg.side_stream_run = lambda fn, device, keep=(): (fn(), True)[1]
    try:
        for k in (4, 3, 2, 1):
            layer = getattr(model.backbone, f'layer{k}')
            Fg._stage_grads_done([layer[i] for i in range(layer.depth_state)])
    finally:
        Fg.side_stream_run = Fg_side
    b4, b3 = fp._stage_blocks[4][1], fp._stage_blocks[3][1]
    # stage 4: its ONE active block, then (not merged: the inactive block 1 lies between) the decode head it feeds
    assert calls[0] == b4[0] and calls[1] == fp._ranges_of(model.decode_head)[0]
    assert (b3[0][0], b3[1][1]) in calls and fp._ranges_of(model.auxiliary_head)[0] in calls
    # after the backward pass: only the stem is left; nothing overlaps, nothing inactive is exchanged, all active covered
    rest = fp._pending_ranges()
    stem_hi = fp._stage_blocks[1][1][0][0]
    assert rest == [(0, stem_hi)]
    covered = FlatParams._merge(calls + rest + inactive)
    assert covered == [(0, total)]
    for a in calls + rest:
        for b in inactive:
            assert a[1] <= b[0] or b[1] <= a[0]
    # an untagged stage (single GPU, or a model without a plan): nothing happens
    calls.clear()
    Fg._stage_grads_done([torch.nn.Sequential(torch.nn.Conv2d(4, 4, 1))])
    assert calls == []
