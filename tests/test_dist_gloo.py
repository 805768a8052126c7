"""N > 1 host logic with world_size-2 gloo process groups on CPU: arch broadcast, shared-seed sampling, log-var
reduction, result collection, SyncBN == global-batch BN in the oracle."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f'{fn_name}: worker exit code {p.exitcode}'
    return dict(ret)


def w_broadcast(rank, world):
    import gaia_seg_b200 as gs
    meta = {'arch': {'backbone': {'stem': {'width': 32 + 16 * rank}, 'body': {'width': [48, 96, 192, 384], 'depth': [2, 2, 5, 2 + rank]}}}}
    got = gs.broadcast_object(meta)
    return got


def w_hook(rank, world):
    import gaia_seg_b200 as gs
    import gs_checks as C
    model = gs.build_segmentor(C.small_cfg(), train_cfg=dict(), test_cfg=dict(mode='whole'))
    # every rank seeds its sampler DIFFERENTLY; the hook's broadcast must still make the arch identical
    rnd = dict(type='composite', model_samplers=[
        dict(type='range', key='arch.backbone.stem.width', start=16, end=32, step=16),
        dict(type='range', key='arch.backbone.body.width', start=[16, 32, 48, 64], end=[32, 48, 64, 80], step=[16, 16, 16, 16]),
        dict(type='range', key='arch.backbone.body.depth', start=[1, 1, 2, 1], end=[2, 2, 3, 2], step=[1, 1, 1, 1])])
    sampler = gs.build_model_sampler(dict(rnd, seed=100 + rank))
    hook = gs.ManipulateArchHook(sampler)
    runner = type('R', (), {'model': model})()
    out = []
    for _ in range(5):
        hook.before_train_iter(runner)
        out.append((model.backbone.conv1.width_state, model.backbone.layer3.depth_state, model.backbone.layer4[0].conv3.width_state))
    return out


def w_logvars(rank, world):
    import gaia_seg_b200 as gs
    losses = {'decode.loss_seg': torch.tensor(1.0 + rank), 'decode.acc_seg': torch.tensor(10.0 * (rank + 1)),
              'aux.loss_seg': torch.tensor(0.5)}
    loss, lv = gs.EncoderDecoder._parse_losses(losses)
    return float(loss), dict(lv.items())


def w_collect(rank, world):
    import numpy as np
    from gaia_seg_b200.apis import collect_results_cpu, collect_results_gpu
    part = [np.full((2, 2), 10 * i + rank) for i in range(3)]      # sample index i*world + rank
    a = collect_results_cpu(list(part), 6)
    b = collect_results_gpu(list(part), 6)
    return None if a is None else ([int(x[0, 0]) for x in a], [int(x[0, 0]) for x in b])


def w_syncbn(rank, world):
    """packed (sum, sumsq) all-reduce -> same statistics as BN over the concatenated batch."""
    from oracle import ref_model as O
    g = torch.Generator().manual_seed(0)
    full = torch.randn(2 * world, 6, 5, 7, generator=g)
    x = full[2 * rank:2 * rank + 2]
    stats = torch.cat([x.double().sum((0, 2, 3)), (x.double() ** 2).sum((0, 2, 3))])
    dist.all_reduce(stats)
    n = full.numel() // 6
    mean = stats[:6] / n
    var = stats[6:] / n - mean ** 2
    y = (x - mean.float().view(1, -1, 1, 1)) / torch.sqrt(var.float().view(1, -1, 1, 1) + 1e-5)
    bn = O.DynamicBatchNorm2d(6, affine=True)
    ref = bn(full)[2 * rank:2 * rank + 2]
    return float((y - ref).abs().max())


def test_broadcast_object_makes_ranks_agree():
    out = _run('w_broadcast')
    assert out[0] == out[1] and out[0]['arch']['backbone']['stem']['width'] == 32


def test_manipulate_arch_hook_applies_the_same_subnet_on_every_rank():
    out = _run('w_hook')
    assert out[0] == out[1] and len(set(out[0])) > 1


def test_parse_losses_averages_log_vars_over_ranks():
    out = _run('w_logvars')
    (l0, lv0), (l1, lv1) = out[0], out[1]
    assert abs(l0 - 1.5) < 1e-6 and abs(l1 - 2.5) < 1e-6                     # the loss itself stays local
    assert lv0 == lv1 and abs(lv0['loss'] - 2.0) < 1e-6 and abs(lv0['decode.acc_seg'] - 15.0) < 1e-6


def test_collect_results_interleaves_ranks():
    out = _run('w_collect')
    assert out[1] is None
    assert out[0][0] == [0, 1, 10, 11, 20, 21] and out[0][1] == [0, 1, 10, 11, 20, 21]


def test_packed_stat_allreduce_equals_global_batch_bn():
    out = _run('w_syncbn')
    assert max(out.values()) < 1e-4
