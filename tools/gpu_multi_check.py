"""Multi-GPU parity (run under torchrun, one rank per GPU): SyncBN over R ranks + gradient all-reduce must equal the
oracle on the concatenated global batch.  Prints one JSON line from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import gaia_seg_b200 as gs  # noqa: E402
import gs_checks as C  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg = C.small_cfg(aux=True, deep_stem=True, os8=True)
    om, gm, _ = C.build_pair(gs, cfg)
    arch = {'backbone': {'stem': {'width': [8, 8, 16]}, 'body': {'width': [16, 32, 48, 64], 'depth': [1, 1, 2, 1]}}}
    om.manipulate_arch(arch)
    gm.manipulate_arch(arch)
    g = torch.Generator().manual_seed(21)
    img = C.bf16r(torch.randn(2 * world, 3, 64, 96, generator=g))
    lab = C._labels(g, 2 * world, 19, 64, 96)
    sd0 = {k: v.clone() for k, v in om.state_dict().items()}
    C.emulate_bf16_storage(om)
    om = om.double()
    om.train()
    lo = om.forward_train(img.double(), None, lab)
    loss_o = om.parse_losses(lo)
    loss_o.backward()
    opt = gs.GsSGD(gm, lr=0.01, momentum=0.9, weight_decay=5e-4)
    gm.train()
    sl = slice(2 * rank, 2 * rank + 2)
    out = gm.train_step(dict(img=img[sl].cuda(), img_metas=[{}, {}], gt_semantic_seg=lab[sl].cuda()), opt)
    opt.zero_grad()
    out['loss'].backward()
    w = opt.flat.all_reduce_grads()
    torch.cuda.synchronize()
    res = {}
    res['world'] = w
    res['loss_mean_over_ranks'] = out['log_vars']['loss']
    res['loss_oracle_global_batch'] = float(loss_o)
    g_cuda = {n: (p.grad.detach() / w).cpu() for n, p in gm.named_parameters() if p.grad is not None}
    g_or = {n: p.grad.detach() for n, p in om.named_parameters() if p.grad is not None}
    d = C._grad_cos(g_cuda, g_or)
    res['grad_mean_1mcos'] = sum(d.values()) / len(d)
    res['grad_worst'] = max(d.items(), key=lambda kv: kv[1])
    rm = {n: b for n, b in gm.named_buffers() if n.endswith('running_mean')}
    rmo = dict(om.named_buffers())
    res['running_mean_err'] = max(float((b.cpu().double() - rmo[n]).abs().max()) for n, b in rm.items())
    # buffers must be bit-identical on every rank (broadcast_buffers=False relies on it)
    flat = torch.cat([b.flatten() for b in rm.values()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    res['buffers_identical_across_ranks'] = bool(torch.equal(flat, ref))
    opt.grad_scale = 1.0 / w
    opt.step()
    pflat = opt.flat.flat_p.clone()
    pref = pflat.clone()
    dist.broadcast(pref, 0)
    res['params_identical_after_step'] = bool(torch.equal(pflat, pref))
    ok = (abs(res['loss_mean_over_ranks'] - res['loss_oracle_global_batch']) < 2e-3 * abs(res['loss_oracle_global_batch'])
          and res['grad_mean_1mcos'] < 0.05 and res['running_mean_err'] < 2e-3
          and res['buffers_identical_across_ranks'] and res['params_identical_after_step'])
    # CUDA vs CUDA: collectives bit-exact against ordered sums, N ranks vs ONE rank on the concatenated batch (tight),
    # overlapped vs plain vs NCCL all-reduce (tools/multi_rank_parity.py)
    from tools.multi_rank_parity import run as cuda_vs_cuda
    res['vs_1rank'] = cuda_vs_cuda(gs)
    ok = ok and res['vs_1rank']['ok']
    res['ok'] = bool(ok)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
