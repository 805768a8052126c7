#!/usr/bin/env python
"""count_flops -- the reference's tools/count_flops.py:58-179: one JSON line `{overhead: {flops, params}, arch, data}` per
sub-net of `cfg.train_sampler` (traverse mode) into <work_dir>/flops.json, the model-space table that
tools/test_supernet.py / finetune_supernet.py read back through ModelSpaceManager.  The reference runs a hook-based
counter (one forward pass per sub-net on the GPU); here the count is analytic (gaia_seg_b200.complexity) and needs no
device.  Same CLI: config, work_dir, --only_backbone_flops, --as_strings, --launcher."""
import argparse
import json
import os
import os.path as osp
import shutil
from collections.abc import Sequence

from _common import setup_dist

import gaia_seg_b200 as gs
from gaia_seg_b200.complexity import get_model_complexity_info


def main():
    p = argparse.ArgumentParser(description='Count flops of each subnet')
    p.add_argument('config')
    p.add_argument('work_dir')
    p.add_argument('--only_backbone_flops', action='store_true')
    p.add_argument('--as_strings', action='store_true')
    p.add_argument('--launcher', choices=['none', 'pytorch'], default='none')
    p.add_argument('--local_rank', type=int, default=0)
    args = p.parse_args()
    cfg = gs.Config.fromfile(args.config)
    cfg.work_dir = args.work_dir
    rank, world = 0, 1
    if args.launcher != 'none':
        setup_dist(args, cfg)
        rank, world = gs.get_dist_info()
    tmpdir = osp.join(cfg.work_dir, '.rank_flops')
    if rank == 0:
        shutil.rmtree(tmpdir, ignore_errors=True)
        os.makedirs(tmpdir)
    model = gs.build_segmentor(cfg.model, train_cfg=cfg.get('train_cfg'), test_cfg=cfg.get('test_cfg'))
    model.eval()
    # the table normally covers the training space; a config whose train sampler is assembled at run time (sandwich
    # rule) names the sub-nets to tabulate in `flops_sampler`, else the validation anchors are used
    sampler = gs.build_model_sampler(cfg.get('flops_sampler') or cfg.get('train_sampler') or cfg.val_sampler)
    sampler.set_mode('traverse')
    metas = []
    for meta in list(sampler.traverse())[rank::world]:
        meta = gs.fold_dict(meta)
        data = meta.get('data') or {'input_shape': 512}
        model.manipulate_arch(meta['arch'])
        shape = data['input_shape']
        if isinstance(shape, str):
            shape = tuple(int(v) for v in shape.strip().split(','))
        elif not isinstance(shape, Sequence):
            shape = (3, 512, 2048)                       # tools/count_flops.py:139-140
        flops, params = get_model_complexity_info(model, tuple(shape), as_strings=args.as_strings,
                                                  only_backbone_flops=args.only_backbone_flops)
        metas.append({'overhead': {'flops': flops, 'params': params}, 'arch': meta['arch'], 'data': data})
    if world > 1:
        import torch.distributed as dist
        with open(osp.join(tmpdir, f'flops.json.{rank}'), 'w') as f:
            for m in metas:
                f.write(json.dumps(m, ensure_ascii=False) + '\n')
        dist.barrier()
        if rank == 0:
            metas = []
            for r in range(world):
                metas += [json.loads(l) for l in open(osp.join(tmpdir, f'flops.json.{r}'))]
    if rank == 0:
        with open(osp.join(cfg.work_dir, 'flops.json'), 'w') as f:
            for m in metas:
                f.write(json.dumps(m, ensure_ascii=False) + '\n')
        shutil.rmtree(tmpdir, ignore_errors=True)
        print(f'{len(metas)} sub-nets -> {osp.join(cfg.work_dir, "flops.json")}')


if __name__ == '__main__':
    main()
