#!/usr/bin/env python
"""test_supernet -- the search sweep of the reference's tools/test_supernet.py:131-327: load checkpoint -> model
space (JSON lines) -> rules -> for each model_meta: broadcast, manipulate_arch(meta['arch']), multi_gpu_test,
dataset.evaluate, append metric -> dump metrics json.  `caliberate_bn.use_minibatch_stats` drops the running
statistics (:190-198).  Sub-nets can also be sharded across ranks (--shard-subnets): independent units, no
communication, the natural multi-GPU mode for a 50-sub-net sweep."""
import argparse
import os
import json
import os
import os.path as osp
import time

import torch
from torch.nn.modules.batchnorm import _BatchNorm

from _common import DictAction, setup_dist

import gaia_seg_b200 as gs


def parse_args():
    p = argparse.ArgumentParser(description='Evaluate sub-nets of a supernet')
    p.add_argument('config')
    p.add_argument('checkpoint')
    p.add_argument('--model-space-path', dest='model_space_path')
    p.add_argument('--work-dir')
    p.add_argument('--out-name', default='metrics.json')
    p.add_argument('--save-results', action='store_true')
    p.add_argument('--eval', type=str, nargs='+', default=['mIoU'])
    p.add_argument('--gpu-collect', action='store_true')
    p.add_argument('--tmpdir')
    p.add_argument('--metric-tag', default='test')
    p.add_argument('--shard-subnets', action='store_true')
    p.add_argument('--cfg-options', nargs='+', action=DictAction)
    p.add_argument('--launcher', choices=['none', 'pytorch', 'slurm', 'mpi'], default='none')
    p.add_argument('--local_rank', type=int, default=0)
    return p.parse_args()


def main():
    args = parse_args()
    cfg = gs.Config.fromfile(args.config)
    if args.cfg_options:
        cfg.merge_from_dict(args.cfg_options)
    distributed = setup_dist(args, cfg)
    rank, world = gs.get_dist_info()
    model = gs.build_segmentor(cfg.model, train_cfg=None, test_cfg=cfg.get('test_cfg'))
    if args.checkpoint and args.checkpoint != 'none':
        gs.load_checkpoint(model, args.checkpoint, map_location='cpu')
    calib = cfg.get('caliberate_bn', None)
    if calib and calib.get('use_minibatch_stats', False):
        for m in model.modules():
            if isinstance(m, _BatchNorm):
                m.track_running_stats = False
                m.running_mean = m.running_var = None
    model = model.cuda()
    if args.model_space_path:
        ms = gs.ModelSpaceManager.load(args.model_space_path)
        metas = ms.pack()
    else:
        sampler = gs.build_model_sampler(cfg.val_sampler)
        metas = [gs.fold_dict(m) for m in sampler.traverse()]
        ms = gs.ModelSpaceManager(metas)
    dataset = gs.build_dataset(cfg.data.test, dict(test_mode=True))
    results_all = []
    for i, meta in enumerate(metas):
        if args.shard_subnets and i % world != rank:
            continue
        meta = meta if args.shard_subnets else gs.broadcast_object(meta)
        loader = gs.build_dataloader(dataset, 1, cfg.data.get('workers_per_gpu', 0), dist=distributed and not args.shard_subnets,
                                     shuffle=False)
        model.manipulate_arch(meta['arch'])
        model.eval()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if distributed and not args.shard_subnets:
            outputs = gs.multi_gpu_test(model, loader, args.tmpdir, args.gpu_collect)
        else:
            outputs = gs.single_gpu_test(model, loader)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if outputs is not None:
            metric = dataset.evaluate(outputs, metric=args.eval)
            meta = dict(meta)
            meta.setdefault('metric', {})[args.metric_tag] = metric
            meta['imgs_per_s'] = len(dataset) / dt
            results_all.append(meta)
            print(f'[{i + 1}/{len(metas)}] {metric} {len(dataset) / dt:.1f} img/s', flush=True)
    if args.shard_subnets and world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, results_all)
        results_all = [m for part in gathered for m in part]
    if rank == 0 and args.work_dir:
        os.makedirs(args.work_dir, exist_ok=True)
        with open(osp.join(args.work_dir, args.out_name), 'w') as f:
            for m in results_all:
                f.write(json.dumps(m) + '\n')


if __name__ == '__main__':
    main()
