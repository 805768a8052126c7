#!/usr/bin/env python
"""train_supernet -- same CLI and flow as the reference's tools/train_supernet.py:99-214:
cfg -> build_segmentor -> build_model_sampler(train_sampler / val_sampler) -> build_dataset -> train_segmentor.
`use_distillation` builds the sandwich sampler concat[max_net, min_net, random_subnet x N] (:180-187; the upstream
NameError on `sample_subnet_num` is not reproduced)."""
import argparse
import os
import os
import os.path as osp
import time

from _common import DictAction, setup_dist

import gaia_seg_b200 as gs


def parse_args():
    p = argparse.ArgumentParser(description='Train a supernet segmentor')
    p.add_argument('config')
    p.add_argument('--work-dir')
    p.add_argument('--resume-from')
    p.add_argument('--load-from')
    p.add_argument('--no-validate', action='store_true')
    g = p.add_mutually_exclusive_group()
    g.add_argument('--gpus', type=int)
    g.add_argument('--gpu-ids', type=int, nargs='+')
    p.add_argument('--seed', type=int, default=None)
    p.add_argument('--deterministic', action='store_true')
    p.add_argument('--options', nargs='+', action=DictAction)
    p.add_argument('--cfg-options', nargs='+', action=DictAction)
    p.add_argument('--launcher', choices=['none', 'pytorch', 'slurm', 'mpi'], default='none')
    p.add_argument('--local_rank', type=int, default=0)
    args = p.parse_args()
    if args.options and args.cfg_options:
        raise ValueError('--options and --cfg-options cannot be both specified')
    args.cfg_options = args.cfg_options or args.options
    return args


def main():
    args = parse_args()
    cfg = gs.Config.fromfile(args.config)
    if args.cfg_options:
        cfg.merge_from_dict(args.cfg_options)
    cfg.work_dir = args.work_dir or cfg.get('work_dir') or osp.join('./work_dirs', osp.splitext(osp.basename(args.config))[0])
    if args.resume_from:
        cfg.resume_from = args.resume_from
    if args.load_from:
        cfg.load_from = args.load_from
    cfg.gpu_ids = args.gpu_ids if args.gpu_ids is not None else list(range(args.gpus or 1))
    distributed = setup_dist(args, cfg)
    os.makedirs(cfg.work_dir, exist_ok=True)
    timestamp = time.strftime('%Y%m%d_%H%M%S', time.localtime())
    if args.seed is not None:
        gs.set_random_seed(args.seed, deterministic=args.deterministic)
    cfg.seed = args.seed
    model = gs.build_segmentor(cfg.model, train_cfg=cfg.get('train_cfg'), test_cfg=cfg.get('test_cfg'))
    if cfg.get('use_distillation', False) or cfg.get('sandwich', False):
        n = cfg.get('sample_subnet_num', 2)
        cfg.train_sampler = gs.sandwich_sampler_cfg(cfg.max_net, cfg.min_net, cfg.random_subnet, n, seed=cfg.get('sampler_seed', 0))
    train_sampler = gs.build_model_sampler(cfg.train_sampler)
    val_sampler = gs.build_model_sampler(cfg.val_sampler) if cfg.get('val_sampler') else None
    datasets = [gs.build_dataset(cfg.data.train)]
    meta = dict(config=args.config, CLASSES=datasets[0].CLASSES, seed=args.seed)
    gs.train_segmentor(model, train_sampler, val_sampler, datasets, cfg, distributed=distributed,
                       validate=not args.no_validate and val_sampler is not None, timestamp=timestamp, meta=meta)


if __name__ == '__main__':
    main()
