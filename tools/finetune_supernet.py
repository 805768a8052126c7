#!/usr/bin/env python
"""finetune_supernet -- the reference's tools/finetune_supernet.py:139-366: for every rule-selected model_meta:
one-anchor sampler with that single arch -> train_segmentor -> reload latest.pth -> test -> metrics json."""
import argparse
import os
import json
import os
import os.path as osp

from _common import DictAction, setup_dist

import gaia_seg_b200 as gs


def main():
    p = argparse.ArgumentParser(description='Finetune sub-nets of a supernet')
    p.add_argument('config')
    p.add_argument('--work-dir')
    p.add_argument('--load-from')
    p.add_argument('--model-space-path', dest='model_space_path')
    p.add_argument('--tmpdir')
    p.add_argument('--metric-tag', default='finetune')
    p.add_argument('--out-name', default='metrics.json')
    p.add_argument('--eval', type=str, nargs='+', default=['mIoU'])
    p.add_argument('--seed', type=int, default=None)
    p.add_argument('--cfg-options', nargs='+', action=DictAction)
    p.add_argument('--launcher', choices=['none', 'pytorch', 'slurm', 'mpi'], default='none')
    p.add_argument('--local_rank', type=int, default=0)
    args = p.parse_args()
    cfg = gs.Config.fromfile(args.config)
    if args.cfg_options:
        cfg.merge_from_dict(args.cfg_options)
    cfg.work_dir = args.work_dir or cfg.get('work_dir', './work_dirs/finetune')
    distributed = setup_dist(args, cfg)
    rank, _ = gs.get_dist_info()
    if args.seed is not None:
        gs.set_random_seed(args.seed)
    if args.model_space_path:
        metas = gs.ModelSpaceManager.load(args.model_space_path).pack()
    else:
        metas = [gs.fold_dict(m) for m in gs.build_model_sampler(cfg.val_sampler).traverse()]
    results = []
    for i, meta in enumerate(metas):
        meta = gs.broadcast_object(meta)
        sub_dir = osp.join(cfg.work_dir, f'subnet_{i}')
        cfg.work_dir, base_dir = sub_dir, cfg.work_dir
        model = gs.build_segmentor(cfg.model, train_cfg=cfg.get('train_cfg'), test_cfg=cfg.get('test_cfg'))
        if args.load_from:
            gs.load_checkpoint(model, args.load_from, map_location='cpu')
        anchor = gs.build_model_sampler(dict(type='anchor', anchors=[gs.unfold_dict(meta)]))
        dataset = gs.build_dataset(cfg.data.train)
        gs.train_segmentor(model, anchor, anchor, [dataset], cfg, distributed=distributed, validate=False,
                           meta=dict(CLASSES=dataset.CLASSES))
        test_set = gs.build_dataset(cfg.data.test, dict(test_mode=True))
        loader = gs.build_dataloader(test_set, 1, 0, dist=distributed, shuffle=False)
        model.manipulate_arch(meta['arch'])
        model.eval()
        outputs = gs.multi_gpu_test(model, loader, args.tmpdir) if distributed else gs.single_gpu_test(model, loader)
        if outputs is not None:
            meta = dict(meta)
            meta.setdefault('metric', {})[args.metric_tag] = test_set.evaluate(outputs, metric=args.eval)
            results.append(meta)
        cfg.work_dir = base_dir
    if rank == 0:
        os.makedirs(cfg.work_dir, exist_ok=True)
        with open(osp.join(cfg.work_dir, args.out_name), 'w') as f:
            for m in results:
                f.write(json.dumps(m) + '\n')


if __name__ == '__main__':
    main()
