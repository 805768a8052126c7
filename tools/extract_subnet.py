#!/usr/bin/env python
"""extract_subnet -- the reference's tools/extract_subnet.py:65-152: build + load + eval() + deploy(); per sampler
anchor: manipulate_arch -> deepcopy -> one dummy forward (every DynConv2d / DynBN slices its tensors, every stage
drops unused blocks) -> save_checkpoint(md5(meta)[:8].pth).  All norm cfgs are switched to DynBN first (:54-62)."""
import argparse
import copy
import hashlib
import json
import os
import os.path as osp

import torch

from _common import DictAction, setup_dist

import gaia_seg_b200 as gs


def prepare_cfg(cfg):
    def swap(d):
        if isinstance(d, dict):
            if d.get('type') in ('DynSyncBN', 'SyncBN'):
                d['type'] = 'DynBN'
                d.pop('group_size', None)
            for v in d.values():
                swap(v)
    swap(cfg.model)
    return cfg


def main():
    p = argparse.ArgumentParser(description='Extract sub-nets from a supernet checkpoint')
    p.add_argument('src_ckpt')
    p.add_argument('out_dir')
    p.add_argument('config')
    p.add_argument('--input-shape', type=int, nargs=2, default=[128, 256])
    p.add_argument('--cfg-options', nargs='+', action=DictAction)
    p.add_argument('--local_rank', type=int, default=0)
    args = p.parse_args()
    cfg = prepare_cfg(gs.Config.fromfile(args.config))
    if args.cfg_options:
        cfg.merge_from_dict(args.cfg_options)
    model = gs.build_segmentor(cfg.model, train_cfg=None, test_cfg=cfg.get('test_cfg'))
    if args.src_ckpt != 'none':
        gs.load_checkpoint(model, args.src_ckpt, map_location='cpu')
    model = model.cuda().eval()
    model.deploy()
    sampler = gs.build_model_sampler(cfg.get('extract_sampler', cfg.get('val_sampler')))
    sampler.set_mode('traverse')
    os.makedirs(args.out_dir, exist_ok=True)
    H, W = args.input_shape
    for meta in sampler.traverse():
        meta = gs.fold_dict(meta)
        model.manipulate_arch(meta['arch'])
        sub = copy.deepcopy(model)
        with torch.no_grad():
            sub(return_loss=False, img=[torch.zeros(1, 3, H, W, device='cuda')],
                img_metas=[[dict(ori_shape=(H, W, 3), flip=False)]])
        name = hashlib.md5(json.dumps(meta, sort_keys=True).encode()).hexdigest()[:8]
        gs.save_checkpoint(sub, osp.join(args.out_dir, f'{name}.pth'), meta=dict(meta=meta))
        print(f'{name}.pth  {sum(p_.numel() for p_ in sub.parameters()) / 1e6:.2f} M params  {meta.get("name", "")}')


if __name__ == '__main__':
    main()
