"""Entry points of the hot path, with the reference's names and argument meaning:

    train_segmentor(model, train_sampler, val_sampler, dataset, cfg, distributed, validate, timestamp, meta)
                                                                 gaiaseg/apis/train.py:47-186
    single_gpu_test / multi_gpu_test / collect_results_{cpu,gpu}   gaiaseg/apis/test.py:13-186
    CrossArchEvalHook / DistCrossArchEvalHook                     gaiaseg/core/evaluation/cross_arch_eval_hooks.py:24-167

plus the synthetic Cityscapes / ADE20K-shaped dataset the benchmarks use (no real data in this environment).
"""
import math
import os.path as osp
import pickle
import random
import shutil
import tempfile

import numpy as np
import torch
import torch.distributed as dist
from torch.nn.modules.batchnorm import _BatchNorm
from torch.utils.data import DataLoader, Dataset

from .core import DynamicMixin, Registry, build_from_cfg
from .model_space import ManipulateArchHook, broadcast_object, fold_dict
from .runner import GsDataParallel, Hook, build_optimizer, build_runner, get_dist_info

DATASETS = Registry('dataset')


def set_random_seed(seed, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
@DATASETS.register_module()
class SyntheticSegDataset(Dataset):
    """Seeded random images / label maps with the shapes of the benchmark configs: img ~ N(0,1)
    fp32 [3,H,W]; labels uniform in [0, num_classes) with `ignore_ratio` of the pixels set to 255."""
    CLASSES = None

    def __init__(self, size=(512, 1024), num_classes=19, length=64, ignore_ratio=0.1, seed=1234, test_mode=False,
                 ori_size=None, **kw):
        self.size, self.num_classes, self.length = tuple(size), num_classes, length
        self.ignore_ratio, self.seed, self.test_mode = ignore_ratio, seed, test_mode
        self.ori_size = tuple(ori_size) if ori_size is not None else self.size
        self.CLASSES = tuple(str(i) for i in range(num_classes))

    def __len__(self):
        return self.length

    def _meta(self):
        H, W = self.size
        return dict(ori_shape=(*self.ori_size, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=1.0,
                    flip=False, flip_direction='horizontal')

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 100003 + idx)
        H, W = self.size
        img = torch.randn(3, H, W, generator=g)
        if self.test_mode:
            return dict(img=[img], img_metas=[self._meta()])
        lab = torch.randint(0, self.num_classes, (1, H, W), generator=g)
        if self.ignore_ratio > 0:
            lab[torch.rand(1, H, W, generator=g) < self.ignore_ratio] = 255
        return dict(img=img, img_metas=self._meta(), gt_semantic_seg=lab)

    def labels(self, idx):
        g = torch.Generator().manual_seed(self.seed * 100003 + idx)
        H, W = self.size
        torch.randn(3, H, W, generator=g)
        lab = torch.randint(0, self.num_classes, (1, H, W), generator=g)
        if self.ignore_ratio > 0:
            lab[torch.rand(1, H, W, generator=g) < self.ignore_ratio] = 255
        return lab[0].numpy()

    def evaluate(self, results, metric='mIoU', logger=None, **kw):
        """mIoU / aAcc of predicted label maps against the synthetic labels (mmseg CustomDataset.evaluate)."""
        K = self.num_classes
        inter, union, correct, total = np.zeros(K), np.zeros(K), 0, 0
        for i, pred in enumerate(results):
            gt = self.labels(i)
            if pred.shape != gt.shape:
                continue
            mask = gt != 255
            p, g_ = pred[mask], gt[mask]
            correct += int((p == g_).sum())
            total += int(mask.sum())
            hit = p[p == g_]
            ai = np.bincount(hit, minlength=K)[:K]
            ap = np.bincount(p, minlength=K)[:K]
            ag = np.bincount(g_, minlength=K)[:K]
            inter += ai
            union += ap + ag - ai
        iou = inter / np.maximum(union, 1)
        return dict(mIoU=float(iou.mean()), aAcc=float(correct / max(total, 1)))


def build_dataset(cfg, default_args=None):
    return build_from_cfg(cfg, DATASETS, default_args)


def _collate_train(batch):
    return dict(img=torch.stack([b['img'] for b in batch]), img_metas=[b['img_metas'] for b in batch],
                gt_semantic_seg=torch.stack([b['gt_semantic_seg'] for b in batch]))


def _collate_test(batch):
    n_aug = len(batch[0]['img'])
    return dict(img=[torch.stack([b['img'][a] for b in batch]) for a in range(n_aug)],
                img_metas=[[b['img_metas'][a] for b in batch] for a in range(n_aug)])


def build_dataloader(dataset, samples_per_gpu, workers_per_gpu=0, num_gpus=1, dist=False, shuffle=True, seed=None,
                     drop_last=False, pin_memory=True, **kw):
    sampler = None
    if dist:
        rank, world = get_dist_info()
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, world, rank, shuffle=shuffle,
                                                                  seed=seed or 0)
        shuffle = False
    test = getattr(dataset, 'test_mode', False)
    return DataLoader(dataset, batch_size=samples_per_gpu, sampler=sampler, shuffle=shuffle and not test,
                      num_workers=workers_per_gpu, collate_fn=_collate_test if test else _collate_train,
                      pin_memory=pin_memory and torch.cuda.is_available(), drop_last=drop_last)


# ------------------------------------------------------------------------------------------------
# test loops
# ------------------------------------------------------------------------------------------------
def _to_device(data):
    dev = torch.device('cuda', torch.cuda.current_device())
    return dict(img=[t.to(dev, non_blocking=True) for t in data['img']], img_metas=data['img_metas'])


def single_gpu_test(model, data_loader, show=False, out_dir=None):
    """model(return_loss=False, rescale=True, **data) under no_grad (gaiaseg/apis/test.py:13-65)."""
    model.eval()
    results = []
    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, rescale=True, **_to_device(data))
        results.extend(result if isinstance(result, list) else [result])
    return results


def multi_gpu_test(model, data_loader, tmpdir=None, gpu_collect=False):
    """As the reference (gaiaseg/apis/test.py:68-109) the model is NOT switched to eval here -- `model.eval()`
    is commented out upstream (:85) so BN calibration modes keep working; the caller decides."""
    results = []
    for data in data_loader:
        with torch.no_grad():
            result = model(return_loss=False, rescale=True, **_to_device(data))
        results.extend(result if isinstance(result, list) else [result])
    if gpu_collect:
        return collect_results_gpu(results, len(data_loader.dataset))
    return collect_results_cpu(results, len(data_loader.dataset), tmpdir)


def collect_results_cpu(result_part, size, tmpdir=None):
    rank, world = get_dist_info()
    if world == 1:
        return result_part[:size]
    if tmpdir is None:
        tmpdir = broadcast_object(tempfile.mkdtemp() if rank == 0 else None)
    with open(osp.join(tmpdir, f'part_{rank}.pkl'), 'wb') as f:
        pickle.dump(result_part, f)
    dist.barrier()
    if rank != 0:
        return None
    parts = []
    for i in range(world):
        with open(osp.join(tmpdir, f'part_{i}.pkl'), 'rb') as f:
            parts.append(pickle.load(f))
    ordered = [r for group in zip(*parts) for r in group]
    for p in parts:  # ragged tail
        ordered.extend(p[min(len(q) for q in parts):])
    shutil.rmtree(tmpdir, ignore_errors=True)
    return ordered[:size]


def collect_results_gpu(result_part, size):
    rank, world = get_dist_info()
    if world == 1:
        return result_part[:size]
    parts = [None] * world
    dist.all_gather_object(parts, result_part)
    if rank != 0:
        return None
    ordered = [r for group in zip(*parts) for r in group]
    return ordered[:size]


# ------------------------------------------------------------------------------------------------
# cross-arch evaluation hooks
# ------------------------------------------------------------------------------------------------
class CrossArchEvalHook(Hook):
    """Every `interval` iterations: for meta in val_sampler.traverse(): broadcast -> manipulate_arch -> test ->
    dataset.evaluate (cross_arch_eval_hooks.py:44-92)."""

    def __init__(self, dataloader, model_sampler=None, interval=1, by_epoch=False, **eval_kwargs):
        assert model_sampler is not None, 'In cross arch mode, the val sampler should be specified in cfg'
        self.dataloader, self.model_sampler = dataloader, model_sampler
        self.interval, self.by_epoch, self.eval_kwargs = interval, by_epoch, eval_kwargs
        self.results = {}

    def _test(self, runner):
        return single_gpu_test(runner.model, self.dataloader)

    def after_train_iter(self, runner):
        if self.by_epoch or not self.every_n_iters(runner, self.interval):
            return
        if not hasattr(self.model_sampler, 'traverse'):
            raise AttributeError(f'{type(self.model_sampler)} has no attribute `traverse`')
        for i, meta in enumerate(self.model_sampler.traverse()):
            anchor_id = self.model_sampler.anchor_name(i) if hasattr(self.model_sampler, 'anchor_name') else i
            meta = broadcast_object(fold_dict(meta))
            ManipulateArchHook.manipulate_arch(runner, meta['arch'])
            results = self._test(runner)
            if runner.rank == 0 and results is not None:
                res = self.dataloader.dataset.evaluate(results, **self.eval_kwargs)
                self.results[anchor_id] = res
                print(f'[eval iter {runner.iter + 1}] {anchor_id}: {res}', flush=True)
        if dist.is_available() and dist.is_initialized():
            dist.barrier()        # rank 0 alone ran dataset.evaluate: re-align before the next training iteration
        runner.model.train()


class DistCrossArchEvalHook(CrossArchEvalHook):
    def __init__(self, dataloader, model_sampler=None, interval=1, by_epoch=False, tmpdir=None, gpu_collect=False,
                 **eval_kwargs):
        super().__init__(dataloader, model_sampler, interval, by_epoch, **eval_kwargs)
        self.tmpdir, self.gpu_collect = tmpdir, gpu_collect

    def _test(self, runner):
        runner.model.eval()
        return multi_gpu_test(runner.model, self.dataloader, tmpdir=self.tmpdir, gpu_collect=self.gpu_collect)


# ------------------------------------------------------------------------------------------------
# train_segmentor
# ------------------------------------------------------------------------------------------------
def train_segmentor(model, train_sampler, val_sampler, dataset, cfg, distributed=False, validate=False,
                    timestamp=None, meta=None):
    dataset = dataset if isinstance(dataset, (list, tuple)) else [dataset]
    data_loaders = [build_dataloader(ds, cfg.data.samples_per_gpu, cfg.data.workers_per_gpu,
                                     len(cfg.get('gpu_ids', [0])), dist=distributed, seed=cfg.get('seed'),
                                     drop_last=True) for ds in dataset]
    model = GsDataParallel(model.cuda(), device_ids=[torch.cuda.current_device()], broadcast_buffers=False,
                           find_unused_parameters=cfg.get('find_unused_parameters', True))
    from .runner import reserve_activation_pool
    reserve_activation_pool(cfg.get('activation_pool_gb', 32.0))
    _, world_size = get_dist_info()
    lr_scaler_config = cfg.get('lr_scaler', None)
    if lr_scaler_config is not None:
        total_batch_size = world_size * cfg.data.samples_per_gpu
        base_lr = lr_scaler_config['base_lr']
        if lr_scaler_config.get('policy', 'linear') == 'linear':
            cfg.optimizer.lr = base_lr * total_batch_size
        else:
            cfg.optimizer.lr = base_lr * math.pow(total_batch_size, lr_scaler_config.get('temperature', 0.7))
    optimizer = build_optimizer(model, cfg.optimizer)
    if cfg.get('runner') is None:
        cfg.runner = {'type': 'IterBasedRunner', 'max_iters': cfg.total_iters}
    runner = build_runner(cfg.runner, default_args=dict(model=model, batch_processor=None, optimizer=optimizer,
                                                        work_dir=cfg.get('work_dir'), logger=None, meta=meta))
    runner.register_training_hooks(cfg.get('lr_config'), cfg.get('optimizer_config'), cfg.get('checkpoint_config'),
                                   cfg.get('log_config'), cfg.get('momentum_config', None))
    runner.timestamp = timestamp
    if cfg.get('manipulate_arch', True):
        runner.register_hook(ManipulateArchHook(train_sampler))
    if validate:
        eval_cfg = dict(cfg.get('evaluation', {}))
        eval_cfg['by_epoch'] = False
        vals = cfg.data.val if isinstance(cfg.data.val, (list, tuple)) else [cfg.data.val]
        for each in vals:
            val_dataset = build_dataset(each, dict(test_mode=True))
            val_loader = build_dataloader(val_dataset, 1, cfg.data.workers_per_gpu, dist=distributed, shuffle=False)
            hook = DistCrossArchEvalHook if distributed else CrossArchEvalHook
            runner.register_hook(hook(val_loader, val_sampler, **eval_cfg))
    if cfg.get('resume_from'):
        runner.resume(cfg.resume_from)
    elif cfg.get('load_from'):
        runner.load_checkpoint(cfg.load_from)
    calib_bn = cfg.get('caliberate_bn', None)
    if calib_bn and calib_bn.get('reset_stats', False):
        def clean_bn_stats(m):
            if isinstance(m, _BatchNorm):
                m.running_mean.zero_()
                m.running_var.fill_(1)
        model.apply(clean_bn_stats)
    runner.run(data_loaders, cfg.get('workflow'))
    return runner
