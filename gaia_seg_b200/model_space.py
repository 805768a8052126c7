"""Sub-net sampling -- mirror of the `gaiavision.model_space` / `gaiavision.core` pieces the train / test drivers
use ([EXT], contracts from the call sites: tools/train_supernet.py:180-190, gaiaseg/apis/train.py:142-146,
gaiaseg/core/evaluation/cross_arch_eval_hooks.py:50-60, configs/_dynamic_/model_samplers/ar50to101v2.py:2-116):

    build_model_sampler(cfg)  types 'anchor' | 'range' | 'candidate' | 'composite' | 'repeat' | 'concat'
        .sample() -> flat dict with dotted keys      .traverse() -> generator      .set_mode(m)   .anchor_name(i)
    fold_dict / unfold_dict, broadcast_object, ManipulateArchHook, ModelSpaceManager (JSON-lines model space)

Pure host-side integer logic (bit-exact by construction): no tensors are touched here.
"""
import itertools
import json
import pickle
import random

import torch
import torch.distributed as dist

from .core import DynamicMixin, Registry, build_from_cfg

MODEL_SAMPLERS = Registry('model sampler')


def fold_dict(flat, sep='.'):
    """{'arch.backbone.stem.width': 32} -> {'arch': {'backbone': {'stem': {'width': 32}}}}"""
    out = {}
    for key, val in flat.items():
        parts = key.split(sep)
        d = out
        for p in parts[:-1]:
            nxt = d.setdefault(p, {})
            if not isinstance(nxt, dict):
                raise ValueError(f'key "{key}" collides with a leaf value')
            d = nxt
        d[parts[-1]] = val
    return out


def unfold_dict(nested, sep='.', prefix=''):
    out = {}
    for k, v in nested.items():
        key = f'{prefix}{sep}{k}' if prefix else str(k)
        if isinstance(v, dict) and v:
            out.update(unfold_dict(v, sep, key))
        else:
            out[key] = v
    return out


def broadcast_object(obj, src=0, group=None):
    """pickle -> dist.broadcast from `src`, so every rank applies the SAME sub-net.  No-op without a
    process group.  Works on NCCL (byte tensor on the current CUDA device) and gloo."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return obj
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    rank = dist.get_rank(group)
    if rank == src:
        buf = pickle.dumps(obj)
        size = torch.tensor([len(buf)], dtype=torch.int64, device=dev)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(size, src, group=group)
    if rank == src:
        data = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(dev)
    else:
        data = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(data, src, group=group)
    return obj if rank == src else pickle.loads(data.cpu().numpy().tobytes())


class BaseSampler:
    def __init__(self, mode='sample', seed=None):
        self._mode = mode
        self._rng = random.Random(seed)

    def set_mode(self, mode):
        assert mode in ('sample', 'traverse')
        self._mode = mode

    def period(self):
        return 1

    def sample(self):
        raise NotImplementedError

    def traverse(self):
        raise NotImplementedError

    def __call__(self):
        return self.traverse() if self._mode == 'traverse' else self.sample()


def _strip(meta):
    return dict(meta)


@MODEL_SAMPLERS.register_module(name='anchor')
class AnchorSampler(BaseSampler):
    """Fixed sub-nets; `sample()` walks them round-robin (one per iteration)."""

    def __init__(self, anchors, **kw):
        super().__init__(**kw)
        assert len(anchors) > 0
        self.anchors = [dict(a) for a in anchors]
        self._i = 0

    def period(self):
        return len(self.anchors)

    def anchor_name(self, i):
        return self.anchors[i].get('name', str(i))

    def sample(self):
        a = self.anchors[self._i % len(self.anchors)]
        self._i += 1
        return _strip(a)

    def traverse(self):
        for a in self.anchors:
            yield _strip(a)


@MODEL_SAMPLERS.register_module(name='range')
class RangeSampler(BaseSampler):
    """Uniform over start..end step `step` (scalars or per-element lists); `ascending=True` keeps the
    sampled list non-decreasing (ar50to101v2.py:8-14)."""

    def __init__(self, key, start, end, step, ascending=False, **kw):
        super().__init__(**kw)
        self.key, self.ascending = key, ascending
        self.scalar = not isinstance(start, (list, tuple))
        s, e, st = ([start], [end], [step]) if self.scalar else (list(start), list(end), list(step))
        assert len(s) == len(e) == len(st)
        self.grids = [list(range(a, b + 1, c)) for a, b, c in zip(s, e, st)]

    def _ok(self, vals):
        return (not self.ascending) or all(vals[i] <= vals[i + 1] for i in range(len(vals) - 1))

    def sample(self):
        for _ in range(1000):
            vals = [self._rng.choice(g) for g in self.grids]
            if self._ok(vals):
                break
        else:
            vals = sorted(vals)
        return {self.key: vals[0] if self.scalar else vals}

    def traverse(self):
        for vals in itertools.product(*self.grids):
            if self._ok(list(vals)):
                yield {self.key: vals[0] if self.scalar else list(vals)}


@MODEL_SAMPLERS.register_module(name='candidate')
class CandidateSampler(BaseSampler):
    def __init__(self, key, candidates, **kw):
        super().__init__(**kw)
        self.key, self.candidates = key, list(candidates)

    def sample(self):
        return {self.key: self._rng.choice(self.candidates)}

    def traverse(self):
        for c in self.candidates:
            yield {self.key: c}


def _build_children(cfgs, seed):
    out = []
    for i, c in enumerate(cfgs):
        c = dict(c)
        if seed is not None and 'seed' not in c:
            c['seed'] = seed * 1000003 + i + 1
        out.append(build_model_sampler(c))
    return out


@MODEL_SAMPLERS.register_module(name='composite')
class CompositeSampler(BaseSampler):
    """Merge of independent samplers (one key each)."""

    def __init__(self, model_samplers, **kw):
        super().__init__(**kw)
        self.children = _build_children(model_samplers, kw.get('seed'))

    def sample(self):
        out = {}
        for c in self.children:
            out.update(c.sample())
        return out

    def traverse(self):
        for combo in itertools.product(*[list(c.traverse()) for c in self.children]):
            out = {}
            for d in combo:
                out.update(d)
            yield out


@MODEL_SAMPLERS.register_module(name='repeat')
class RepeatSampler(BaseSampler):
    def __init__(self, model_sampler, times=1, **kw):
        super().__init__(**kw)
        self.times = times
        self.child = _build_children([model_sampler], kw.get('seed'))[0]

    def period(self):
        return self.times * self.child.period()

    def sample(self):
        return self.child.sample()

    def traverse(self):
        return self.child.traverse()


@MODEL_SAMPLERS.register_module(name='concat')
class ConcatSampler(BaseSampler):
    """Cycle through the children: child k is asked `period(k)` times in a row.  The shipped train sampler is
    concat[anchor(5), repeat x3(composite(range x3))] = 5 anchors + 3 random sub-nets per cycle; the sandwich
    rule of tools/train_supernet.py:180-187 is concat[anchor(MAX), anchor(MIN), repeat xN(random)]."""

    def __init__(self, model_samplers, **kw):
        super().__init__(**kw)
        self.children = _build_children(model_samplers, kw.get('seed'))
        self._i = 0

    def period(self):
        return sum(c.period() for c in self.children)

    def anchor_name(self, i):
        for c in self.children:
            n = c.period()
            if i < n:
                return c.anchor_name(i) if hasattr(c, 'anchor_name') else str(i)
            i -= n
        return str(i)

    def sample(self):
        i = self._i % self.period()
        self._i += 1
        for c in self.children:
            n = c.period()
            if i < n:
                return c.sample()
            i -= n
        raise AssertionError

    def traverse(self):
        return itertools.chain(*[c.traverse() for c in self.children])


def build_model_sampler(cfg):
    return build_from_cfg(cfg, MODEL_SAMPLERS)


def sandwich_sampler_cfg(max_net, min_net, random_cfg, num_random=2, seed=0):
    """concat[max_net, min_net, random_subnet x N] (tools/train_supernet.py:180-187)."""
    return dict(type='concat', seed=seed, model_samplers=[
        dict(type='anchor', anchors=[dict(max_net)]), dict(type='anchor', anchors=[dict(min_net)]),
        dict(type='repeat', times=num_random, model_sampler=random_cfg)])


class ManipulateArchHook:
    """before_train_iter: meta = broadcast_object(fold_dict(sampler.sample())) ->
    model.manipulate_arch(meta['arch'])  -- one sub-net per iteration (gaiaseg/apis/train.py:142-146;
    mirrors cross_arch_eval_hooks.py:59-60,85-92)."""
    priority = 'NORMAL'

    def __init__(self, model_sampler, broadcast=True):
        self.model_sampler = model_sampler
        self.broadcast = broadcast
        self.last_meta = None

    @staticmethod
    def manipulate_arch(runner, arch_meta):
        model = runner.model
        if isinstance(model, DynamicMixin):
            model.manipulate_arch(arch_meta)
        elif hasattr(model, 'module') and isinstance(model.module, DynamicMixin):
            model.module.manipulate_arch(arch_meta)
        else:
            raise Exception('Current model does not support arch manipulation.')

    def before_train_iter(self, runner):
        meta = fold_dict(self.model_sampler.sample())
        if self.broadcast:
            meta = broadcast_object(meta)
        self.last_meta = meta
        self.manipulate_arch(runner, meta['arch'])


# ------------------------------------------------------------------------------------------------
# model space (JSON-lines table of {overhead:{flops,params}, arch, data[, metric]}; tools/count_flops.py:153-158)
# ------------------------------------------------------------------------------------------------
class ModelSpaceManager:
    def __init__(self, metas=None):
        self.metas = list(metas or [])

    @classmethod
    def load(cls, path_or_list):
        if isinstance(path_or_list, (list, tuple)):
            return cls(path_or_list)
        metas = []
        with open(path_or_list) as f:
            for line in f:
                line = line.strip()
                if line:
                    metas.append(json.loads(line))
        return cls(metas)

    @property
    def ms_manager(self):
        return self

    def apply_rule(self, rule):
        self.metas = list(rule(self.metas)) if rule is not None else self.metas
        return self

    def pack(self):
        return [fold_dict(m) if any('.' in k for k in m) else m for m in self.metas]

    def dump(self, path):
        with open(path, 'w') as f:
            for m in self.metas:
                f.write(json.dumps(m) + '\n')


def eval_rule(func_str=None, sample=None, seed=0):
    """Minimal rule: keep metas for which `func_str` (a lambda over the unfolded meta) holds, then optionally
    sample `sample` of them (configs/_dynamic_/rules/ar50to101v2_rules.py:2-39 uses this vocabulary)."""
    fn = eval(func_str) if isinstance(func_str, str) else func_str  # noqa: S307 - config-provided lambda, as upstream

    def _rule(metas):
        out = [m for m in metas if fn is None or fn(unfold_dict(m) if isinstance(m, dict) else m)]
        if sample is not None and sample < len(out):
            out = random.Random(seed).sample(out, sample)
        return out
    return _rule
