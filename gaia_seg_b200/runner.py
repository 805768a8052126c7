"""Training plumbing around the hot path: config loading, the data-parallel wrapper, the fused SGD optimizer,
an iteration-based runner with the hooks the reference registers (gaiaseg/apis/train.py:118-186: poly LR,
optimizer, checkpoint, text logger, ManipulateArchHook, eval hooks) and checkpoint IO in the reference's
format (max-width OIHW fp32 state_dict + meta).

B200-first choices:
  * all parameters live in ONE flat fp32 master buffer (+ one flat gradient buffer, one flat momentum buffer,
    one flat bf16 forward-shadow buffer at the same offsets) -> the optimizer is ONE kernel launch, the
    gradient exchange is a hand-written two-shot all-reduce over NVLink PEER MEMORY of ranges of the flat buffer
    (gs_grad_allreduce; per res stage on a side stream during backward, blocks outside the sampled depth never travel;
    bucketed NCCL only as the collectively agreed fallback) instead of DDP's reducer hooks + `find_unused_parameters`
    graph walk (gaiaseg/apis/train.py:88-95);
  * SyncBN statistics are packed (sum, sumsq) exchanges over peer memory issued by the layers themselves.
"""
import importlib.util
import math
import os
import os.path as osp
import time
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as F_gs
from ._lib import call
from .core import DynamicConv2d


# ------------------------------------------------------------------------------------------------
# Config (mmcv.Config subset: python files, `_base_` inheritance, attribute access, dotted overrides)
# ------------------------------------------------------------------------------------------------
class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _to_cfgdict(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: _to_cfgdict(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cfgdict(v) for v in obj)
    return obj


def _merge(base, new):
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            out[k] = v
    return out


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, '_cfg_dict', _to_cfgdict(cfg_dict or {}))
        object.__setattr__(self, 'filename', filename)

    @staticmethod
    def _file2dict(filename):
        filename = osp.abspath(osp.expanduser(filename))
        if not osp.isfile(filename):
            raise FileNotFoundError(filename)
        spec = importlib.util.spec_from_file_location('_gs_cfg_' + str(abs(hash(filename))), filename)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        cfg = {k: v for k, v in vars(mod).items()
               if not k.startswith('__') and not isinstance(v, type(os)) and not callable(v)}
        base = cfg.pop('_base_', None)
        if base is not None:
            merged = {}
            for b in (base if isinstance(base, (list, tuple)) else [base]):
                merged = _merge(merged, Config._file2dict(osp.join(osp.dirname(filename), b)))
            cfg = _merge(merged, cfg)
        return cfg

    @staticmethod
    def fromfile(filename):
        return Config(Config._file2dict(filename), filename)

    def merge_from_dict(self, options):
        for key, val in options.items():
            d = self._cfg_dict
            parts = key.split('.')
            for p in parts[:-1]:
                d = d.setdefault(p, ConfigDict())
            d[parts[-1]] = _to_cfgdict(val)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _to_cfgdict(value)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def to_dict(self):
        return dict(self._cfg_dict)


# ------------------------------------------------------------------------------------------------
# distributed helpers
# ------------------------------------------------------------------------------------------------
def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_dist(launcher='pytorch', backend='nccl', **kwargs):
    """mmcv.runner.init_dist for the pytorch launcher (tools/train_supernet.py:138): env:// rendezvous."""
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', rank % max(torch.cuda.device_count(), 1)))
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
        kwargs.setdefault('device_id', torch.device('cuda', local_rank))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    dist.init_process_group(backend=backend, **kwargs)


# ------------------------------------------------------------------------------------------------
# flat parameter storage + fused optimizer
# ------------------------------------------------------------------------------------------------
_ALIGN = 64  # elements; keeps every tensor 256-byte (fp32) / 128-byte (bf16 shadow) aligned


class FlatParams:
    """Moves every trainable parameter of `model` into one flat fp32 buffer (views keep the logical shapes and
    the KRSC memory order of conv weights), with matching flat gradient and bf16 shadow buffers."""

    def __init__(self, model):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError('model has no trainable parameters')
        dev = params[0].device
        if dev.type != 'cuda':
            raise F_gs.GsError('FlatParams: the model must be on a CUDA device (call model.cuda() first)')
        for m in model.modules():
            if isinstance(m, DynamicConv2d):
                F_gs.ensure_krsc(m)
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        # several ranks: the gradient buffer lives in IPC-shareable memory so that the all-reduce can run over NVLink
        # peer memory (functional.PeerGrad); the ranks must construct FlatParams collectively
        self.peer_grad = F_gs.PeerGrad.create(total, dev)
        self.flat_g = self.peer_grad.tensor if self.peer_grad else torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_shadow = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.params, self.offsets = params, offs
        for p, o in zip(params, offs):
            n = p.numel()
            if p.dim() == 4:  # conv weight: memory order [Co][kh][kw][Ci]
                Co, Ci, kh, kw = p.shape
                vp = self.flat_p[o:o + n].view(Co, kh, kw, Ci).permute(0, 3, 1, 2)
                vg = self.flat_g[o:o + n].view(Co, kh, kw, Ci).permute(0, 3, 1, 2)
            else:
                vp = self.flat_p[o:o + n].view(p.shape)
                vg = self.flat_g[o:o + n].view(p.shape)
            vp.copy_(p.data)
            p.data = vp
            p.grad = vg
            p._gs_flat_off = o
        self.convs = []
        off_of = {id(p): o for p, o in zip(params, offs)}
        for m in model.modules():
            if isinstance(m, DynamicConv2d) and id(m.weight) in off_of:
                self.convs.append(m)
                o, n = off_of[id(m.weight)], m.weight.numel()
                if not F_gs.is_image_conv(m):
                    m._gs_w_krsc = self.flat_shadow[o:o + n]   # forward shadow lives in the flat buffer
                m._gs_key = None
        for m in self.convs:
            F_gs.conv_shadows(m)
        self.image_convs = [m for m in self.convs if F_gs.is_image_conv(m)]

        # gradient exchange plan: which flat ranges become final when (see functional._stage_grads_done), and which
        # ranges belong to blocks a depth-truncated sub-net never runs (all-zero on every rank -> not exchanged at all)
        self._sizes = {o: (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN for p, o in zip(params, offs)}
        self._off_of = off_of
        self._stage_blocks = {}      # stage index (1-based) -> (layer module, [(lo, hi) per block])
        self._stage_extra = {}       # stage index -> [(lo, hi)] of heads whose input feature that stage produces
        self._done = []              # ranges already exchanged (or known to be all-zero) in this iteration
        self._overlap = self.peer_grad is not None and os.environ.get('GS_GRAD_OVERLAP', '1') != '0'
        self._check = os.environ.get('GS_GRAD_OVERLAP_CHECK', '0') == '1'
        self._snapshots = []
        self._plan(model)

    # -- plan -----------------------------------------------------------------------------------
    def _ranges_of(self, module):
        """Merged flat ranges [(lo, hi)] covering the trainable parameters of `module`."""
        offs = sorted(self._off_of[id(p)] for p in module.parameters() if id(p) in self._off_of)
        out = []
        for o in offs:
            hi = o + self._sizes[o]
            if out and out[-1][1] == o:
                out[-1][1] = hi
            else:
                out.append([o, hi])
        return [tuple(r) for r in out]

    def _plan(self, model):
        import weakref
        backbone = getattr(model, 'backbone', None)
        names = getattr(backbone, 'res_layers', None)
        if backbone is None or not names:
            return
        ref = weakref.ref(self)
        for k, name in enumerate(names, start=1):
            layer = getattr(backbone, name)
            blocks = []
            for blk in layer:
                r = self._ranges_of(blk)
                if len(r) != 1:          # a block's parameters are not one contiguous range: leave the stage unplanned
                    blocks = None
                    break
                blocks.append(r[0])
            if not blocks:
                continue
            self._stage_blocks[k] = (weakref.ref(layer), blocks)
            if self._overlap:
                layer[0]._gs_grad_stage = (ref, k)
        out_idx = list(getattr(backbone, 'out_indices', range(len(names))))
        heads = []
        for attr in ('decode_head', 'auxiliary_head'):
            h = getattr(model, attr, None)
            if h is None:
                continue
            heads += list(h) if isinstance(h, nn.ModuleList) else [h]
        for h in heads:
            idx = getattr(h, 'in_index', None)
            if not isinstance(idx, int):
                continue
            try:
                k = out_idx[idx] + 1
            except IndexError:
                continue
            if k in self._stage_blocks:
                self._stage_extra.setdefault(k, []).extend(self._ranges_of(h))

    @staticmethod
    def _merge(ranges):
        out = []
        for lo, hi in sorted(r for r in ranges if r[1] > r[0]):
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return [tuple(r) for r in out]

    def _inactive_ranges(self):
        """Flat ranges of blocks beyond the sampled depth of every planned stage (the reference's DDP
        all-reduces the zeros of the whole max-width model, gaiaseg/apis/train.py:88-95; here they never travel)."""
        out = []
        for k, (lref, blocks) in self._stage_blocks.items():
            layer = lref()
            d = getattr(layer, 'depth_state', len(blocks)) if layer is not None else len(blocks)
            if d < len(blocks):
                out.append((blocks[d][0], blocks[-1][1]))
        return out

    def _pending_ranges(self):
        """What is left to exchange after the backward pass: [0, total) minus the per-stage chunks already done minus
        the blocks outside the sampled depth."""
        out, cur = [], 0
        for lo, hi in self._merge(self._done + self._inactive_ranges()) + [(self.total, self.total)]:
            if lo > cur:
                out.append((cur, lo))
            cur = max(cur, hi)
        return out

    def _reduce_stage(self, k):
        """Backward of res stage k has been enqueued: exchange its active blocks and the heads it feeds on the side
        stream (their gradients are final by data dependency)."""
        if not self._overlap or k not in self._stage_blocks:
            return
        lref, blocks = self._stage_blocks[k]
        layer = lref()
        d = getattr(layer, 'depth_state', len(blocks)) if layer is not None else len(blocks)
        todo = self._merge([(blocks[0][0], blocks[min(d, len(blocks)) - 1][1])] + self._stage_extra.get(k, []))
        pg = self.peer_grad
        for lo, hi in todo:
            F_gs.side_stream_run(lambda lo=lo, hi=hi: pg.all_reduce(lo, hi - lo), self.flat_g.device)
            self._done.append((lo, hi))
            if self._check:
                self._snapshots.append((lo, hi, None))

    def zero_grad(self):
        """Contract of one iteration: zero_grad() -> ONE backward pass -> all_reduce_grads() -> optimizer step."""
        F_gs.reset_side_state()
        self.flat_g.zero_()
        self._done = []
        self._snapshots = []

    def all_reduce_grads(self, group=None, bucket_bytes=64 << 20):
        """Sum the flat gradient over the data-parallel group (the mean is folded into the optimizer's grad_scale).
        Peer-memory path: everything the per-stage chunks have not exchanged yet, minus the blocks outside the sampled
        depth (zero everywhere).  NCCL path (fallback): bucketed all_reduce of the whole buffer."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return 1
        if self.peer_grad is not None and group is None:
            F_gs.wgrad_join()
            if self._check and self._snapshots:
                self._verify_chunks()
            for lo, hi in self._pending_ranges():
                self.peer_grad.all_reduce(lo, hi - lo)
            self._done = []
            return self.peer_grad.world
        n = bucket_bytes // 4
        for s in range(0, self.total, n):
            dist.all_reduce(self.flat_g[s:s + n], group=group)
        return dist.get_world_size(group)

    def _verify_chunks(self):
        """GS_GRAD_OVERLAP_CHECK=1 (debug): a range that was exchanged during the backward pass must be identical on
        every rank afterwards -- a gradient accumulated into it AFTER its exchange would be rank-local and break that."""
        for lo, hi, _ in self._snapshots:
            mine = self.flat_g[lo:hi]
            ref = mine.clone()
            dist.broadcast(ref, 0)
            if not torch.equal(mine, ref):
                raise F_gs.GsError(f'overlapped gradient all-reduce: range [{lo}, {hi}) was modified after its exchange '
                                   f'(a late local gradient); set GS_GRAD_OVERLAP=0 and report the model structure')
        self._snapshots = []


class GsSGD(torch.optim.Optimizer):
    """SGD(momentum, weight_decay) as one fused kernel over the flat master buffer; the same launch rewrites
    the bf16 shadow weights (forward and dgrad read the same buffer).
    Semantics of torch.optim.SGD with dampening 0 / nesterov False
    (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:175).  Every trainable parameter is updated every
    step (zero gradient for channels / blocks outside the sampled sub-net, like the reference after
    `optimizer.zero_grad()` once each block has been visited by the MAX sub-net)."""

    def __init__(self, model_or_flat, lr=0.01, momentum=0.9, weight_decay=0.0, grad_scale=1.0):
        self.flat = model_or_flat if isinstance(model_or_flat, FlatParams) else FlatParams(model_or_flat)
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay, initial_lr=lr)
        super().__init__(self.flat.params, defaults)
        self.momentum_buf = torch.zeros_like(self.flat.flat_p)
        self.grad_scale = grad_scale
        self._steps = 0
        self._hyper_dev = torch.zeros(4, dtype=torch.float32, device=self.flat.flat_p.device)
        self._hyper_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self._hyper_last = None

    def sync_hyper(self):
        """Push {lr, momentum, weight_decay, grad_scale} to the device (only when they changed).  Called by step();
        a graph-replaying caller calls it BEFORE the replay, outside the captured region."""
        g = self.param_groups[0]
        cur = (float(g['lr']), float(g['momentum']), float(g['weight_decay']), float(self.grad_scale))
        if cur != self._hyper_last:
            for i, v in enumerate(cur):
                self._hyper_host[i] = v
            self._hyper_dev.copy_(self._hyper_host, non_blocking=True)
            self._hyper_last = cur

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        f = self.flat
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        call('gs_sgd_flat', f.flat_p.data_ptr(), f.flat_g.data_ptr(), self.momentum_buf.data_ptr(), f.total,
             self._hyper_dev.data_ptr(), 0, f.flat_shadow.data_ptr(), F_gs._stream())   # buf starts at zero: mom * 0 + d == d
        for m in f.image_convs:
            F_gs.refresh_image_shadow(m)
        self._steps += 1

    def _logical(self, flat, p, off):
        """View of `flat` at a parameter's offset with the parameter's logical (OIHW) shape."""
        v = flat[off:off + p.numel()]
        if p.dim() == 4:
            Co, Ci, kh, kw = p.shape
            return v.view(Co, kh, kw, Ci).permute(0, 3, 1, 2)
        return v.view(p.shape)

    def state_dict(self):
        """torch.optim.SGD layout (what an mmcv / reference checkpoint carries): state[i]['momentum_buffer'] with the
        parameter's logical OIHW shape + param_groups with `params` indices, so checkpoints written here resume in the
        reference and vice versa.  (No state before the first step, like torch.)"""
        n = len(self.flat.params)
        state = {}
        if self._steps > 0:
            for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
                state[i] = {'momentum_buffer': self._logical(self.momentum_buf, p, off).detach().clone().contiguous()}
        groups = []
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != 'params'}
            d.update(dampening=0, nesterov=False, params=list(range(n)))
            groups.append(d)
        return dict(state=state, param_groups=groups)

    def load_state_dict(self, sd):
        """Accepts the torch.optim.SGD layout (reference checkpoints) and the round-1 private layout (flat buffer)."""
        if 'momentum_buf' in sd:                      # round-1 layout
            self.momentum_buf.copy_(sd['momentum_buf'])
            self._steps = int(sd.get('steps', 1))
        else:
            state = sd.get('state', {})
            if len(sd['param_groups']) != 1 or len(sd['param_groups'][0]['params']) != len(self.flat.params):
                raise ValueError('optimizer state: expected ONE param group over %d parameters' % len(self.flat.params))
            self.momentum_buf.zero_()                 # parameters without state (never had a gradient) start from zero
            with torch.no_grad():
                for i, (p, off) in enumerate(zip(self.flat.params, self.flat.offsets)):
                    st = state.get(i, state.get(str(i)))
                    if st is None or st.get('momentum_buffer') is None:
                        continue
                    mb = st['momentum_buffer']
                    if tuple(mb.shape) != tuple(p.shape):
                        raise ValueError(f'momentum buffer {i}: shape {tuple(mb.shape)} vs parameter {tuple(p.shape)}')
                    self._logical(self.momentum_buf, p, off).copy_(mb)
            self._steps = 1 if state else 0
        for g, s_ in zip(self.param_groups, sd['param_groups']):
            g.update({k: v for k, v in s_.items() if k not in ('params', 'dampening', 'nesterov')})
        self._hyper_last = None


def reserve_activation_pool(gigabytes=32.0, device=None):
    """Every iteration trains a DIFFERENT sub-net, i.e. a different set of activation sizes.  With a cold caching
    allocator each new size class costs a device-synchronising cudaMalloc in the middle of the step (measured: +24 ms per
    sandwich cycle until the pool has grown).  Allocating one large block up front and releasing it to the cache gives
    the allocator a single segment it can split for every later request (180 GB of HBM: 32 GB is affordable)."""
    if gigabytes <= 0:
        return
    device = device or torch.device('cuda', torch.cuda.current_device())
    free, _ = torch.cuda.mem_get_info(device)
    n = min(int(gigabytes * 2 ** 30), int(free * 0.6))
    block = torch.empty(n, dtype=torch.uint8, device=device)
    del block


class GraphedTrainStep:
    """One whole training iteration -- forward, fused loss, backward, gradient all-reduce, fused SGD -- replayed as
    a CUDA graph for sub-nets that recur (the anchors of the sampler: MAX / MIN of the sandwich rule, the single arch
    of finetune_supernet).  A MAX iteration is ~1500 kernel launches; replaying them costs one graph launch instead
    of ~25 ms of Python.  Sub-nets seen fewer than `graph_after` times run eagerly (random sub-nets never repeat:
    85 293 of them), and at most `max_graphs` graphs are kept (each owns the activation memory of its sub-net).

    The graph bakes in pointers, shapes and the sub-net; everything that varies per iteration lives in device
    memory: the input batch (static buffers), the SGD hyper-parameters (GsSGD._hyper_dev) and dropout's Philox
    offsets (torch's graph-safe generator)."""

    def __init__(self, model, optimizer, graph_after=2, max_graphs=6, pool_gb=32.0):
        self.model, self.opt = model, optimizer
        self.pool_gb = pool_gb
        self.graph_after, self.max_graphs = graph_after, max_graphs
        self.seen, self.graphs = {}, OrderedDict()
        # ONE stepper (= one stream) per optimizer: autograd's gradient accumulators remember the stream they were created
        # on, a second stepper on the same parameters would make a captured backward depend on uncaptured work of the
        # first one's stream (cudaErrorStreamCaptureIsolation)
        prev = getattr(optimizer, '_gs_stepper', None)
        if prev is not None and prev() is not None and prev() is not self and max_graphs > 0:
            raise F_gs.GsError('GraphedTrainStep: this optimizer is already driven by another GraphedTrainStep; use ONE '
                               'stepper per model and key the sub-nets through arch_key')
        import weakref
        optimizer._gs_stepper = weakref.ref(self)
        # every iteration (eager, capture, replay) runs on ONE dedicated side stream: a graph may only be captured on a
        # stream whose tensors carry no dependency on uncaptured work of another stream (autograd would otherwise
        # insert cross-stream waits during the captured backward: cudaErrorStreamCaptureIsolation)
        self.stream = None

    def _module(self):
        return self.model.module if hasattr(self.model, 'module') else self.model

    def _eager(self, batch):
        out = self.model.train_step(batch, self.opt)
        self.opt.zero_grad()
        out['loss'].backward()
        w = self.opt.flat.all_reduce_grads()
        self.opt.grad_scale = 1.0 / w
        self.opt.step()
        return out

    def _graphable(self):
        """With several ranks the SyncBN exchange must be the capturable peer-memory kernel (not NCCL)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return True
        return bool(F_gs.PeerExchange.get(None))

    def __call__(self, arch_key, batch):
        """arch_key: hashable id of the currently applied sub-net (e.g. json.dumps(meta['arch'], sort_keys=True))."""
        first = self.stream is None
        if first:
            self.stream = torch.cuda.Stream(priority=int(os.environ.get("GS_MAIN_STREAM_PRIORITY", "-1")))
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if first:
                reserve_activation_pool(self.pool_gb)     # block pools are per stream: reserve on OUR stream
            out = self._run(arch_key, batch)
        cur.wait_stream(self.stream)
        return out

    def _run(self, arch_key, batch):
        img, gt = batch['img'], batch['gt_semantic_seg']
        key = (arch_key, tuple(img.shape), tuple(gt.shape), self._module().training)
        n = self.seen.get(key, 0) + 1
        self.seen[key] = n
        entry = self.graphs.get(key)
        if entry is None:
            if n <= self.graph_after or not self._graphable():
                return self._eager(batch)
            entry = self._capture(key, batch)
        else:
            self.graphs.move_to_end(key)
        entry['img'].copy_(img, non_blocking=True)
        entry['gt'].copy_(gt, non_blocking=True)
        w = 1
        if dist.is_available() and dist.is_initialized():
            w = dist.get_world_size()
        self.opt.grad_scale = 1.0 / w
        self.opt.sync_hyper()
        entry['graph'].replay()
        if entry['tail_eager']:               # several ranks: NCCL exchanges stay outside the graph
            self.opt.flat.all_reduce_grads()
            self.opt.step()
        else:
            self.opt._steps += 1
        for bn in entry['bns']:
            bn._gs_nbt_pending = getattr(bn, '_gs_nbt_pending', 0) + 1
        out = dict(entry['out'])
        lv = out.get('log_vars')
        if lv is not None and hasattr(lv, 'fresh'):
            if entry['tail_eager']:
                red = lv._stacked / w
                dist.all_reduce(red)
                out['log_vars'] = type(lv)(list(lv.keys()), red)
            else:
                out['log_vars'] = lv.fresh()  # same device tensor, values of THIS replay
        return out

    def _capture(self, key, batch):
        while len(self.graphs) >= self.max_graphs:
            self.graphs.popitem(last=False)
        dev = batch['img'].device
        entry = dict(img=batch['img'].clone(), gt=batch['gt_semantic_seg'].clone())
        static = dict(batch, img=entry['img'], gt_semantic_seg=entry['gt'])
        w = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.opt.grad_scale = 1.0 / w
        self.opt.sync_hyper()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        steps0 = self.opt._steps
        # several ranks without the peer-memory gradient buffer: the graph holds forward + backward (+ the peer-memory
        # SyncBN exchanges) only, the NCCL all-reduce and the optimizer launch stay eager
        tail_eager = w > 1 and self.opt.flat.peer_grad is None
        mod = self._module()
        F_gs._touched_bns = []
        try:
            if tail_eager:
                mod.log_vars_reduce = False
            with torch.cuda.graph(graph, stream=self.stream):
                F_gs._capture_arena = F_gs.CaptureArena(dev)
                F_gs._capture_arena.begin()
                out = self.model.train_step(static, self.opt)
                self.opt.zero_grad()
                out['loss'].backward()
                if not tail_eager:
                    self.opt.flat.all_reduce_grads()
                    self.opt.step()
        finally:
            arena, F_gs._capture_arena = F_gs._capture_arena, None
            bns, F_gs._touched_bns = F_gs._touched_bns, None
            if tail_eager:
                mod.log_vars_reduce = True
        self.opt._steps = steps0            # capturing enqueued nothing; the replay below is the real iteration
        for bn in bns:
            bn._gs_nbt_pending -= 1
        entry.update(graph=graph, out=out, bns=bns, arena=arena, tail_eager=tail_eager)
        reserve_activation_pool(self.pool_gb)   # torch.cuda.graph() empties the allocator cache before capturing
        self.graphs[key] = entry
        return entry


def build_optimizer(model, cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type', 'SGD')
    if typ != 'SGD':
        raise NotImplementedError(f'optimizer {typ}: the GAIA-seg recipe is SGD')
    if cfg.pop('nesterov', False) or cfg.pop('dampening', 0):
        raise NotImplementedError('nesterov / dampening are not used by the GAIA-seg recipe')
    module = model.module if hasattr(model, 'module') else model
    return GsSGD(module, **cfg)


class GsDataParallel(nn.Module):
    """Stand-in for MMDistributedDataParallel(model.cuda(), broadcast_buffers=False,
    find_unused_parameters=True) (gaiaseg/apis/train.py:88-95): one process per GPU, `.module` access,
    `train_step` / `val_step` forwarding.  Gradient exchange is `FlatParams.all_reduce_grads` called by the
    optimizer hook after backward."""

    def __init__(self, module, device_ids=None, broadcast_buffers=False, find_unused_parameters=True, **kw):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def train_step(self, *args, **kwargs):
        return self.module.train_step(*args, **kwargs)

    def val_step(self, *args, **kwargs):
        return self.module.val_step(*args, **kwargs)


# ------------------------------------------------------------------------------------------------
# checkpoint IO
# ------------------------------------------------------------------------------------------------
# the stem norm's attribute name is whatever gaiavision's build_norm_layer returned -- unverifiable, so
# accept the common spellings when loading
_KEY_REMAP = (('backbone.norm1.', 'backbone.bn1.'), ('norm1.', 'bn1.'))


def _unwrap(model):
    return model.module if hasattr(model, 'module') and isinstance(model.module, nn.Module) else model


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def save_checkpoint(model, filename, optimizer=None, meta=None):
    model = _unwrap(model)
    sd = OrderedDict((k, v.detach().cpu().contiguous()) for k, v in model.state_dict().items())
    ckpt = dict(meta=dict(meta or {}, time=time.asctime()), state_dict=sd)
    if optimizer is not None:
        ckpt['optimizer'] = _to_cpu(optimizer.state_dict())
    os.makedirs(osp.dirname(osp.abspath(filename)), exist_ok=True)
    torch.save(ckpt, filename)


def load_checkpoint(model, filename, map_location='cpu', strict=False, logger=None):
    ckpt = torch.load(filename, map_location=map_location, weights_only=False)
    sd = ckpt.get('state_dict', ckpt) if isinstance(ckpt, dict) else ckpt
    sd = OrderedDict((k[7:] if k.startswith('module.') else k, v) for k, v in sd.items())
    model = _unwrap(model)
    own = model.state_dict()
    fixed = OrderedDict()
    for k, v in sd.items():
        if k not in own:
            for a, b in _KEY_REMAP:
                if k.startswith(a) and (b + k[len(a):]) in own:
                    k = b + k[len(a):]
                    break
        fixed[k] = v
    with torch.no_grad():
        missing, unexpected = [], [k for k in fixed if k not in own]
        for k, dst in own.items():
            if k not in fixed:
                missing.append(k)
                continue
            src = fixed[k]
            if tuple(src.shape) != tuple(dst.shape):
                raise RuntimeError(f'size mismatch for {k}: checkpoint {tuple(src.shape)} vs model {tuple(dst.shape)}')
            dst.copy_(src)   # in place: keeps the flat-buffer views and bumps _version (shadows refresh)
    if strict and (missing or unexpected):
        raise RuntimeError(f'missing keys {missing}, unexpected keys {unexpected}')
    return ckpt


# ------------------------------------------------------------------------------------------------
# hooks + runner
# ------------------------------------------------------------------------------------------------
class Hook:
    def before_run(self, runner): pass
    def after_run(self, runner): pass
    def before_train_iter(self, runner): pass
    def after_train_iter(self, runner): pass

    @staticmethod
    def every_n_iters(runner, n):
        return (runner.iter + 1) % n == 0 if n > 0 else False


class PolyLrUpdaterHook(Hook):
    """lr = (base - min_lr) * (1 - iter/max_iters)^power + min_lr   (mmcv PolyLrUpdaterHook, by_epoch=False)."""

    def __init__(self, power=1.0, min_lr=0.0, by_epoch=False, **kw):
        self.power, self.min_lr = power, min_lr

    def before_run(self, runner):
        self.base_lr = [g.setdefault('initial_lr', g['lr']) for g in runner.optimizer.param_groups]

    def before_train_iter(self, runner):
        coeff = (1 - runner.iter / runner.max_iters) ** self.power
        for g, base in zip(runner.optimizer.param_groups, self.base_lr):
            g['lr'] = (base - self.min_lr) * coeff + self.min_lr


class OptimizerHook(Hook):
    """zero_grad -> backward -> (gradient all-reduce) -> step  (mmcv OptimizerHook.after_train_iter)."""

    def __init__(self, grad_clip=None, **kw):
        if grad_clip is not None:
            raise NotImplementedError('grad_clip is not used by the GAIA-seg recipe')

    def after_train_iter(self, runner):
        if getattr(runner, '_step_done', False):     # GraphedTrainStep already ran backward + all-reduce + step
            return
        opt = runner.optimizer
        opt.zero_grad()
        runner.outputs['loss'].backward()
        if isinstance(opt, GsSGD):
            world = opt.flat.all_reduce_grads()
            opt.grad_scale = 1.0 / world
        opt.step()


class CheckpointHook(Hook):
    def __init__(self, interval=-1, by_epoch=False, out_dir=None, **kw):
        self.interval, self.out_dir = interval, out_dir

    def after_train_iter(self, runner):
        if self.every_n_iters(runner, self.interval) and runner.rank == 0:
            out = self.out_dir or runner.work_dir
            path = osp.join(out, f'iter_{runner.iter + 1}.pth')
            save_checkpoint(runner.model, path, optimizer=runner.optimizer, meta=dict(runner.meta or {}, iter=runner.iter + 1))
            latest = osp.join(out, 'latest.pth')
            if osp.lexists(latest):
                os.remove(latest)
            os.symlink(osp.basename(path), latest)
        if self.every_n_iters(runner, self.interval) and dist.is_available() and dist.is_initialized():
            # rank 0 alone wrote the file: re-align the ranks on the host BEFORE they enqueue the next iteration -- the
            # peer-memory exchange kernels spin on the device with a bounded budget (GS_COMM_TIMEOUT_S)
            dist.barrier()


class TextLoggerHook(Hook):
    def __init__(self, interval=50, by_epoch=False, **kw):
        self.interval = interval
        self._t = None

    def before_run(self, runner):
        self._t = time.time()

    def after_train_iter(self, runner):
        if self.every_n_iters(runner, self.interval):
            lv = runner.outputs['log_vars']
            now = time.time()
            dt = (now - self._t) / self.interval
            self._t = now
            if runner.rank == 0:
                msg = ', '.join(f'{k}: {v:.4f}' for k, v in lv.items())
                lr = runner.optimizer.param_groups[0]['lr']
                print(f'Iter [{runner.iter + 1}/{runner.max_iters}] lr: {lr:.3e}, time: {dt:.3f}, {msg}', flush=True)


class IterBasedRunner:
    """mmcv IterBasedRunner for the single ('train', 1) workflow of the reference config.  `graph_replay` (default on,
    `runner=dict(type='IterBasedRunner', graph_replay=False)` to disable): iterations run through GraphedTrainStep, so
    sub-nets that recur (the anchors of the sampler, the single arch of a finetune) are replayed as CUDA graphs."""

    def __init__(self, model, optimizer=None, work_dir=None, logger=None, meta=None, max_iters=None, graph_replay=True,
                 **kw):
        self.model, self.optimizer, self.work_dir, self.logger, self.meta = model, optimizer, work_dir, logger, meta
        self.max_iters = max_iters
        self.graph_replay, self._graphed, self._step_done = graph_replay, None, False
        self.iter = 0
        self.hooks = []
        self.outputs = None
        self.rank, self.world_size = get_dist_info()
        self.timestamp = None

    def register_hook(self, hook, priority='NORMAL'):
        self.hooks.append(hook)

    def register_training_hooks(self, lr_config, optimizer_config=None, checkpoint_config=None, log_config=None,
                                momentum_config=None):
        if lr_config is not None:
            cfg = dict(lr_config)
            policy = cfg.pop('policy', 'poly')
            if policy != 'poly':
                raise NotImplementedError(f'lr policy {policy}: the GAIA-seg recipe is poly')
            self.register_hook(PolyLrUpdaterHook(**cfg))
        self.register_hook(OptimizerHook(**dict(optimizer_config or {})))
        if checkpoint_config is not None:
            self.register_hook(CheckpointHook(**dict(checkpoint_config)))
        if log_config is not None:
            for h in log_config.get('hooks', []):
                if h.get('type') == 'TextLoggerHook':
                    self.register_hook(TextLoggerHook(interval=log_config.get('interval', 50)))

    def call_hook(self, name):
        # LR and arch hooks run before the iteration in registration order; the optimizer hook first after it
        for h in self.hooks:
            getattr(h, name, lambda r: None)(self)

    def _arch_key(self):
        """Identity of the sub-net the ManipulateArchHook applied for this iteration (graph cache key)."""
        import json
        for h in self.hooks:
            meta = getattr(h, 'last_meta', None)
            if meta is not None:
                return json.dumps(meta.get('arch', meta), sort_keys=True, default=str)
        return 'fixed-arch'

    def load_checkpoint(self, filename, map_location='cpu', strict=False):
        return load_checkpoint(self.model, filename, map_location, strict)

    def resume(self, checkpoint):
        ckpt = self.load_checkpoint(checkpoint)
        self.iter = ckpt.get('meta', {}).get('iter', 0)
        if 'optimizer' in ckpt and self.optimizer is not None:
            self.optimizer.load_state_dict(ckpt['optimizer'])

    def run(self, data_loaders, workflow=None, max_iters=None, **kw):
        if max_iters is not None:
            self.max_iters = max_iters
        loader = data_loaders[0] if isinstance(data_loaders, (list, tuple)) else data_loaders
        self.model.train()
        self.call_hook('before_run')
        epoch = 0

        def new_iter():
            # mmcv IterLoader: every pass over the dataset reshuffles (DistributedSampler.set_epoch)
            sampler = getattr(loader, 'sampler', None)
            if hasattr(sampler, 'set_epoch'):
                sampler.set_epoch(epoch)
            return iter(loader)

        it = new_iter()
        while self.iter < self.max_iters:
            try:
                data_batch = next(it)
            except StopIteration:
                epoch += 1
                it = new_iter()
                data_batch = next(it)
            self.call_hook('before_train_iter')
            data_batch = scatter_batch(data_batch)
            if self.graph_replay and isinstance(self.optimizer, GsSGD):
                if self._graphed is None:
                    self._graphed = GraphedTrainStep(self.model, self.optimizer)
                self.outputs = self._graphed(self._arch_key(), data_batch)
                self._step_done = True
            else:
                self.outputs = self.model.train_step(data_batch, self.optimizer)
                self._step_done = False
            self.call_hook('after_train_iter')
            self.iter += 1
        self.call_hook('after_run')


def scatter_batch(batch, device=None):
    """host -> device copy of one batch (MMDDP.scatter): img fp32, gt_semantic_seg int64, metas untouched."""
    device = device or torch.device('cuda', torch.cuda.current_device())
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=True)
        elif isinstance(v, list) and v and torch.is_tensor(v[0]):
            out[k] = [t.to(device, non_blocking=True) for t in v]
        else:
            out[k] = v
    return out


def build_runner(cfg, default_args=None):
    cfg = dict(cfg)
    typ = cfg.pop('type', 'IterBasedRunner')
    if typ != 'IterBasedRunner':
        raise NotImplementedError(f'runner {typ}: the GAIA-seg recipe uses IterBasedRunner')
    args = dict(default_args or {})
    args.update(cfg)
    args.pop('batch_processor', None)
    return IterBasedRunner(**args)
