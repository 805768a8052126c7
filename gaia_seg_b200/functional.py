"""Host-side launch layer: tensors in, C-ABI calls out.

Every function here enqueues hand-written sm_100a kernels from libgaiaseg_b200.so on torch's current
CUDA stream.  PyTorch is only used for device memory (caching allocator), streams and autograd
bookkeeping.  Activations are bf16 "NHWC with pitch": a torch tensor of logical shape [N, C, H, W]
whose memory is [N][H][W][ld] (ld >= C), i.e. channels_last or a channel slice of a wider buffer.

Reference semantics being replaced (all [EXT] gaiavision / mmseg):
  DynamicConv2d.forward      F.conv2d(x, W[:width_state, :x.size(1)], b[:width_state], s, p, d)
  DynamicBatchNorm2d.forward F.batch_norm(x, rm[:C], rv[:C], w[:C], b[:C], training, 0.1, 1e-5)
  DynamicBottleneck.forward  relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1 x))))))) + identity)
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ConvGeom, GsError, call

BF16 = torch.bfloat16

# bench.py sets this to a list to time every convolution launch with CUDA events on the launching stream:
# entries are (kind, algorithmic_flops, start_event, end_event)
PROFILE = None


def _timed_call(kind, g, fn, *args):
    if PROFILE is None:
        call(fn, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(fn, *args)
    e1.record()
    flops = 2.0 * g.N * g.Ho * g.Wo * g.Co * g.Ci * g.kh * g.kw
    PROFILE.append((kind, flops, e0, e1, (g.N * g.Ho * g.Wo, g.Ci, g.Co, g.kh, g.stride, g.dil)))


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (raw handle: torch.cuda.current_stream()
    costs ~15 us of Python per call, this ~0.3 us)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def round_up(a, b):
    return (a + b - 1) // b * b


# ------------------------------------------------------------------------------------------------
# activation tensors
# ------------------------------------------------------------------------------------------------
def new_act(N, C, H, W, device, dtype=BF16, ld=None):
    """Uninitialised [N, C, H, W] activation with NHWC memory and pixel pitch ld (default C)."""
    ld = C if ld is None else ld
    t = torch.empty((N, H, W, ld), dtype=dtype, device=device).permute(0, 3, 1, 2)
    return t if ld == C else t[:, :C]


def act_ld(t):
    """Pixel pitch (elements) of an NHWC-with-pitch tensor, or None when the layout is anything else."""
    N, C, H, W = t.shape
    s = t.stride()
    if C > 1 and s[1] != 1:
        return None
    if W > 1:
        ld = s[3]
    elif H > 1:
        ld = s[2]
    elif N > 1:
        ld = s[0]
    else:
        ld = C
    if ld < C:
        return None
    if (W > 1 and s[3] != ld) or (H > 1 and s[2] != W * ld) or (N > 1 and s[0] != H * W * ld):
        return None
    return ld


def as_act(t, dtype=BF16):
    """Return `t` as an NHWC-with-pitch tensor of `dtype` the kernels can address (16-byte aligned
    vectors); converts only when the layout does not already qualify."""
    if not t.is_cuda:
        raise GsError('gaia_seg_b200: activation is not on a CUDA device -- the hot path has no CPU fallback')
    if t.dim() != 4:
        raise GsError(f'expected a 4-D activation, got shape {tuple(t.shape)}')
    if t.dtype != dtype:
        t = t.to(dtype)
    ld = act_ld(t)
    esz = t.element_size()
    if ld is None or (ld * esz) % 16 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous(memory_format=torch.channels_last)
        if act_ld(t) is None:  # degenerate shapes for which torch keeps NCHW strides
            N, C, H, W = t.shape
            o = new_act(N, C, H, W, t.device, dtype)
            o.copy_(t)
            t = o
    return t


def _pixels(t):
    return t.shape[0] * t.shape[2] * t.shape[3]


class _ZeroPool:
    """Zero-initialised fp64 scratch for the per-layer statistic accumulators (the kernels add into them with
    atomics).  One big buffer per device, handed out in slices and re-zeroed with ONE memset when it wraps, instead
    of a torch.zeros launch per layer.  Safe because every slice is consumed in stream order right after it is
    produced (stats -> apply, sums -> bwd_apply) and the memset is enqueued on the same stream."""
    SIZE = 1 << 22   # doubles (32 MB)

    def __init__(self):
        self.bufs = {}

    def take(self, n, device):
        n = (n + 15) // 16 * 16
        st = self.bufs.get(device)
        if st is None:
            st = self.bufs[device] = [torch.zeros(self.SIZE, dtype=torch.float64, device=device), 0]
        if n > self.SIZE:
            return torch.zeros(n, dtype=torch.float64, device=device)
        if st[1] + n > self.SIZE:
            st[0].zero_()
            st[1] = 0
        out = st[0][st[1]:st[1] + n]
        st[1] += n
        return out


_zero_pool = _ZeroPool()


class CaptureArena:
    """Statistic accumulators of ONE captured CUDA graph: a single buffer inside the graph's memory pool, zeroed by
    one memset node at the start of the graph (a replay must start from zeros, the wrap-around policy of the eager
    pool cannot guarantee that)."""

    def __init__(self, device, size=1 << 20):
        self.buf = torch.zeros(size, dtype=torch.float64, device=device)
        self.off = 0

    def begin(self):
        self.buf.zero_()
        self.off = 0

    def take(self, n):
        n = (n + 15) // 16 * 16
        if self.off + n > self.buf.numel():
            raise GsError('CaptureArena exhausted')
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out


_capture_arena = None   # set by runner.GraphedTrainStep while a graph is being captured


def zeros_f64(n, device):
    if _capture_arena is not None:
        return _capture_arena.take(n)[:n]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not PeerExchange.get():
        return torch.zeros(n, dtype=torch.float64, device=device)   # NCCL-reduced buffers: keep them private
    return _zero_pool.take(n, device)[:n]


_touched_bns = None     # list collecting the BN modules whose running stats a captured graph updates


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
class ConvCall:
    """Geometry of one DynamicConv2d invocation (mirrors struct gs_conv_geom)."""
    __slots__ = ('geom', 'image', 'Kpad', 'N', 'Ho', 'Wo', 'Co', 'Ci')


def _conv_attrs(conv):
    kh, kw = conv.kernel_size
    if conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] or conv.dilation[0] != conv.dilation[1]:
        raise GsError('DynamicConv2d: only square stride / padding / dilation are supported')
    if conv.groups != 1:
        raise GsError('DynamicConv2d: groups != 1 is not on the GAIA-seg hot path')
    return kh, kw, conv.stride[0], conv.padding[0], conv.dilation[0]


def is_image_conv(conv):
    return conv.in_channels < 8


def image_kpad(conv):
    kh, kw = conv.kernel_size
    return round_up(kh * kw * conv.in_channels, 16)


def ensure_krsc(conv):
    """The fp32 master weight must be stored [Co][kh][kw][Ci] (channels_last memory of OIHW)."""
    w = conv.weight
    if not w.permute(0, 2, 3, 1).is_contiguous():
        w.data = w.data.contiguous(memory_format=torch.channels_last)
        if not w.permute(0, 2, 3, 1).is_contiguous():
            Co, Ci, kh, kw = w.shape
            buf = torch.empty((Co, kh, kw, Ci), dtype=w.dtype, device=w.device)
            buf.copy_(w.data.permute(0, 2, 3, 1))
            w.data = buf.permute(0, 3, 1, 2)
    return w


def conv_shadows(conv):
    """bf16 shadow of the max-width weight, [Co_max][kh][kw][Ci_max] (the forward B operand; dgrad reads the same
    buffer as an MN-major operand).  Rebuilt when the fp32 master was modified by anything other than the fused
    optimizer (tracked through tensor._version)."""
    w = ensure_krsc(conv)
    if w.dtype != torch.float32 or not w.is_cuda:
        raise GsError('DynamicConv2d: master weight must be fp32 on a CUDA device')
    key = (w.data_ptr(), w._version)
    if getattr(conv, '_gs_key', None) == key:
        return conv._gs_w_krsc
    Co, Ci, kh, kw = w.shape
    R = kh * kw
    st = _stream()
    krsc = getattr(conv, '_gs_w_krsc', None)
    if is_image_conv(conv):
        K, Kpad = R * Ci, image_kpad(conv)
        if krsc is None or krsc.numel() != Co * Kpad or krsc.device != w.device:
            krsc = torch.empty(Co * Kpad, dtype=BF16, device=w.device)
        call('gs_cast_f32_bf16', w.data_ptr(), K, krsc.data_ptr(), Kpad, Co, K, st)
    else:
        n = w.numel()
        if krsc is None or krsc.numel() != n or krsc.device != w.device:
            krsc = torch.empty(n, dtype=BF16, device=w.device)
        call('gs_cast_f32_bf16', w.data_ptr(), n, krsc.data_ptr(), n, 1, n, st)
    conv._gs_w_krsc, conv._gs_key = krsc, key
    return krsc


def refresh_image_shadow(conv):
    """Called by the fused optimizer for the (zero-padded, im2col-ordered) shadow of the first conv; every other
    shadow lives inside the flat bf16 buffer the optimizer kernel rewrites itself."""
    w = conv.weight
    Co, Ci, kh, kw = w.shape
    K, Kpad = kh * kw * Ci, image_kpad(conv)
    call('gs_cast_f32_bf16', w.data_ptr(), K, conv._gs_w_krsc.data_ptr(), Kpad, Co, K, _stream())


def _geom(N, H, W, Ci, Co, Ci_max, Co_max, kh, kw, stride, pad, dil, x_ld, y_ld):
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    if Ho <= 0 or Wo <= 0:
        raise GsError(f'conv: empty output for input {H}x{W}')
    return ConvGeom(N, H, W, Ho, Wo, Ci, Co, Ci_max, Co_max, kh, kw, stride, pad, dil, x_ld, y_ld)


# GS_CONV_BN_FUSE=1: training forward of conv -> DynBN (-> + residual) (-> ReLU) on a single rank is ONE launch
# (gs_conv2d_fwd_bn: the persistent conv grid meets at a barrier after its statistic flush and normalises the tiles it has
# just written); 0 (default): conv kernel + gs_bn_apply_train.  Parity-tested (tests/test_gpu_path.py::
# test_conv_bn_one_launch_equals_two_kernels) and MEASURED SLOWER (profiles/r02_experiments.md: 53.2 vs 50.2 ms per cycle): with
# one 576-thread CTA per SM and the shared memory taken by the operand ring the in-kernel apply pass streams at ~2.3-4 TB/s,
# and flush -> barrier -> finalize cost ~6 us -- more than the launch they replace.  Several ranks always take the two-kernel
# path (the statistic exchange sits between the two).
CONV_BN_FUSE = os.environ.get('GS_CONV_BN_FUSE', '0') != '0'


def conv_forward(x, conv, Co, scale=None, shift=None, residual=None, relu=False, out_f32=False, want_stats=False, sync=None,
                 fuse_bn=None):
    """y = epi(conv(x, W[:Co, :Ci])).  Returns (y, stats, a_operand, geom): `a_operand` is the tensor the
    weight gradient must be taken against (x itself, or the im2col matrix of the image conv).
    fuse_bn = dict(bn=, residual=, relu=): the fused conv + DynBN apply launch; the dict receives z / aff / count."""
    _lib.require_device()
    krsc = conv_shadows(conv)
    kh, kw, stride, pad, dil = _conv_attrs(conv)
    Co_max, Ci_max = conv.out_channels, conv.in_channels
    if not (0 < Co <= Co_max):
        raise GsError(f'DynamicConv2d: active width {Co} outside (0, {Co_max}]')
    dev = x.device
    st = _stream()
    if is_image_conv(conv):
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        N, Ci, H, W = x.shape
        if Ci != Ci_max:
            raise GsError(f'image conv expects {Ci_max} channels, got {Ci}')
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        if dil != 1:
            raise GsError('image conv: dilation is not supported')
        Kpad = image_kpad(conv)
        cols = new_act(N, Kpad, Ho, Wo, dev)
        call('gs_im2col_image', x.data_ptr(), N, Ci, H, W, kh, kw, stride, pad, Ho, Wo, Kpad, cols.data_ptr(), st)
        a = cols
        g = _geom(N, Ho, Wo, Kpad, Co, Kpad, Co_max, 1, 1, 1, 0, 1, Kpad, 0)
    else:
        a = as_act(x)
        N, Ci, H, W = a.shape
        if Ci > Ci_max:
            raise GsError(f'DynamicConv2d: input has {Ci} channels, max-width weight only {Ci_max}')
        g = _geom(N, H, W, Ci, Co, Ci_max, Co_max, kh, kw, stride, pad, dil, act_ld(a), 0)
    conv._gs_last_ci = Ci
    y_ld = Co if not out_f32 else Co
    y = new_act(N, Co, g.Ho, g.Wo, dev, torch.float32 if out_f32 else BF16, ld=y_ld)
    g.y_ld = y_ld
    stats = None
    if want_stats:
        stats = zeros_f64(2 * Co + 2, dev)[:2 * Co]   # + two zeroed scratch words behind the sums (gs_sync_desc contract)
    flags = (1 if relu else 0) | (2 if out_f32 else 0)
    res_ld = 0
    if residual is not None:
        residual = as_act(residual)
        res_ld = act_ld(residual)
    if fuse_bn is not None:
        bn, bres = fuse_bn['bn'], fuse_bn['residual']
        count = float(N * g.Ho * g.Wo)
        aff = torch.empty((4, Co), dtype=torch.float32, device=dev)
        upd = bn.training and bn.track_running_stats and bn.running_mean is not None
        if upd and bn.momentum is None:
            raise GsError('DynamicBatchNorm2d: momentum=None (cumulative average) is not supported on the CUDA path')
        z = new_act(N, Co, g.Ho, g.Wo, dev)
        bres_ld = 0
        if bres is not None:
            bres = as_act(bres)
            bres_ld = act_ld(bres)
        _timed_call('fwd', g, 'gs_conv2d_fwd_bn', ctypes.byref(g), a.data_ptr(), krsc.data_ptr(), y.data_ptr(), _ptr(shift),
                    stats.data_ptr(), count, _ptr(bn.weight), _ptr(bn.bias),
                    bn.running_mean.data_ptr() if upd else None, bn.running_var.data_ptr() if upd else None,
                    float(bn.momentum if bn.momentum is not None else 0.0), float(bn.eps), aff.data_ptr(), _ptr(bres), bres_ld,
                    1 if fuse_bn['relu'] else 0, z.data_ptr(), act_ld(z), st)
        if upd:
            bn._gs_nbt_pending = getattr(bn, '_gs_nbt_pending', 0) + 1
            if _touched_bns is not None:
                _touched_bns.append(bn)
        fuse_bn['z'], fuse_bn['aff'], fuse_bn['count'] = z, aff, count
    elif sync is not None:     # several ranks: the conv kernel's last CTA pushes the statistics to the SyncBN peers
        _timed_call('fwd', g, 'gs_conv2d_fwd_syncbn', ctypes.byref(g), a.data_ptr(), krsc.data_ptr(), y.data_ptr(), _ptr(scale),
                    _ptr(shift), _ptr(residual), res_ld, flags, _ptr(stats), sync, st)
    else:
        _timed_call('fwd', g, 'gs_conv2d_fwd', ctypes.byref(g), a.data_ptr(), krsc.data_ptr(), y.data_ptr(), _ptr(scale), _ptr(shift),
                    _ptr(residual), res_ld, flags, _ptr(stats), st)
    return y, stats, a, g


def _weight_grad(conv):
    w = conv.weight
    if w.grad is None:
        w.grad = torch.zeros_like(w, memory_format=torch.preserve_format)
        if not w.grad.permute(0, 2, 3, 1).is_contiguous():
            w.grad = torch.zeros_like(w).contiguous(memory_format=torch.channels_last)
    return w.grad


def conv_wgrad(conv, a, dy, g):
    """conv.weight.grad[:Co, :Ci] += dy^T * im2col(a)  (in place, KRSC fp32)."""
    if not conv.weight.requires_grad:
        return
    st = _stream()
    gw = _weight_grad(conv)
    g.y_ld = act_ld(dy)
    fn = 'gs_conv2d_wgrad'
    if is_image_conv(conv):
        Co_max = conv.out_channels
        K = conv.kernel_size[0] * conv.kernel_size[1] * conv.in_channels
        Kpad = image_kpad(conv)
        tmp = torch.zeros((Co_max, Kpad), dtype=torch.float32, device=dy.device)
        _timed_call('wgrad', g, fn, ctypes.byref(g), a.data_ptr(), dy.data_ptr(), tmp.data_ptr(), st)
        gw.permute(0, 2, 3, 1).reshape(Co_max, K).add_(tmp[:, :K])
    else:
        _timed_call('wgrad', g, fn, ctypes.byref(g), a.data_ptr(), dy.data_ptr(), gw.data_ptr(), st)


def conv_dgrad(conv, dy, g, x_shape, add=None):
    """dx = conv_transpose(dy, W[:Co, :Ci]) (+ add): the residual gradient `add` is summed in the dgrad epilogue."""
    krsc = conv_shadows(conv)
    N, Ci, H, W = x_shape
    dx = new_act(N, Ci, H, W, dy.device)
    g.x_ld = Ci
    g.y_ld = act_ld(dy)
    add_ld = 0
    if add is not None:
        add = as_act(add)
        add_ld = act_ld(add)
    st = _stream()
    ws = None
    nbytes = _lib.load().gs_conv2d_dgrad_workspace_bytes(ctypes.byref(g))
    if nbytes > 0:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dy.device)
    _timed_call('dgrad', g, 'gs_conv2d_dgrad', ctypes.byref(g), dy.data_ptr(), krsc.data_ptr(), dx.data_ptr(),
                _ptr(add), add_ld, _ptr(ws), None, st)
    return dx


# ------------------------------------------------------------------------------------------------
# batch norm pieces
# ------------------------------------------------------------------------------------------------
# GS_BN_FUSED_BWD=1: one cooperative kernel for the BN backward (single rank; measured SLOWER than reduce + apply, kept as
# an experiment); GS_SYNCBN_FOLD=1: block 0 of the forward apply kernel runs the SyncBN exchange itself instead of a separate
# gs_syncbn_allreduce launch (parity-green at N = 2, no measurable gain there: 59.4 vs 59.1 ms per cycle)
# BN backward in ONE launch (gs_bn_bwd: channel-partitioned clusters, csrc/gs_norm.cu) instead of reduce + apply:
#   '0' (default) never | '1' always | 'auto' when one activation is at most GS_BN_FUSED_MAX_MB.
# Parity-tested (tests/test_gpu_path.py::test_bn_backward_one_pass_and_two_kernels) but MEASURED SLOWER inside the training
# step (profiles/r02_experiments.md): the 16-CTA clusters are placed as a unit and hold up the side-stream weight gradients,
# which costs more than the saved launch.  Kept as an option; with several ranks each cluster exchanges its own channels.
FUSED_BN_BWD = os.environ.get('GS_BN_FUSED_BWD', '0')
FUSED_BN_MAX_BYTES = float(os.environ.get('GS_BN_FUSED_MAX_MB', '24')) * 1e6
# SyncBN statistic exchange over NVLink peer memory (several ranks), GS_SYNCBN_FOLD:
#   0  one small exchange kernel per BN layer and direction between producer and consumer;
#   1  folded into the consumer: block 0 of gs_bn_apply_train pushes + polls, the other blocks wait on a flag;
#   2  SPLIT: the producer kernel (conv forward / BN-backward reduction) pushes the final local sums from its last block --
#      compute and the send half of the collective in one kernel -- and the consumer (BN apply / BN-backward apply) polls:
#      no exchange launch, and the NVLink flight overlaps the producer's tail and the consumer's launch.
# Default 0 -- MEASURED (profiles/r02_experiments.md, N = 2): 0 -> 59.0 ms, 1 -> 58.6-60 ms, 2 -> 64.6 ms per cycle: the
# fences + ticket at the tail of every producer and the flag wait of every consumer block cost more than the launch saved.
FOLD_EXCHANGE = int(os.environ.get('GS_SYNCBN_FOLD', '0'))


def bn_batch_mode(bn):
    """True when the layer normalises with mini-batch statistics (train mode, or running stats dropped
    by `caliberate_bn.use_minibatch_stats`, tools/test_supernet.py:190-198)."""
    return bn.training or not bn.track_running_stats or bn.running_mean is None


def _all_agree(ok, group=None):
    """Collective AND over the group: the peer-memory path is taken only when EVERY rank could set it up -- a rank that
    fell back to NCCL alone while its peers spin in the peer kernels would hang the job."""
    on_gpu = torch.cuda.is_available() and str(dist.get_backend(group)).lower() != 'gloo'
    t = torch.tensor([1 if ok else 0], dtype=torch.int32,
                     device=torch.device('cuda', torch.cuda.current_device()) if on_gpu else torch.device('cpu'))
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()))


def _ipc_alloc(nbytes):
    lib = _lib.load()
    ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    _lib.check(lib.gs_ipc_alloc(nbytes, ctypes.byref(ptr), handle), 'gs_ipc_alloc')
    return ptr.value, bytes(handle.raw)


def _ipc_open(handle):
    p = ctypes.c_void_p()
    _lib.check(_lib.load().gs_ipc_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(p)), 'gs_ipc_open')
    return p.value


def _peer_map(sizes, group, what):
    """Allocate one IPC-shareable buffer per entry of `sizes` on every rank and map every peer's buffers.  Returns
    (own_ptrs, tables) with tables[i] = (c_void_p * 8) of buffer i on every rank, or None when ANY rank failed (agreed
    collectively; everything that was allocated / opened is released again)."""
    import warnings
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    own, handles, err = [], [], None
    try:
        for nbytes in sizes:
            ptr, h = _ipc_alloc(nbytes)
            own.append(ptr)
            handles.append(h)
    except Exception as e:   # noqa: BLE001  (no P2P / IPC on this system)
        err = e
    if not _all_agree(err is None, group):
        for ptr in own:
            lib.gs_ipc_free(ptr)
        warnings.warn(f'gaia_seg_b200: {what}: IPC allocation failed on at least one rank ({err}); ALL ranks use NCCL')
        return None
    gathered = [None] * world
    dist.all_gather_object(gathered, handles, group=group)
    tables = [(ctypes.c_void_p * 8)() for _ in sizes]
    opened = []
    try:
        for r, hs in enumerate(gathered):
            for i, h in enumerate(hs):
                if r == rank:
                    tables[i][r] = own[i]
                else:
                    tables[i][r] = _ipc_open(h)
                    opened.append(tables[i][r])
    except Exception as e:   # noqa: BLE001
        err = e
    if not _all_agree(err is None, group):
        for ptr in opened:
            lib.gs_ipc_close(ptr)
        dist.barrier(group=group)          # nobody frees while a peer still has the buffer mapped
        for ptr in own:
            lib.gs_ipc_free(ptr)
        warnings.warn(f'gaia_seg_b200: {what}: peer mapping failed on at least one rank ({err}); ALL ranks use NCCL')
        return None
    dist.barrier(group=group)
    return own, tables


class PeerExchange:
    """NVLink peer-memory all-reduce of the packed SyncBN sums (gs_syncbn_allreduce): one small kernel per layer and
    direction instead of a host-launched NCCL collective.  Set up once per process group: every rank allocates an
    IPC-shareable inbox, the 64-byte handles travel through torch.distributed, peers map each other's inbox.  Whether
    the peer path or the NCCL fallback is used is agreed COLLECTIVELY (all ranks or none)."""
    _instances = {}

    def __init__(self, group, own, table):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.own, self.ptrs = own, table
        self.seq = torch.zeros(1, dtype=torch.int64, device=torch.device('cuda', torch.cuda.current_device()))
        # descriptor handed to the DynBN kernels that run the exchange themselves (gs_bn_apply_train / gs_bn_bwd)
        mk = lambda phase: _lib.SyncDesc(ctypes.cast(self.ptrs, ctypes.c_void_p), self.rank, self.world, self.seq.data_ptr(), phase, 0)
        self.desc, self.desc_push, self.desc_poll = mk(0), mk(1), mk(2)     # whole exchange | producer half | consumer half
        self.desc_ref, self.push_ref, self.poll_ref = (ctypes.byref(d) for d in (self.desc, self.desc_push, self.desc_poll))

    def all_reduce(self, stats, dgamma=None, dbeta=None):
        call('gs_syncbn_allreduce', stats.data_ptr(), stats.numel(), self.ptrs, self.rank, self.world, self.seq.data_ptr(),
             dgamma, dbeta, _stream())

    @classmethod
    def get(cls, group=None):
        key = id(group) if group is not None else 0
        inst = cls._instances.get(key)
        if inst is None:
            world = dist.get_world_size(group)
            want = os.environ.get('GS_SYNCBN_PEER', '1') != '0' and world <= 8 and torch.cuda.is_available()
            if not _all_agree(want, group):
                inst = False
            else:
                m = _peer_map([_lib.load().gs_comm_inbox_bytes(world)], group, 'SyncBN peer exchange')
                inst = cls(group, m[0][0], m[1][0]) if m is not None else False
            cls._instances[key] = inst
        return inst


class _RawCuda:
    """A raw device allocation presented through __cuda_array_interface__ (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, n, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (ptr, False), 'version': 2}


class PeerGrad:
    """Flat fp32 gradient buffer in IPC-shareable memory + the peer-memory all-reduce over it (gs_grad_allreduce):
    every rank maps every other rank's buffer, a range is summed in three capturable launches (no NCCL call)."""

    def __init__(self, numel, device, group, own, tables):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.numel = numel
        self.gptrs, self.fptrs = tables
        self.tensor = torch.as_tensor(_RawCuda(own[0], numel), device=device)     # zero-initialised by gs_ipc_alloc
        self.seq = torch.zeros(1, dtype=torch.int64, device=device)

    def all_reduce(self, offset=0, count=None):
        count = self.numel - offset if count is None else count
        call('gs_grad_allreduce', self.gptrs, offset, count, self.fptrs, self.rank, self.world, self.seq.data_ptr(), _stream())

    @classmethod
    def create(cls, numel, device, group=None):
        """PeerGrad, or None when the job is single-rank / peer access is disabled or unavailable on ANY rank (then
        NCCL is used by all of them).  Collective: every rank must call it."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        want = os.environ.get('GS_GRAD_PEER', '1') != '0' and bool(PeerExchange.get(group))
        if not _all_agree(want, group):
            return None
        m = _peer_map([numel * 4, int(_lib.load().gs_comm_flags_bytes())], group, 'peer-memory gradient buffer')
        return cls(numel, device, group, m[0], m[1]) if m is not None else None


def stats_all_reduce(stats, group=None, dgamma=None, dbeta=None):
    """Sum the packed fp64 statistics over the SyncBN group (peer-memory kernel, NCCL as fallback).  `dgamma` / `dbeta`
    (device pointers or None): BN parameter gradients, incremented by the LOCAL sums before the exchange."""
    ex = PeerExchange.get(group)
    if ex:
        ex.all_reduce(stats, dgamma, dbeta)
    else:
        if dgamma or dbeta:
            call('gs_bn_bwd_param', stats.data_ptr(), stats.numel() // 2, dgamma, dbeta, 1, _stream())
        dist.all_reduce(stats, group=group)


# `group_size` of DynSyncBN (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23 passes group_size=1): gaiavision is
# not in the reference tree, so its meaning is an OPEN QUESTION (DESIGN.md "open questions").  Two readings, selectable:
#   GS_SYNCBN_GROUP_SIZE=world (default)  the argument is accepted and ignored: statistics over the whole DP world -- what
#                                         north_star prescribes ("SyncBN per-channel sum/sumsq across the 8 GPUs");
#   GS_SYNCBN_GROUP_SIZE=ranks            group_size = ranks per synchronisation group (the linklink / SenseTime SyncBN
#                                         convention): 1 -> per-GPU statistics (no exchange), k -> groups of k consecutive ranks.
GROUP_SIZE_MODE = os.environ.get('GS_SYNCBN_GROUP_SIZE', 'world')
_subgroups = {}


def _rank_subgroup(k):
    """Process group of the k consecutive ranks containing this rank (all ranks create all groups, once)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if k not in _subgroups:
        mine = None
        for lo in range(0, world, k):
            pg = dist.new_group(list(range(lo, min(lo + k, world))))
            if lo <= rank < lo + k:
                mine = pg
        _subgroups[k] = mine
    return _subgroups[k]


def _sync_group(bn):
    """(process_group, world) when `bn` synchronises statistics across ranks, else (None, 1)."""
    if not getattr(bn, 'sync', False) or not dist.is_available() or not dist.is_initialized():
        return None, 1
    pg = getattr(bn, 'process_group', None)
    if pg is None and GROUP_SIZE_MODE == 'ranks':
        k = getattr(bn, 'group_size', None)
        if k is not None and k < dist.get_world_size():
            if k <= 1:
                return None, 1
            pg = _rank_subgroup(int(k))
    world = dist.get_world_size(pg)
    return pg, world


def syncbn_push_desc(bn):
    """Descriptor for the PRODUCER of `bn`'s statistics (GS_SYNCBN_FOLD=2, several ranks, peer memory available), else None."""
    if FOLD_EXCHANGE != 2:
        return None
    pg, world = _sync_group(bn)
    if world <= 1:
        return None
    ex = PeerExchange.get(pg)
    return ex.push_ref if ex else None


def bn_train_apply(bn, y, stats, C, residual=None, relu=False, pushed=False):
    """all-reduce the packed (sum, sumsq) over the SyncBN group, then ONE kernel: finalize (mean / invstd / scale /
    shift, running-stat update of the channel prefix) + normalise + residual + ReLU.  Returns (z, aff, count)."""
    pg, world = _sync_group(bn)
    sync = None
    if world > 1:
        ex = PeerExchange.get(pg)
        if ex and pushed:
            sync = ex.poll_ref            # the producer of `stats` has pushed them: block 0 of the apply kernel only polls
        elif ex and FOLD_EXCHANGE == 1:
            sync = ex.desc_ref            # block 0 of the apply kernel exchanges the sums itself: no extra launch
        else:
            if pushed:
                raise GsError('bn_train_apply: statistics were pushed but the peer exchange is not available')
            stats_all_reduce(stats, pg)
    count = float(_pixels(y)) * world
    aff = torch.empty((4, C), dtype=torch.float32, device=y.device)
    upd = bn.training and bn.track_running_stats and bn.running_mean is not None
    if upd and bn.momentum is None:
        raise GsError('DynamicBatchNorm2d: momentum=None (cumulative average) is not supported on the CUDA path')
    N, _, H, W = y.shape
    z = new_act(N, C, H, W, y.device)
    res_ld = 0
    if residual is not None:
        residual = as_act(residual)
        res_ld = act_ld(residual)
    call('gs_bn_apply_train', y.data_ptr(), act_ld(y), stats.data_ptr(), count, _ptr(bn.weight), _ptr(bn.bias),
         bn.running_mean.data_ptr() if upd else None, bn.running_var.data_ptr() if upd else None,
         float(bn.momentum if bn.momentum is not None else 0.0), float(bn.eps), aff.data_ptr(), _ptr(residual), res_ld,
         1 if relu else 0, z.data_ptr(), C, _pixels(y), C, sync, _stream())
    if upd:
        bn._gs_nbt_pending = getattr(bn, '_gs_nbt_pending', 0) + 1
        if _touched_bns is not None:
            _touched_bns.append(bn)
    return z, aff, count


def bn_backward(bn, dz, y, aff, count, zmask, relu, want_dres):
    """Per-channel sums (one reduction kernel), all-reduce over the SyncBN group, then ONE kernel for dy (+ dres,
    + parameter grads)."""
    N, C, H, W = dz.shape
    P, dev, st = N * H * W, dz.device, _stream()
    mean, invstd, scale, shift = aff[0], aff[1], aff[2], aff[3]
    zl = act_ld(zmask) if zmask is not None else 0
    gw = bn.weight is not None and bn.weight.requires_grad
    gb = bn.bias is not None and bn.bias.requires_grad
    dgam = _param_grad(bn.weight).data_ptr() if gw else None
    dbet = _param_grad(bn.bias).data_ptr() if gb else None
    pg, world = _sync_group(bn)
    ex = PeerExchange.get(pg) if world > 1 else None
    sums = zeros_f64(2 * C + 2, dev)[:2 * C]
    dy = new_act(N, C, H, W, dev)
    dres = new_act(N, C, H, W, dev) if want_dres else None
    fused = FUSED_BN_BWD == '1' or (FUSED_BN_BWD == 'auto' and P * C * 2 <= FUSED_BN_MAX_BYTES)
    if fused and (world == 1 or ex):
        # ONE launch: partial sums -> cluster barrier (-> per-cluster peer exchange, parameter gradients from the local
        # sums) -> apply; the second pass over dz / y comes from L1 / L2
        call('gs_bn_bwd', dz.data_ptr(), act_ld(dz), y.data_ptr(), act_ld(y), _ptr(zmask), zl, mean.data_ptr(),
             invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), 1 if relu else 0, _ptr(bn.weight), sums.data_ptr(),
             float(count), P, C, dy.data_ptr(), C, _ptr(dres), C, dgam, dbet, ex.desc_ref if world > 1 else None, st)
        return dy, dres
    split = ex is not None and bool(ex) and FOLD_EXCHANGE == 2      # reduce pushes, apply polls: no exchange launch
    call('gs_bn_bwd_reduce', dz.data_ptr(), act_ld(dz), y.data_ptr(), act_ld(y), _ptr(zmask), zl, mean.data_ptr(),
         invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), 1 if relu else 0, P, C, sums.data_ptr(),
         ex.push_ref if split else None, st)
    if world > 1 and not split:
        # parameter gradients come from the LOCAL sums (the gradient all-reduce averages them later): same kernel
        stats_all_reduce(sums, pg, dgam, dbet)
        dgam = dbet = None
    call('gs_bn_bwd_apply', dz.data_ptr(), act_ld(dz), y.data_ptr(), act_ld(y), _ptr(zmask), zl, mean.data_ptr(),
         invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), 1 if relu else 0, _ptr(bn.weight), sums.data_ptr(),
         float(count), P, C, dy.data_ptr(), C, _ptr(dres), C, dgam, dbet, ex.poll_ref if split else None, st)
    return dy, dres


def bn_eval_affine(bn, C):
    aff = torch.empty((2, C), dtype=torch.float32, device=bn.running_mean.device)
    call('gs_bn_eval_affine', C, _ptr(bn.weight), _ptr(bn.bias), bn.running_mean.data_ptr(),
         bn.running_var.data_ptr(), float(bn.eps), aff[0].data_ptr(), aff[1].data_ptr(), _stream())
    return aff


def bn_stats(x):
    x = as_act(x)
    C = x.shape[1]
    stats = zeros_f64(2 * C + 2, x.device)[:2 * C]
    call('gs_bn_stats', x.data_ptr(), _pixels(x), C, act_ld(x), stats.data_ptr(), _stream())
    return stats


def bn_apply(y, scale, shift, residual=None, relu=False):
    N, C, H, W = y.shape
    z = new_act(N, C, H, W, y.device)
    res_ld = 0
    if residual is not None:
        residual = as_act(residual)
        res_ld = act_ld(residual)
    call('gs_bn_apply', y.data_ptr(), act_ld(y), _ptr(scale), _ptr(shift), _ptr(residual), res_ld, 1 if relu else 0,
         z.data_ptr(), C, _pixels(y), C, _stream())
    return z


def _param_grad(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


# ------------------------------------------------------------------------------------------------
# fused conv -> (dynamic BN) -> (ReLU) (+ residual): forward record + hand-written backward
# ------------------------------------------------------------------------------------------------
class LayerRec:
    __slots__ = ('conv', 'bn', 'mode', 'a', 'x_shape', 'geom', 'y', 'z', 'aff', 'count', 'relu', 'has_res', 'Co')


def cba_forward(x, conv, bn=None, relu=False, residual=None, Co=None, save=True):
    """z = relu?( BN( conv(x) ) + residual ).  Returns (z, rec); rec drives cba_backward."""
    Co = conv.width_state if Co is None else Co
    rec = LayerRec() if save else None
    if bn is not None and bn_batch_mode(bn):
        push = syncbn_push_desc(bn)
        if CONV_BN_FUSE and push is None and _sync_group(bn)[1] == 1 and Co % 8 == 0:
            fb = dict(bn=bn, residual=residual, relu=relu)
            y, stats, a, g = conv_forward(x, conv, Co, shift=_bias(conv, Co), want_stats=True, fuse_bn=fb)
            z, aff, count = fb['z'], fb['aff'], fb['count']
        else:
            y, stats, a, g = conv_forward(x, conv, Co, shift=_bias(conv, Co), want_stats=True, sync=push)
            z, aff, count = bn_train_apply(bn, y, stats, Co, residual, relu, pushed=push is not None)
        mode = 'bn_batch'
    else:
        y = None
        if bn is not None:
            aff = bn_eval_affine(bn, Co)
            scale, shift = aff[0], aff[1]
            if conv.bias is not None:
                raise GsError('conv bias followed by eval-mode BN is not supported')
            mode = 'affine'
        else:
            aff, scale, shift = None, None, _bias(conv, Co)
            mode = 'plain'
        z, _, a, g = conv_forward(x, conv, Co, scale=scale, shift=shift, residual=residual, relu=relu)
        count = 0.0
    if save:
        rec.conv, rec.bn, rec.mode, rec.a, rec.x_shape, rec.geom = conv, bn, mode, a, tuple(x.shape), g
        rec.y, rec.z, rec.aff, rec.count, rec.relu, rec.has_res, rec.Co = y, z, aff, count, relu, residual is not None, Co
    return z, rec


def _bias(conv, Co):
    return None if conv.bias is None else conv.bias[:Co]


# ------------------------------------------------------------------------------------------------
# weight gradients on a side stream
# ------------------------------------------------------------------------------------------------
# wgrad(L) needs only dy(L) and the saved input of layer L; nothing on the backward chain (dgrad(L) -> BN backward of
# L-1 -> ...) waits for it.  It is therefore enqueued on a second stream, where the tensor-core kernel overlaps the
# memory-bound BN-backward kernels of the following layers (they co-reside on the SMs: wgrad takes the shared memory,
# the BN blocks only registers).  ONE join per backward pass, installed as an autograd final callback; the operands
# are kept alive until then.  In a captured iteration the fork / join become parallel branches of the graph.
WGRAD_STREAM = os.environ.get('GS_WGRAD_STREAM', '1') == '1'
_side_streams = {}
_side_state = {'pending': [], 'main': None, 'queued': False}


def _side_stream(device):
    st = _side_streams.get(device.index)
    if st is None:
        st = _side_streams[device.index] = torch.cuda.Stream(device=device)
    return st


def wgrad_join():
    """Make the stream the backward pass ran on wait for the side-stream work (idempotent)."""
    stt = _side_state
    if stt['pending']:
        dev = stt['pending'][0][0]
        (stt['main'] or torch.cuda.current_stream(dev)).wait_stream(_side_stream(dev))
        stt['pending'].clear()
    stt['queued'], stt['main'] = False, None


def side_stream_run(fn, device, keep=()):
    """Run `fn()` on the side stream, ordered after everything enqueued so far on the current stream; the current
    stream re-joins at the end of the running backward pass.  Outside a backward pass `fn` simply runs in stream order.
    `keep`: objects that must stay alive until the join (operands of the side-stream kernels).
    Returns True: `fn` HAS run (been enqueued) in either case -- callers that track "this range is done" rely on it."""
    stt = _side_state
    main = torch.cuda.current_stream(device)
    if not stt['queued']:
        try:
            torch.autograd.Variable._execution_engine.queue_callback(wgrad_join)
        except RuntimeError:          # not inside a backward pass: keep the plain stream order
            fn()
            return True
        stt['queued'], stt['main'] = True, main
    side = _side_stream(device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        fn()
    stt['pending'].append((device, keep))
    return True


def reset_side_state():
    """Called at zero_grad(): a backward pass that raised before its final callback ran must not leave the
    'callback queued' flag set (the next backward would then never re-join the side stream)."""
    stt = _side_state
    if stt['pending']:
        wgrad_join()
    stt['queued'], stt['main'] = False, None


def _wgrad_async(conv, a, dy, geom):
    if not WGRAD_STREAM or PROFILE is not None or _lib.PROFILE_CALLS is not None:
        conv_wgrad(conv, a, dy, geom)
        return
    side_stream_run(lambda: conv_wgrad(conv, a, dy, geom), dy.device, keep=(a, dy))


# Gradient chunks (overlapped all-reduce).  runner.FlatParams builds a PLAN from the model structure: every flat range is
# tagged with the res stage whose backward, once enqueued, makes the range final BY DATA DEPENDENCY:
#   * the parameters of backbone stage k           -> final when StageFn.backward of stage k has been enqueued;
#   * a head reading feature `in_index` = stage k  -> final when the backward of stage k has been enqueued (that node
#     consumes the head's input gradient, so the whole head ran before it -- no reliance on autograd's queue priority);
#   * everything else (stem, necks, ...)           -> reduced by all_reduce_grads() after the backward pass.
# The first block of a planned stage carries (weakref(owner), stage index); StageFn.backward reports it here.
def _stage_grads_done(blocks):
    if PROFILE is not None or _lib.PROFILE_CALLS is not None:
        return
    tag = getattr(blocks[0], '_gs_grad_stage', None)
    if tag is None:
        return
    owner = tag[0]()
    if owner is not None:
        owner._reduce_stage(tag[1])


def cba_backward(rec, dz, need_dx=True, dx_add=None):
    """Returns (dx, dres): gradient w.r.t. the conv input (None unless need_dx; `dx_add` is summed in the dgrad
    epilogue) and w.r.t. the residual."""
    dz = as_act(dz)
    conv, bn, C = rec.conv, rec.bn, rec.Co
    N, _, Ho, Wo = dz.shape
    P = N * Ho * Wo
    dev = dz.device
    st = _stream()
    zmask = rec.z if rec.relu else None
    if rec.mode == 'bn_batch':
        # without a residual the ReLU mask is recomputed from y (one tensor read less than reading z)
        dy, dres = bn_backward(bn, dz, rec.y, rec.aff, rec.count, zmask if rec.has_res else None, rec.relu, rec.has_res)
    else:
        dres = new_act(N, C, Ho, Wo, dev) if rec.has_res else None
        scale = rec.aff[0] if rec.mode == 'affine' else None
        if scale is None and zmask is None and dres is None:
            dy = dz
        else:
            dy = new_act(N, C, Ho, Wo, dev)
            call('gs_affine_bwd', dz.data_ptr(), act_ld(dz), _ptr(zmask), act_ld(zmask) if zmask is not None else 0,
                 _ptr(scale), P, C, dy.data_ptr(), C, _ptr(dres), C, st)
        if conv.bias is not None and conv.bias.requires_grad:
            s = bn_stats(dy)
            call('gs_bn_bwd_param', s.data_ptr(), C, None, _param_grad(conv.bias).data_ptr(), 1, st)
    _wgrad_async(conv, rec.a, dy, rec.geom)
    dx = None
    if need_dx:
        if is_image_conv(conv):
            raise GsError('gradient w.r.t. the input image is not provided by the hot path')
        dx = conv_dgrad(conv, dy, rec.geom, rec.x_shape, add=dx_add)
    return dx, dres


# ------------------------------------------------------------------------------------------------
# autograd wrappers
# ------------------------------------------------------------------------------------------------
class ConvBnActFn(torch.autograd.Function):
    """One fused layer.  Parameter gradients are accumulated IN PLACE into `.grad` (KRSC fp32 for the
    conv weight) by the kernels; autograd only routes activation gradients."""

    @staticmethod
    def forward(ctx, x, residual, weight, conv, bn, relu, Co, carrier=None):
        save = any(ctx.needs_input_grad)
        z, rec = cba_forward(x, conv, bn, relu, residual, Co, save=save)
        ctx.rec, ctx.carrier = rec, carrier
        return z

    @staticmethod
    def backward(ctx, dz):
        # a second consumer of x (the skip branch of a concat, see GradCarrier) hands its gradient over here: it is summed
        # in the dgrad epilogue instead of by a separate strided add pass of the autograd engine
        add = ctx.carrier.take() if ctx.carrier is not None else None
        need_dx = ctx.needs_input_grad[0]
        dx, dres = cba_backward(ctx.rec, dz, need_dx=need_dx, dx_add=add if need_dx else None)
        ctx.rec = ctx.carrier = None
        return dx, dres, None, None, None, None, None, None


SKIP_GRAD_CARRIER = os.environ.get('GS_SKIP_CARRIER', '1') != '0'   # 0: let autograd add the two gradients (A/B, tests)


class GradCarrier:
    """Side channel for the gradient of a tensor x that feeds BOTH a conv layer and a channel concat (FCN head with
    concat_input, fcn_head.py:68-81): the concat's backward deposits its channel-slice VIEW of the incoming gradient here
    instead of returning it, and the conv layer's backward -- which by data dependency runs later -- adds it inside its
    dgrad epilogue.  Saves autograd's accumulation pass over two differently-pitched [N, C, H, W] tensors."""
    __slots__ = ('g',)

    def __init__(self):
        self.g = None

    def put(self, g):
        self.g = g

    def take(self):
        g, self.g = self.g, None
        return g


def conv_bn_act(x, conv, bn=None, relu=False, residual=None, Co=None, grad_carrier=None):
    return ConvBnActFn.apply(x, residual, conv.weight, conv, bn, relu, Co, grad_carrier)


class BottleneckFn(torch.autograd.Function):
    """DynamicBottleneck as ONE autograd node: 3 (4 with downsample) fused conv/BN layers; the residual
    gradient is folded into conv1's dgrad epilogue, so no separate add pass exists in either direction."""

    @staticmethod
    def forward(ctx, x, weight, block):
        z3, ctx.recs = _bottleneck_forward(x, block, any(ctx.needs_input_grad))
        return z3

    @staticmethod
    def backward(ctx, dout):
        recs, ctx.recs = ctx.recs, None
        dx = _bottleneck_backward(recs, dout, ctx.needs_input_grad[0])
        return dx, None, None


def _bottleneck_forward(x, block, save):
    z1, r1 = cba_forward(x, block.conv1, block.norm1, True, save=save)
    z2, r2 = cba_forward(z1, block.conv2, block.norm2, True, save=save)
    rd = None
    identity = x
    if block.downsample is not None:
        identity, rd = cba_forward(x, block.downsample[0], block.downsample[1], False, save=save)
    z3, r3 = cba_forward(z2, block.conv3, block.norm3, True, residual=identity, save=save)
    return z3, (r1, r2, r3, rd)


def _bottleneck_backward(recs, dout, need_dx=True):
    """Hand-written backward of one bottleneck: the residual gradient (through the downsample branch when there is one)
    is added inside conv1's dgrad epilogue.  Returns dx."""
    r1, r2, r3, rd = recs
    d2, dres = cba_backward(r3, dout)
    d1, _ = cba_backward(r2, d2)
    add = dres
    if rd is not None:
        add, _ = cba_backward(rd, dres)
    if not need_dx:
        cba_backward(r1, d1, need_dx=False)
        return None
    dx, _ = cba_backward(r1, d1, dx_add=add)
    return dx


class StageFn(torch.autograd.Function):
    """DynamicResLayer.forward (the first `depth` blocks of a stage) as ONE autograd node: ~10x fewer autograd nodes
    than one per layer, and the end of its backward is where the gradients of the stage (and of everything after it)
    are final -> start of their all-reduce chunk."""

    @staticmethod
    def forward(ctx, x, weight, blocks):
        save = any(ctx.needs_input_grad)
        recs = []
        for blk in blocks:
            x, r = _bottleneck_forward(x, blk, save)
            recs.append(r)
        ctx.recs = recs
        ctx.blocks = blocks
        return x

    @staticmethod
    def backward(ctx, dout):
        recs, ctx.recs = ctx.recs, None
        d = dout
        for b in range(len(recs) - 1, -1, -1):
            d = _bottleneck_backward(recs[b], d, ctx.needs_input_grad[0] or b > 0)
        _stage_grads_done(ctx.blocks)
        return d, None, None


def _grad_anchor(modules):
    """A parameter that requires grad among `modules` (autograd builds the node iff an input needs a gradient: with a
    frozen first block and an input without grad -- frozen_stages / frozen_layers -- the later trainable blocks would
    otherwise silently train nothing); falls back to the first conv weight."""
    for m in modules:
        for p in m.parameters():
            if p.requires_grad:
                return p
    return modules[0].conv1.weight


def res_stage(x, blocks):
    return StageFn.apply(x, _grad_anchor(blocks), blocks)


def bottleneck(x, block):
    return BottleneckFn.apply(x, _grad_anchor([block]), block)


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _lib.require_device()
        x = as_act(x)
        N, C, H, W = x.shape
        Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        y = new_act(N, C, Ho, Wo, x.device)
        idx = None
        if ctx.needs_input_grad[0]:
            idx = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=x.device)
        call('gs_maxpool3x3s2_fwd', x.data_ptr(), N, H, W, C, act_ld(x), y.data_ptr(), Ho, Wo, C, _ptr(idx), _stream())
        ctx.idx, ctx.xs = idx, (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_act(dy)
        N, C, H, W = ctx.xs
        dx = new_act(N, C, H, W, dy.device)
        call('gs_maxpool3x3s2_bwd', dy.data_ptr(), act_ld(dy), ctx.idx.data_ptr(), N, H, W, C, dy.shape[2], dy.shape[3],
             dx.data_ptr(), C, _stream())
        ctx.idx = None
        return dx


def maxpool3x3s2(x):
    return MaxPoolFn.apply(x)


class CatFn(torch.autograd.Function):
    """torch.cat([a, b], dim=1) of NHWC activations as two pitched copies; the backward is two channel-slice
    VIEWS of the incoming gradient (the kernels take pitches), i.e. free."""

    @staticmethod
    def forward(ctx, carrier, *xs):
        xs = [as_act(x) for x in xs]
        N, _, H, W = xs[0].shape
        Cs = [x.shape[1] for x in xs]
        out = new_act(N, sum(Cs), H, W, xs[0].device)
        P, off, st = N * H * W, 0, _stream()
        for x, C in zip(xs, Cs):
            call('gs_copy_channels', x.data_ptr(), act_ld(x), out[:, off:off + C].data_ptr(), sum(Cs), P, C, st)
            off += C
        ctx.Cs, ctx.carrier = Cs, carrier
        return out

    @staticmethod
    def backward(ctx, d):
        d = as_act(d)
        outs, off = [], 0
        for C in ctx.Cs:
            outs.append(d[:, off:off + C])
            off += C
        if ctx.carrier is not None and ctx.needs_input_grad[1]:
            ctx.carrier.put(outs[0])          # picked up by the conv layer that also consumes xs[0] (GradCarrier)
            outs[0] = None
        ctx.carrier = None
        return (None,) + tuple(outs)


def cat_channels(xs, skip_carrier=None):
    """Channel concat.  skip_carrier: see GradCarrier -- xs[0]'s gradient goes through the carrier instead of autograd."""
    return CatFn.apply(skip_carrier, *xs)


# parity tests inject the Bernoulli draw here: callable(N, C, keep, device) -> fp32 [N, C] mask with values in
# {0, 1/keep}; None = torch's generator (the product behaviour)
DROPOUT_MASK_FN = None


class Dropout2dFn(torch.autograd.Function):
    """nn.Dropout2d (fcn_head.py:248-253): the per-(n, c) Bernoulli mask comes from torch's generator (a
    [N, C] tensor -- control-plane sized); applying it to the feature map is the kernel."""

    @staticmethod
    def forward(ctx, x, p):
        x = as_act(x)
        N, C, H, W = x.shape
        keep = 1.0 - p
        if DROPOUT_MASK_FN is not None:
            mask = DROPOUT_MASK_FN(N, C, keep, x.device).to(device=x.device, dtype=torch.float32).contiguous()
        else:
            mask = torch.bernoulli(torch.full((N, C), keep, dtype=torch.float32, device=x.device)).div_(keep)
        y = new_act(N, C, H, W, x.device)
        call('gs_scale_nc', x.data_ptr(), act_ld(x), mask.data_ptr(), y.data_ptr(), C, N, H * W, C, _stream())
        ctx.mask = mask
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = as_act(dy)
        N, C, H, W = dy.shape
        dx = new_act(N, C, H, W, dy.device)
        call('gs_scale_nc', dy.data_ptr(), act_ld(dy), ctx.mask.data_ptr(), dx.data_ptr(), C, N, H * W, C, _stream())
        return dx, None


def dropout2d(x, p, training):
    if not training or p <= 0:
        return x
    return Dropout2dFn.apply(x, p)


class ConvSegFn(torch.autograd.Function):
    """conv_seg: 1x1 DynamicConv2d + bias producing fp32 logits (the head's `@force_fp32` boundary)."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv, Co):
        y, _, a, g = conv_forward(x, conv, Co, shift=_bias(conv, Co), out_f32=True)
        ctx.conv, ctx.a, ctx.g, ctx.xs, ctx.Co = conv, a, g, tuple(x.shape), Co
        return y

    @staticmethod
    def backward(ctx, dlogits):
        conv, Co = ctx.conv, ctx.Co
        dl = as_act(dlogits, torch.float32) if act_ld(dlogits) is None or dlogits.dtype != torch.float32 else dlogits
        N, K, h, w = dl.shape
        P, st = N * h * w, _stream()
        if conv.bias is not None and conv.bias.requires_grad:
            call('gs_colsum_f32', dl.data_ptr(), act_ld(dl), P, K, _param_grad(conv.bias).data_ptr(), st)
        Kp = round_up(K, 8)
        dy = new_act(N, K, h, w, dl.device, BF16, ld=Kp)
        call('gs_cast_f32_bf16', dl.data_ptr(), act_ld(dl), dy.data_ptr(), Kp, P, K, st)
        conv_wgrad(conv, ctx.a, dy, ctx.g)
        dx = conv_dgrad(conv, dy, ctx.g, ctx.xs) if ctx.needs_input_grad[0] else None
        ctx.a = None
        return dx, None, None, None, None


def conv_seg(x, conv, Co=None):
    return ConvSegFn.apply(x, conv.weight, conv.bias, conv, conv.width_state if Co is None else Co)


class UpsampleCEFn(torch.autograd.Function):
    """losses(): bilinear resize to the label size -> CE(ignore_index) mean over ALL pixels * loss_weight,
    plus top-1 accuracy (dynamic_fcn_head.py:137-159).  Returns (loss, acc_seg, n_ignored, n_correct)."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index, loss_weight):
        _lib.require_device()
        lg = logits if (logits.dtype == torch.float32 and act_ld(logits) is not None) else as_act(logits, torch.float32)
        N, K, h, w = lg.shape
        lab = labels.reshape(N, labels.shape[-2], labels.shape[-1])
        if lab.dtype != torch.int64 or not lab.is_contiguous():
            lab = lab.long().contiguous()
        H, W = lab.shape[-2:]
        dev = lg.device
        out_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        rec = None
        if ctx.needs_input_grad[0]:
            rec = torch.empty(_lib.load().gs_upsample_ce_record_bytes(N, H, W), dtype=torch.uint8, device=dev)
        call('gs_upsample_ce_fwd', lg.data_ptr(), N, h, w, K, act_ld(lg), lab.data_ptr(), H, W, int(ignore_index),
             out_sum.data_ptr(), counts.data_ptr(), _ptr(rec), _stream())
        numel = float(N * H * W)
        loss = (out_sum * (loss_weight / numel)).float().squeeze(0)
        acc = (counts[1].double() * (100.0 / numel)).float()
        ctx.lg, ctx.rec, ctx.dims, ctx.gscale = lg, rec, (N, K, h, w, H, W), loss_weight / numel
        ctx.mark_non_differentiable(acc, counts)
        return loss, acc, counts

    @staticmethod
    def backward(ctx, dloss, dacc, dcounts):
        N, K, h, w, H, W = ctx.dims
        lg = ctx.lg
        dl = new_act(N, K, h, w, lg.device, torch.float32)
        ds = dloss.reshape(1).float().contiguous()
        call('gs_upsample_ce_bwd', lg.data_ptr(), N, h, w, K, act_ld(lg), ctx.rec.data_ptr(), H, W, float(ctx.gscale),
             ds.data_ptr(), dl.data_ptr(), K, _stream())
        ctx.rec = None
        return dl, None, None, None


def upsample_ce(logits, labels, ignore_index=255, loss_weight=1.0):
    return UpsampleCEFn.apply(logits, labels, ignore_index, loss_weight)


def upsample_argmax(logits, size):
    """Fused bilinear resize -> argmax over classes -> int64 [N, H, W] label map (inference)."""
    _lib.require_device()
    lg = logits if (logits.dtype == torch.float32 and act_ld(logits) is not None) else as_act(logits, torch.float32)
    N, K, h, w = lg.shape
    H, W = size
    out = torch.empty((N, H, W), dtype=torch.int64, device=lg.device)
    call('gs_upsample_argmax', lg.data_ptr(), N, h, w, K, act_ld(lg), H, W, out.data_ptr(), _stream())
    return out


def upsample_bilinear_f32(logits, size):
    lg = logits if (logits.dtype == torch.float32 and act_ld(logits) is not None) else as_act(logits, torch.float32)
    N, K, h, w = lg.shape
    H, W = size
    out = new_act(N, K, H, W, lg.device, torch.float32)
    call('gs_upsample_bilinear_f32', lg.data_ptr(), N, h, w, K, act_ld(lg), out.data_ptr(), H, W, K, _stream())
    return out


def to_nchw_f32(x):
    """bf16 NHWC activation -> fp32 NCHW contiguous tensor (what the reference's modules return)."""
    x = as_act(x)
    N, C, H, W = x.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    call('gs_nhwc_bf16_to_nchw_f32', x.data_ptr(), act_ld(x), N, C, H, W, out.data_ptr(), _stream())
    return out


def from_nchw_f32(x, Cpad=None):
    """fp32 NCHW tensor -> bf16 NHWC activation (channels padded with zeros to Cpad)."""
    _lib.require_device()
    x = x.float().contiguous()
    N, C, H, W = x.shape
    Cpad = round_up(C, 8) if Cpad is None else Cpad
    out = new_act(N, Cpad, H, W, x.device)
    call('gs_nchw_f32_to_nhwc_bf16', x.data_ptr(), N, C, H, W, out.data_ptr(), Cpad, Cpad, _stream())
    return out
