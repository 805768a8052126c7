#!/usr/bin/env python
"""bench.py -- supernet sandwich-training throughput (BASELINE.json metric: "supernet train imgs/s @512x1024").

One STEP = one sandwich cycle of the reference recipe (tools/train_supernet.py:180-187): four training iterations
[MAX, MIN, random, random] of the Dynamic ResNet supernet + FCN head, each iteration = sample sub-net ->
manipulate_arch -> forward -> fused upsample+CE loss -> backward -> (gradient all-reduce) -> fused SGD step, on a
batch of 2 synthetic Cityscapes-shaped images (3x512x1024, 19 classes, 10 % ignore) per GPU.  8 images per GPU per
step.  Weak scaling: per-GPU work is fixed as N grows; SyncBN statistics and gradients cross NVLink via NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--variant os8|os32] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same loop fed from
pinned host memory through the public API (`model.train_step` + optimizer), H2D copy of every batch and D2H read
of the loss inside the timed region.  `--impl reference` times the CPU oracle (oracle/ref_model.py -- the
reference's own mmseg/gaiavision stack is not installable here) on the host cores.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W, NUM_CLASSES, BATCH = 512, 1024, 19, 2
CYCLE = 4  # MAX, MIN, rand, rand


def supernet_cfg(variant):
    bb = dict(type='DynamicResNet', in_channels=3, stem_width=64, body_depth=[4, 6, 29, 4],
              body_width=[80, 160, 320, 640], num_stages=4, out_indices=(0, 1, 2, 3), conv_cfg=dict(type='DynConv2d'),
              norm_cfg=dict(type='DynSyncBN', requires_grad=True, group_size=1), style='pytorch')
    if variant == 'os8':   # configs/local_examples/extract_subnet/psp_ar50to101_v1c_extract.py:6-14
        bb.update(deep_stem=True, stem_width=[32, 32, 64], strides=(1, 2, 1, 1), dilations=(1, 1, 2, 4),
                  contract_dilation=True)
    return dict(type='DynamicEncoderDecoder', backbone=bb,
                decode_head=dict(type='DynamicFCNHead', conv_cfg=dict(type='DynConv2d'), in_channels=2560, in_index=3,
                                 channels=512, num_convs=2, concat_input=True, dropout_ratio=0.1,
                                 num_classes=NUM_CLASSES, norm_cfg=dict(type='SyncBN', requires_grad=True),
                                 align_corners=False,
                                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)))


def sampler_cfg(variant, seed=0):
    """sandwich = concat[MAX, MIN, random x2] over the search space of configs/_dynamic_/model_samplers/ar50to101v2.py."""
    deep = variant == 'os8'
    stem_max, stem_min = ([32, 32, 64], [16, 16, 32]) if deep else (64, 32)
    stem_range = (dict(type='range', key='arch.backbone.stem.width', start=[16, 16, 32], end=[32, 32, 64],
                       step=[8, 8, 16], ascending=True) if deep else
                  dict(type='range', key='arch.backbone.stem.width', start=32, end=64, step=16))
    width = dict(type='range', key='arch.backbone.body.width', start=[48, 96, 192, 384], end=[80, 160, 320, 640],
                 step=[16, 32, 64, 128], ascending=True)
    depth = dict(type='range', key='arch.backbone.body.depth', start=[2, 2, 5, 2], end=[4, 6, 29, 4], step=[1, 2, 2, 1])
    MAX = {'name': 'MAX', 'arch.backbone.stem.width': stem_max, 'arch.backbone.body.width': [80, 160, 320, 640],
           'arch.backbone.body.depth': [4, 6, 29, 4]}
    MIN = {'name': 'MIN', 'arch.backbone.stem.width': stem_min, 'arch.backbone.body.width': [48, 96, 192, 384],
           'arch.backbone.body.depth': [2, 2, 5, 2]}
    return MAX, MIN, dict(type='composite', model_samplers=[stem_range, width, depth])


def synth_batch(rank, idx, device=None, pin=False):
    import torch
    g = torch.Generator().manual_seed(1234 + rank * 7919 + idx)
    img = torch.randn(BATCH, 3, IMG_H, IMG_W, generator=g)
    lab = torch.randint(0, NUM_CLASSES, (BATCH, 1, IMG_H, IMG_W), generator=g)
    lab[torch.rand(BATCH, 1, IMG_H, IMG_W, generator=g) < 0.1] = 255
    if pin:
        img, lab = img.pin_memory(), lab.pin_memory()
    if device is not None:
        img, lab = img.to(device), lab.to(device)
    return img, lab


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([time.time()] + [c.strip() for c in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r[1:] for r in self.rows if self.t0 is None or (self.t0 <= r[0] <= self.t1 + 0.25)]
        self.rows = rows or [r[1:] for r in self.rows]
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


class CpuArm:
    """The CPU oracle (restatement of the reference's mmseg/gaiavision path, oracle/ref_model.py) on the host cores.
    One CPU step = one sandwich cycle [MAX, MIN, rand, rand] (forward + loss + backward + SGD each) at the GPU arm's
    per-GPU batch (2 x 3x512x1024 per iteration, 8 images per step, 4 rotating batches): literally one rank's step of
    the GPU arm -- a bounded sample of about 10-30 s of CPU work."""
    IMGS_PER_STEP = CYCLE * BATCH

    def __init__(self, variant):
        import torch
        from oracle import ref_model as O
        from gaia_seg_b200.model_space import build_model_sampler, fold_dict, sandwich_sampler_cfg
        self.torch, self.fold = torch, fold_dict
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.model = O.build_segmentor(supernet_cfg(variant))
        self.model.train()
        MAX, MIN, rnd = sampler_cfg(variant)
        self.sampler = build_model_sampler(sandwich_sampler_cfg(MAX, MIN, rnd, num_random=2, seed=0))
        self.opt = torch.optim.SGD(self.model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
        self.batches = [synth_batch(0, i) for i in range(4)]
        self.n_it = 0
        self.sample = (f'one sandwich cycle [MAX, MIN, rand, rand] of fwd+loss+bwd+SGD iterations at batch {BATCH}x3x{IMG_H}x{IMG_W} '
                       f'({CYCLE * BATCH} images = one rank\'s GPU step, same sub-net sampler and seed), torch CPU fp32 oracle')

    def iteration(self):
        self.model.manipulate_arch(self.fold(self.sampler.sample())['arch'])
        self.opt.zero_grad()
        img, lab = self.batches[self.n_it % len(self.batches)]
        self.n_it += 1
        loss = self.model.parse_losses(self.model.forward_train(img, None, lab))
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def step(self):
        for _ in range(CYCLE):
            self.iteration()


def run_reference(args):
    """--impl reference: times CpuArm steps with all host threads.  Rank 0 only."""
    if int(os.environ.get('RANK', 0)) != 0:
        return
    arm = CpuArm(args.variant)
    t0 = time.perf_counter()
    n_warm = args.warmup
    for i in range(args.warmup):
        arm.step()
        if time.perf_counter() - t0 > 60 and i + 1 < args.warmup:   # a CPU step takes ~15-30 s: bound the warm-up too
            n_warm = i + 1
            break
    per = (time.perf_counter() - t0) / max(n_warm, 1)
    steps = args.steps
    if args.warmup and per * steps > 150:      # keep the whole run within a few minutes
        steps = max(1, int(150 / per))
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step()
    dt = time.perf_counter() - t0
    v = steps * arm.IMGS_PER_STEP / dt
    print(json.dumps({
        'impl': 'reference', 'metric': 'supernet train imgs/s @512x1024', 'value': v, 'unit': 'imgs/s', 'n_gpus': args.gpus,
        'steps': steps, 'steps_requested': args.steps, 'warmup': n_warm, 'warmup_requested': args.warmup,
        'ms_per_step': 1e3 * dt / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(args.variant, args.gpus),
        'run': {'arm': 'reference CPU path (oracle port) on rank 0 only: ONE rank\'s share of the step '
                       f'({arm.IMGS_PER_STEP} images per step); value = its images/s', 'threads': arm.cores},
        'cpu_baseline': {'value': v, 'unit': 'imgs/s', 'cores': arm.cores, 'kind': 'port', 'sample': arm.sample},
        'e2e': {'value': v, 'unit': 'imgs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def bench_config(variant, world):
    """`config` of the JSON line -- the SAME dict for the B200 arm and the reference (CPU) arm."""
    return {'workload': workload_name(variant), 'variant': variant, 'per_gpu_batch': BATCH, 'global_batch': BATCH * world,
            'images_per_step': CYCLE * BATCH * world, 'parallelism': f'dp{world}',
            'l2': 'inputs + activations of every iteration (>1 GB) exceed the 126 MB L2; 4 rotating input batches',
            'subnet_sampler': 'sandwich [MAX, MIN, random x2], shared seed (all ranks draw the same sub-nets)'}


def workload_name(variant):
    return (f'BASELINE configs[1]: supernet sandwich training [MAX, MIN, rand, rand], DynamicResNet({variant}) + FCN head, '
            f'SyncBN, synthetic Cityscapes {IMG_H}x{IMG_W} crops, batch {BATCH}/GPU, one step = one 4-iteration cycle')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--variant', default='os8', choices=['os8', 'os32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--no-infer', action='store_true')
    ap.add_argument('--pool-gb', type=float, default=32.0, help='activation pool reserved up front (see runner.reserve_activation_pool)')
    ap.add_argument('--ncu-cycle', action='store_true')
    ap.add_argument('--graphs', type=int, default=1, help='CUDA-graph replay for recurring sub-nets (MAX / MIN)')
    ap.add_argument('--host-profile', action='store_true', help='cProfile one cycle -> gpurun_out/hostprof.txt')
    ap.add_argument('--kineto', action='store_true', help='torch.profiler (CUPTI) kernel times of one cycle -> gpurun_out/kineto_kernels.json')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import gaia_seg_b200 as gs
    from gaia_seg_b200 import functional as Fg
    from gaia_seg_b200.model_space import build_model_sampler, fold_dict, sandwich_sampler_cfg

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    os.environ.setdefault('GS_COMM_TIMEOUT_S', '120')   # a benchmark must trap, not hang, if a peer never shows up
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node N'
    gs._lib.require_device()
    gs.set_random_seed(0)

    # multi-rank correctness, CUDA vs CUDA on a small model, OUTSIDE the timed region (tools/multi_rank_parity.py): the
    # exchange kernels bit-exact against ordered sums, N ranks vs one rank on the concatenated batch, overlapped vs plain vs
    # NCCL gradients, buffers / parameters identical on all ranks -> `parity_multi` of the JSON line
    parity_multi = None
    if world > 1:
        from tools.multi_rank_parity import run as multi_rank_parity
        try:
            parity_multi = multi_rank_parity(gs)
        except Exception as e:   # noqa: BLE001  (report, do not hide the throughput numbers)
            parity_multi = {'ok': False, 'error': f'{type(e).__name__}: {e}'}

    model = gs.build_segmentor(supernet_cfg(args.variant), train_cfg=dict(), test_cfg=dict(mode='whole')).to(dev)
    model.train()
    opt = gs.GsSGD(model, lr=0.01, momentum=0.9, weight_decay=5e-4)
    if not args.graphs:
        gs.reserve_activation_pool(args.pool_gb, dev)
    MAX, MIN, rnd = sampler_cfg(args.variant)
    def new_sampler(seed):
        return build_model_sampler(sandwich_sampler_cfg(MAX, MIN, rnd, num_random=2, seed=seed))
    sampler_box = [new_sampler(0)]                 # shared seed: all ranks draw the same sub-nets, no broadcast needed
    n_dev_batches = 4
    dev_batches = [synth_batch(rank, i, device=dev) for i in range(n_dev_batches)]
    host_batches = [synth_batch(rank, 100 + i, pin=True) for i in range(n_dev_batches)]
    metas = [dict(ori_shape=(IMG_H, IMG_W, 3), flip=False)] * BATCH
    it_count = [0]

    # algorithmic conv FLOPs of every iteration that runs, counted analytically per sampled sub-net (complexity.py:
    # shape propagation, no device): forward 2*MACs + weight gradient 2*MACs + data gradient 2*(MACs - first conv: the
    # gradient w.r.t. the image is never computed), times the per-GPU batch
    from gaia_seg_b200.complexity import Counter, conv_macs
    flops_box, flops_cache = [0.0], {}

    def train_flops(key):
        f = flops_cache.get(key)
        if f is None:
            macs = conv_macs(model, (3, IMG_H, IMG_W))
            c0 = Counter()
            bb = model.backbone
            c0.conv(bb.stem[0] if bb.deep_stem else bb.conv1, (3, IMG_H, IMG_W))
            f = flops_cache[key] = BATCH * (6.0 * macs - 2.0 * c0.conv_macs)
        return f

    # with several ranks the graph holds forward + backward (incl. the peer-memory SyncBN exchanges); the NCCL gradient
    # all-reduce and the optimizer launch stay eager
    use_graphs = bool(args.graphs)
    graphed = gs.GraphedTrainStep(model, opt, graph_after=2, max_graphs=4, pool_gb=args.pool_gb) if use_graphs else None

    def iteration(img, lab):
        meta = fold_dict(sampler_box[0].sample())
        model.manipulate_arch(meta['arch'])
        batch = dict(img=img, img_metas=metas, gt_semantic_seg=lab)
        akey = json.dumps(meta['arch'], sort_keys=True)
        flops_box[0] += train_flops(akey)
        if graphed is not None and gs._lib.PROFILE_CALLS is None and Fg.PROFILE is None:
            out = graphed(akey, batch)
        else:
            out = model.train_step(batch, opt)
            opt.zero_grad()
            out['loss'].backward()
            w = opt.flat.all_reduce_grads()
            opt.grad_scale = 1.0 / w
            opt.step()
        it_count[0] += 1
        return out

    def step_resident():
        for _ in range(CYCLE):
            img, lab = dev_batches[it_count[0] % n_dev_batches]
            out = iteration(img, lab)
        return out

    # end-to-end feed: every iteration's batch travels pinned host memory -> device on a copy stream, one iteration
    # AHEAD of the compute stream (what a DataLoader with pin_memory + non_blocking copies does), so the 21 MB H2D
    # copy of iteration i+1 overlaps the kernels of iteration i
    copy_stream = torch.cuda.Stream()
    prefetched = [None]

    def h2d_async(idx):
        himg, hlab = host_batches[idx % n_dev_batches]
        with torch.cuda.stream(copy_stream):
            img = himg.to(dev, non_blocking=True)
            lab = hlab.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return img, lab, ev

    def step_e2e():
        total = None
        for _ in range(CYCLE):
            if prefetched[0] is None:
                prefetched[0] = h2d_async(it_count[0])
            img, lab, ev = prefetched[0]
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            img.record_stream(cur)
            lab.record_stream(cur)
            prefetched[0] = h2d_async(it_count[0] + 1)       # pinned host memory -> device, every iteration
            out = iteration(img, lab)
            total = out['loss'].detach() if total is None else total + out['loss'].detach()
        return total.item()                            # device -> host read of the step result (sum of the 4 losses)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    if args.ncu_cycle:   # profiling aid: exactly one sandwich cycle between cudaProfilerStart/Stop, no timing
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({'ncu_cycle': 'done', 'launches_per_cycle': None}))
        return
    if args.kineto and rank != 0:      # the other ranks take part in the profiled cycle (SyncBN exchanges)
        step_resident()
        torch.cuda.synchronize()
    if args.kineto and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_resident()
            torch.cuda.synchronize()
        agg = {}
        for ev in prof.events():
            if ev.device_type is not None and 'cuda' in str(ev.device_type).lower():
                d = agg.setdefault(ev.name.split('(')[0][:90], [0.0, 0])
                d[0] += ev.device_time_total if hasattr(ev, 'device_time_total') else ev.cuda_time_total
                d[1] += 1
        rows = sorted(({'kernel': k, 'ms': v[0] / 1e3, 'launches': v[1]} for k, v in agg.items()), key=lambda r: -r['ms'])
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'kineto_kernels.json'), 'w'), indent=0)
        # compact timeline (name, stream, start us, duration us) of every kernel of the cycle, for gap / overlap analysis
        try:
            tmp = os.path.join(ROOT, 'gpurun_out', '_trace_full.json')
            prof.export_chrome_trace(tmp)
            ev = json.load(open(tmp)).get('traceEvents', [])
            ker = [(e['name'].split('(')[0][:60], e.get('args', {}).get('stream'), e['ts'], e['dur']) for e in ev
                   if e.get('ph') == 'X' and e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
            ker.sort(key=lambda r: r[2])
            t0 = ker[0][2] if ker else 0
            json.dump([(n, s_, round(ts - t0, 3), round(d, 3)) for n, s_, ts, d in ker],
                      open(os.path.join(ROOT, 'gpurun_out', 'kineto_timeline.json'), 'w'))
            os.remove(tmp)
        except Exception as e:   # noqa: BLE001
            print('timeline export failed:', e, file=sys.stderr)
    if args.host_profile and rank != 0:
        step_resident()
        torch.cuda.synchronize()
    if args.host_profile and rank == 0:
        import cProfile
        import pstats
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        step_resident()
        pr.disable()
        torch.cuda.synchronize()
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'hostprof.txt'), 'w') as f:
            pstats.Stats(pr, stream=f).sort_stats('cumulative').print_stats(70)
            pstats.Stats(pr, stream=f).sort_stats('tottime').print_stats(45)
    # host-side enqueue time of one step (no sync inside): if it is close to ms_per_step the loop is CPU-bound
    torch.cuda.synchronize()
    t_h0 = time.perf_counter()
    step_resident()
    host_enqueue_ms = (time.perf_counter() - t_h0) * 1e3
    torch.cuda.synchronize()
    gs._lib.reset_launch_count()
    t_w0 = time.time()
    sampler_box[0] = new_sampler(1)                # the timed loops below all see the SAME sequence of random sub-nets
    ms0 = torch.cuda.memory_stats(dev)
    flops_box[0] = 0.0
    ms = timed(step_resident, args.steps)
    step_flops = flops_box[0] / args.steps           # per GPU, mean over exactly the timed steps
    ms1 = torch.cuda.memory_stats(dev)
    alloc_info = {'cudaMalloc_calls_in_timed_region': ms1.get('num_device_alloc', 0) - ms0.get('num_device_alloc', 0),
                  'reserved_GB': round(ms1.get('reserved_bytes.all.current', 0) / 2 ** 30, 2)}
    clocks.window(t_w0, time.time())
    launches = gs._lib.launch_count()
    clk = clocks.stop() if rank == 0 else {}
    imgs_per_step = CYCLE * BATCH * world
    value = imgs_per_step * args.steps / (ms / 1e3)

    step_e2e()
    sampler_box[0] = new_sampler(1)
    ms_e2e = timed(step_e2e, args.steps)
    sampler_box[0] = new_sampler(1)
    e2e_value = imgs_per_step * args.steps / (ms_e2e / 1e3)
    h2d = CYCLE * (BATCH * 3 * IMG_H * IMG_W * 4 + BATCH * IMG_H * IMG_W * 8)
    d2h = 4

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), CUDA events around every launch ----
    roof, breakdown = None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback (B200_PROFILING.md sustained 1.4 PF)'
    if not args.no_profile:
        # (a) every C-ABI call timed -> share of each kernel family in one cycle
        gs._lib.PROFILE_CALLS = []
        torch.cuda.synchronize()
        step_resident()
        torch.cuda.synchronize()
        calls, gs._lib.PROFILE_CALLS = gs._lib.PROFILE_CALLS, None
        fam = {}
        for name, a, b in calls:
            d = fam.setdefault(name, [0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += 1
        kernel_ms = {k: {'ms': round(v[0], 3), 'launches': v[1]} for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])}
        # (b) convolution launches with their algorithmic FLOPs
        Fg.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        step_resident()
        e1.record()
        torch.cuda.synchronize()
        prof, Fg.PROFILE = Fg.PROFILE, None
        cyc_ms = e0.elapsed_time(e1)
        agg = {}
        shapes = {}
        for kind, flops, a, b, shp in prof:
            sd = shapes.setdefault((kind,) + shp, [0.0, 0.0, 0])
            sd[0] += flops
            sd[1] += a.elapsed_time(b)
            sd[2] += 1
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += flops
            d[1] += a.elapsed_time(b)
            d[2] += 1
        ig_f = agg.get('fwd', [0, 0, 0])[0] + agg.get('dgrad', [0, 0, 0])[0]
        ig_ms = agg.get('fwd', [0, 0, 0])[1] + agg.get('dgrad', [0, 0, 0])[1]
        ig_n = agg.get('fwd', [0, 0, 0])[2] + agg.get('dgrad', [0, 0, 0])[2]
        ach = ig_f / (ig_ms * 1e-3) / 1e12 if ig_ms > 0 else 0.0
        # DRAM bytes per conv launch (dram__bytes_read.sum + dram__bytes_write.sum) come from the committed ncu capture
        # of the same kernels (bench.py cannot run under ncu itself): newest profiles/r*_igemm_dram_traffic.json
        traffic, traffic_src = None, None
        import glob
        cand = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_igemm_dram_traffic.json')))
        if cand:
            try:
                traffic = float(json.load(open(cand[-1]))['dram_bytes_per_launch'])
                traffic_src = os.path.relpath(cand[-1], ROOT)
            except Exception:
                traffic = None
        roof = {'bound': 'tensor', 'kernel': 'gs::igemm_kernel (conv fwd + dgrad)', 'achieved': ach, 'peak': peak_tf,
                'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': peak_src,
                'launches': ig_n, 'avg_launch_ms': ig_ms / max(ig_n, 1), 'flops_per_launch': ig_f / max(ig_n, 1)}
        breakdown = {k: {'tflops': v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0, 'ms': v[1], 'launches': v[2],
                         'share_of_step': v[1] / cyc_ms} for k, v in agg.items()}
        breakdown['profiled_step_ms'] = cyc_ms
        breakdown['c_abi_calls_ms'] = kernel_ms
        if rank == 0:
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            rows = [dict(kind=k[0], P=k[1], Ci=k[2], Co=k[3], k=k[4], stride=k[5], dil=k[6], launches=v[2], ms=v[1],
                         tflops=v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0) for k, v in shapes.items()]
            rows.sort(key=lambda r: -r['ms'])
            json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'conv_shapes.json'), 'w'), indent=0)
        breakdown['profiled_cycle_conv_flops'] = sum(v[0] for v in agg.values())

    # ---- second half of the BASELINE metric: sub-net inference imgs/s (config 5: whole-image 1x3x1024x2048, eval-mode
    # BN folded into the conv epilogue, fused resize+argmax, label map read back to the host), rank 0 only ----
    infer = None
    if rank == 0 and not args.no_infer:
        infer = infer_sweep(gs, model, dev, args.variant)
        model.train()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(args.variant)

    # whole-step roofline: analytic conv FLOPs of exactly the timed steps / measured step time, per GPU
    step_tf = step_flops / (ms / args.steps * 1e-3) / 1e12
    roof_step = {'bound': 'tensor', 'flops_per_step_per_gpu': step_flops, 'ms_per_step': ms / args.steps,
                 'achieved': step_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': step_tf / peak_tf, 'peak_source': peak_src,
                 'how': 'conv FLOPs (fwd + wgrad + dgrad, analytic per sampled sub-net, gaia_seg_b200/complexity.py) summed over '
                        'the timed iterations / CUDA-event time of the timed region; memory-bound kernels count as overhead'}
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {
            'metric': 'supernet train imgs/s @512x1024', 'value': value, 'unit': 'imgs/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': bench_config(args.variant, world),
            'run': {'arm': 'b200', 'cuda_graphs': 'MAX and MIN iterations replayed as CUDA graphs, random sub-nets eager'
                    if use_graphs else 'off',
                    'timing': 'CUDA events on the launching stream, barrier+sync both sides, max over ranks'},
            'clocks': clk, 'gpu_launches': launches, 'allocator': alloc_info, 'host_enqueue_ms_per_step': host_enqueue_ms,
            'e2e': {'value': e2e_value, 'unit': 'imgs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'roofline': roof, 'roofline_step': roof_step, 'parity_multi': parity_multi, 'cpu_baseline': cpu_base,
            'subnet_infer': infer, 'breakdown': breakdown}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def infer_sweep(gs, model, dev, variant, n_subnets=6, reps=3):
    """Per-sub-net whole-image inference throughput: MAX, MIN and random sub-nets (random.Random(0)), input
    1x3x1024x2048 from pinned host memory, output int64 label map on the host (model(return_loss=False, ...))."""
    import torch
    from gaia_seg_b200.model_space import build_model_sampler, fold_dict
    MAX, MIN, rnd = sampler_cfg(variant)
    rs = build_model_sampler(dict(rnd, seed=0))
    metas = [MAX, MIN] + [rs.sample() for _ in range(n_subnets - 2)]
    g = torch.Generator().manual_seed(7)
    himg = torch.randn(1, 3, 1024, 2048, generator=g).pin_memory()
    meta = [[dict(ori_shape=(1024, 2048, 3), flip=False)]]
    model.eval()
    per, tot_imgs, tot_s = [], 0, 0.0
    with torch.no_grad():
        for m in metas:
            model.manipulate_arch(fold_dict(m)['arch'])
            model(return_loss=False, img=[himg.to(dev, non_blocking=True)], img_metas=meta)      # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = model(return_loss=False, img=[himg.to(dev, non_blocking=True)], img_metas=meta)
            dt = time.perf_counter() - t0          # the .cpu() of the label map synchronises every call
            per.append({'subnet': m.get('name', 'random'), 'imgs_per_s': reps / dt})
            tot_imgs += reps
            tot_s += dt
    return {'value': tot_imgs / tot_s, 'unit': 'imgs/s', 'input': '1x3x1024x2048 whole-image, host->device->host',
            'n_subnets': len(metas), 'per_subnet': per, 'label_map_shape': list(out[0].shape), 'label_dtype': str(out[0].dtype)}


def cpu_baseline(variant):
    """cpu_baseline leg of the B200 arm: one CpuArm step (after one warm-up iteration), rank 0, N=1 only."""
    arm = CpuArm(variant)
    arm.iteration()                      # warm-up (MAX sub-net: touches every weight once)
    for _ in range(CYCLE - 1):           # finish the warm-up cycle so the timed step starts at MAX again
        arm.iteration()
    t0 = time.perf_counter()
    arm.step()
    dt = time.perf_counter() - t0
    return {'value': arm.IMGS_PER_STEP / dt, 'unit': 'imgs/s', 'cores': arm.cores, 'kind': 'port', 'sample': arm.sample,
            'seconds_per_step': dt}


if __name__ == '__main__':
    main()
