"""Time individual convolution launches (CUDA events, 20 reps after 5 warm-ups, L2 flushed between reps)."""
import os
import sys
import json
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg

SHAPES = [
    # P = N*H*W via (N,H,W), Ci, Co, k, dil
    ((2, 64, 128), 320, 1280, 1, 1), ((2, 64, 128), 1280, 320, 1, 1), ((2, 64, 128), 320, 320, 3, 2),
    ((2, 64, 128), 320, 256, 1, 1), ((2, 64, 128), 2560, 512, 3, 1), ((2, 128, 256), 320, 80, 1, 1),
    ((2, 128, 256), 80, 80, 3, 1), ((2, 128, 256), 80, 320, 1, 1), ((2, 64, 128), 192, 768, 1, 1),
    ((2, 64, 128), 768, 192, 1, 1), ((2, 64, 128), 640, 2560, 1, 1),
]


def timeit(fn, reps=20, warm=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us


def main():
    dev = torch.device('cuda')
    out = []
    for (N, H, W), Ci, Co, k, dil in SHAPES:
        conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
        x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
        res = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
        flops = 2.0 * N * H * W * Ci * Co * k * k
        row = dict(P=N * H * W, Ci=Ci, Co=Co, k=k, dil=dil)
        y, stats, a, g = Fg.conv_forward(x, conv, Co, want_stats=True)
        dy = Fg.as_act(torch.randn_like(y.float()))
        variants = {
            'fwd_stats': lambda: Fg.conv_forward(x, conv, Co, want_stats=True),
            'fwd_plain': lambda: Fg.conv_forward(x, conv, Co),
            'fwd_relu_res': lambda: Fg.conv_forward(x, conv, Co, relu=True, residual=y),
            'dgrad': lambda: Fg.conv_dgrad(conv, dy, g, tuple(x.shape)),
            'dgrad_add': lambda: Fg.conv_dgrad(conv, dy, g, tuple(x.shape), add=res),
            'wgrad': lambda: Fg.conv_wgrad(conv, a, dy, g),
        }
        for name, fn in variants.items():
            us = timeit(fn)
            row[name + '_us'] = round(us, 1)
            row[name + '_tf'] = round(flops / us / 1e6, 0)
        print(json.dumps(row), flush=True)
        out.append(row)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'conv_microbench.json'), 'w'), indent=0)


if __name__ == '__main__':
    main()
