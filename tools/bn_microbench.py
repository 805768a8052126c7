"""Time the DynBN kernels and wgrad in isolation (CUDA events, L2 flushed)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg
from tools.conv_microbench import timeit

dev = torch.device('cuda')
rows = []
for (N, H, W, C) in ((2, 64, 128, 320), (2, 64, 128, 1280), (2, 128, 256, 80), (2, 64, 128, 2560), (2, 128, 256, 320)):
    bn = gs.DynamicBatchNorm2d(C).to(dev).train()
    y = Fg.as_act(torch.randn(N, C, H, W, device=dev))
    dz = Fg.as_act(torch.randn(N, C, H, W, device=dev))
    stats = Fg.bn_stats(y)
    z, aff, count = Fg.bn_train_apply(bn, y, stats.clone(), C, relu=True)
    E = N * H * W * C
    r = dict(P=N * H * W, C=C, MB=round(E * 2 / 1e6, 1))
    us = timeit(lambda: Fg.bn_train_apply(bn, y, stats, C, relu=True)); r['apply_us'] = round(us, 1); r['apply_TBs'] = round(2 * E * 2 / us / 1e6, 2)
    us = timeit(lambda: Fg.bn_backward(bn, dz, y, aff, count, None, True, False)); r['bwd_us'] = round(us, 1); r['bwd_TBs'] = round(5 * E * 2 / us / 1e6, 2)
    us = timeit(lambda: Fg.bn_stats(y)); r['stats_us'] = round(us, 1); r['stats_TBs'] = round(E * 2 / us / 1e6, 2)
    print(json.dumps(r), flush=True)
for (N, H, W), Ci, Co, k, dil in (((2, 64, 128), 320, 1280, 1, 1), ((2, 64, 128), 1280, 320, 1, 1), ((2, 64, 128), 320, 320, 3, 2), ((2, 128, 256), 80, 320, 1, 1)):
    conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
    x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
    y, _, a, g = Fg.conv_forward(x, conv, Co)
    dy = Fg.as_act(torch.randn_like(y.float()))
    us = timeit(lambda: Fg.conv_wgrad(conv, a, dy, g))
    print(json.dumps(dict(wgrad=[N * H * W, Ci, Co, k], us=round(us, 1), tf=round(2.0 * N * H * W * Ci * Co * k * k / us / 1e6))), flush=True)
