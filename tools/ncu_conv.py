"""One conv shape: forward (+ epilogue statistics), data gradient and weight gradient, a few rounds each -- the target of
    ncu --set full --import-source on -k regex:'igemm_kernel|wgrad_kernel' -s 9 -c 3 python tools/ncu_conv.py N H W Ci Co k dil
(3 warm-up rounds = 9 launches skipped, the 4th round captured)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg
N, H, W, Ci, Co, k, dil = [int(a) for a in sys.argv[1:8]]
rounds = int(sys.argv[8]) if len(sys.argv) > 8 else 4
dev = torch.device('cuda')
conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
y, stats, a, g = Fg.conv_forward(x, conv, Co, want_stats=True)
dy = Fg.as_act(torch.randn_like(y.float()))
torch.cuda.synchronize()
ts = {}
for r in range(rounds):
    for name, fn in (('fwd', lambda: Fg.conv_forward(x, conv, Co, want_stats=True)),
                     ('dgrad', lambda: Fg.conv_dgrad(conv, dy, g, tuple(x.shape))),
                     ('wgrad', lambda: Fg.conv_wgrad(conv, a, dy, g))):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.setdefault(name, []).append(round(e0.elapsed_time(e1) * 1e3, 1))
print({k_: v for k_, v in ts.items()})
