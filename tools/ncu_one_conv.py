"""One conv shape, a few launches (target of `ncu --set full --import-source on -k regex:igemm -s 3 -c 1`)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg
N, H, W, Ci, Co, k, dil, stats = [int(a) for a in sys.argv[1:9]]
dev = torch.device('cuda')
conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
for _ in range(6):
    Fg.conv_forward(x, conv, Co, want_stats=bool(stats))
torch.cuda.synchronize()
