"""Fused upsample + CE forward / backward at the benchmark size (target of ncu -k regex:upsample_ce)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg
N, K, h, w, H, W = [int(a) for a in sys.argv[1:7]] if len(sys.argv) > 6 else (2, 19, 64, 128, 512, 1024)
g = torch.Generator().manual_seed(0)
logits = (torch.randn(N, K, h, w, generator=g) * 2).cuda().contiguous(memory_format=torch.channels_last)
lab = torch.randint(0, K, (N, 1, H, W), generator=g)
lab[torch.rand(N, 1, H, W, generator=g) < 0.1] = 255
lab = lab.cuda()
ts = []
for i in range(5):
    lg = logits.clone().requires_grad_(True)
    loss, _, _ = Fg.upsample_ce(lg, lab, 255, 1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.backward(); e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1) * 1e3, 1))
print('backward us (incl. autograd overhead):', ts)
