"""A few BN launches of one shape (target of ncu --set full -k regex:bn_)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg
N, H, W, C = [int(a) for a in sys.argv[1:5]]
dev = torch.device('cuda')
bn = gs.DynamicBatchNorm2d(C).to(dev).train()
y = Fg.as_act(torch.randn(N, C, H, W, device=dev))
dz = Fg.as_act(torch.randn(N, C, H, W, device=dev))
stats = Fg.bn_stats(y)
for _ in range(3):
    z, aff, count = Fg.bn_train_apply(bn, y, stats, C, relu=True)
    Fg.bn_backward(bn, dz, y, aff, count, None, True, False)
torch.cuda.synchronize()
