"""Pipeline trace of one conv launch (CTA 0): prints when the TMA producer issued each stage, when the MMA warp saw it
full, and the per-tile epilogue phases (ns relative to kernel start)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg

def run(N, H, W, Ci, Co, k, dil, stats, res=False):
    dev = torch.device('cuda')
    conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
    x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
    r = Fg.as_act(torch.randn(N, Co, H, W, device=dev)) if res else None
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    for _ in range(3):
        Fg.conv_forward(x, conv, Co, want_stats=stats, residual=r)
    gs._lib.call('gs_debug_set_trace', buf.data_ptr())
    Fg.conv_forward(x, conv, Co, want_stats=stats, residual=r)
    torch.cuda.synchronize()
    gs._lib.call('gs_debug_set_trace', None)
    t = buf.cpu().tolist()
    t0 = t[0]
    rel = lambda v: (v - t0) if v else None
    prod = [rel(v) for v in t[16:80] if v]
    mma = [rel(v) for v in t[80:144] if v]
    epi = [[rel(v) for v in t[144 + 4 * i:148 + 4 * i]] for i in range(16) if t[144 + 4 * i]]
    print(json.dumps(dict(shape=[N * H * W, Ci, Co, k, dil], stats=stats, res=res, setup_ns=rel(t[1]), producer_issue_ns=prod[:24],
                          mma_full_ns=mma[:24], epilogue_tiles_ns=epi[:6])), flush=True)

run(2, 64, 128, 320, 1280, 1, 1, True)
run(2, 64, 128, 320, 1280, 1, 1, False)
run(2, 64, 128, 320, 1280, 1, 1, False, res=True)
run(2, 64, 128, 320, 320, 3, 2, True)
run(2, 128, 256, 320, 80, 1, 1, True)
