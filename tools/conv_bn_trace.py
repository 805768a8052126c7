"""Pipeline trace of the fused conv + DynBN launch (CTA 0): conv epilogue end, tail entry, barrier, finalize, apply."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg

def run(N, H, W, Ci, Co, k, dil, res):
    dev = torch.device('cuda')
    conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
    bn = gs.DynamicBatchNorm2d(Co).to(dev).train()
    x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
    r = Fg.as_act(torch.randn(N, Co, H, W, device=dev)) if res else None
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    for _ in range(3):
        Fg.cba_forward(x, conv, bn, relu=True, residual=r, save=False)
    gs._lib.call('gs_debug_set_trace', buf.data_ptr())
    Fg.cba_forward(x, conv, bn, relu=True, residual=r, save=False)
    torch.cuda.synchronize()
    gs._lib.call('gs_debug_set_trace', None)
    t = buf.cpu().tolist()
    t0 = t[0]
    rel = lambda v: (v - t0) if v else None
    epi = [[rel(v) for v in t[144 + 4 * i:148 + 4 * i]] for i in range(16) if t[144 + 4 * i]]
    print(json.dumps(dict(shape=[N * H * W, Ci, Co, k, dil], res=res, last_epilogue_ns=epi[-1] if epi else None,
                          tail=dict(entry=rel(t[208]), loads_issued=rel(t[209]), cta_arrived=rel(t[210]), barrier_done=rel(t[211]),
                                    finalize_done=rel(t[212]), apply_done=rel(t[213])))), flush=True)

for a in sys.argv[1:]:
    v = [int(x) for x in a.split(',')]
    run(2, v[0], v[1], v[2], v[3], v[4], v[5], bool(v[6]) if len(v) > 6 else False)
