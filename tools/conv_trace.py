"""Pipeline trace of one conv launch (CTA 0): prints when the TMA producer issued each stage, when the MMA warp saw it
full, and the per-tile epilogue phases (ns relative to kernel start)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gaia_seg_b200 as gs
from gaia_seg_b200 import functional as Fg

FLUSH = None
def run(N, H, W, Ci, Co, k, dil, stats, res=False, cold=False):
    global FLUSH
    dev = torch.device('cuda')
    conv = gs.DynamicConv2d(Ci, Co, k, padding=dil * (k // 2), dilation=dil, bias=False).to(dev)
    x = Fg.as_act(torch.randn(N, Ci, H, W, device=dev))
    r = Fg.as_act(torch.randn(N, Co, H, W, device=dev)) if res else None
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    for _ in range(3):
        Fg.conv_forward(x, conv, Co, want_stats=stats, residual=r)
    if cold:
        if FLUSH is None:
            FLUSH = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        FLUSH.fill_(1)
        torch.cuda.synchronize()
    gs._lib.call('gs_debug_set_trace', buf.data_ptr())
    Fg.conv_forward(x, conv, Co, want_stats=stats, residual=r)
    torch.cuda.synchronize()
    gs._lib.call('gs_debug_set_trace', None)
    t = buf.cpu().tolist()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        Fg.conv_forward(x, conv, Co, want_stats=stats, residual=r)
    e1.record(); torch.cuda.synchronize()
    warm_us = e0.elapsed_time(e1) * 50
    cold_us = 0.0
    if cold:
        for _ in range(5):
            FLUSH.fill_(1)
            e0.record(); Fg.conv_forward(x, conv, Co, want_stats=stats, residual=r); e1.record(); torch.cuda.synchronize()
            cold_us += e0.elapsed_time(e1) * 200
    t0 = t[0]
    rel = lambda v: (v - t0) if v else None
    prod = [rel(v) for v in t[16:80] if v]
    mma = [rel(v) for v in t[80:144] if v]
    epi = [[rel(v) for v in t[144 + 4 * i:148 + 4 * i]] for i in range(16) if t[144 + 4 * i]]
    last = max([v for v in t[144:208] if v] + [t0]) - t0
    print(json.dumps(dict(shape=[N * H * W, Ci, Co, k, dil], stats=stats, res=res, cold=cold, last_epi_ns=last, warm_us_b2b=round(warm_us, 1), cold_us=round(cold_us, 1), setup_ns=rel(t[1]), producer_issue_ns=prod[:24],
                          mma_full_ns=mma[:24], epilogue_tiles_ns=epi[:6])), flush=True)

if len(sys.argv) > 1:
    for a in sys.argv[1:]:
        v = [int(x) for x in a.split(',')]
        run(2, v[0], v[1], v[2], v[3], v[4], v[5], True, cold=bool(v[6]) if len(v) > 6 else False)
else:
    run(2, 64, 128, 320, 1280, 1, 1, True)
    run(2, 64, 128, 320, 1280, 1, 1, False)
    run(2, 64, 128, 320, 1280, 1, 1, False, res=True)
    run(2, 64, 128, 320, 320, 3, 2, True)
    run(2, 128, 256, 320, 80, 1, 1, True)
