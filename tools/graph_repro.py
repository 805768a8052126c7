"""Capture a whole iteration of the small PSP / ASPP / FCN models through GraphedTrainStep (debug aid)."""
import json, os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import gaia_seg_b200 as gs
import gs_checks as C
for name, kw in (('fcn', dict(aux=True)), ('psp', dict(psp=True, aux=True, os8=True)), ('aspp', dict(aspp=True, os8=True))):
    try:
        om, gm, _ = C.build_pair(gs, C.small_cfg(**kw), seed=1)
        opt = gs.GsSGD(gm, lr=0.05, momentum=0.9)
        st = gs.GraphedTrainStep(gm, opt, graph_after=1, max_graphs=2, pool_gb=2)
        gm.train()
        gm.manipulate_arch(C.SMALL_ARCHS['max'])
        g = torch.Generator().manual_seed(1)
        img = C.bf16r(torch.randn(2, 3, 64, 96, generator=g)).cuda()
        lab = C._labels(g, 2, 19, 64, 96).cuda()
        for it in range(4):
            out = st('max', dict(img=img, img_metas=[{}, {}], gt_semantic_seg=lab))
        torch.cuda.synchronize()
        print(name, 'ok', float(out['log_vars']['loss']), 'graphs', len(st.graphs), flush=True)
    except Exception:
        print(name, 'FAILED', flush=True)
        traceback.print_exc()
