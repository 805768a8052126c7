"""Full-size (2x3x512x1024) graph capture of the PSP / ASPP heads on the MIN sub-net (debug aid)."""
import os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bench
import gaia_seg_b200 as gs
from config_cases import batch
dev = torch.device('cuda', 0)
MIN = {'backbone': {'stem': {'width': [16, 16, 32]}, 'body': {'width': [48, 96, 192, 384], 'depth': [2, 2, 5, 2]}}}
if os.environ.get('ARCH') == 'R101':
    MIN = {'backbone': {'stem': {'width': [32, 32, 64]}, 'body': {'width': [64, 128, 256, 512], 'depth': [3, 4, 23, 3]}}}
heads = {
    'psp': dict(type='DynamicPSPHead', conv_cfg=dict(type='DynConv2d'), in_channels=2560, in_index=3, channels=512,
                pool_scales=(1, 2, 3, 6), dropout_ratio=0.1, num_classes=19, norm_cfg=dict(type='SyncBN', requires_grad=True),
                align_corners=False, loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)),
    'aspp': dict(type='DynamicASPPHead', conv_cfg=dict(type='DynConv2d'), in_channels=2560, in_index=3, channels=512,
                 dilations=(1, 12, 24, 36), dropout_ratio=float(os.environ.get('DROP', '0.1')), num_classes=19,
                 norm_cfg=dict(type='SyncBN', requires_grad=True), align_corners=False,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)),
}
for name in sys.argv[1:] or ['psp', 'aspp']:
    try:
        cfg = bench.supernet_cfg('os8')
        cfg['decode_head'] = heads[name]
        if os.environ.get('SEED'):
            gs.set_random_seed(int(os.environ['SEED']))
        model = gs.build_segmentor(cfg, train_cfg=dict(), test_cfg=dict(mode='whole')).to(dev).train()
        opt = gs.GsSGD(model, lr=0.01, momentum=0.9, weight_decay=5e-4)
        model.manipulate_arch(MIN)
        data = batch(2, 512, 1024, 19, dev, 3)
        st = gs.GraphedTrainStep(model, opt, graph_after=2, max_graphs=2, pool_gb=float(os.environ.get('POOL', '8')))
        for it in range(5):
            out = st('min', data)
        torch.cuda.synchronize()
        print(name, 'ok', float(out['log_vars']['loss']), 'graphs', len(st.graphs), flush=True)
    except Exception as e:
        print(name, 'FAILED', str(e).splitlines()[0], flush=True)
        break
