# GAIA-seg supernet (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py recipe, OS8 "V1c" backbone of
# configs/local_examples/extract_subnet/psp_ar50to101_v1c_extract.py) with the FCN decode head of BASELINE
# config[1] on the synthetic Cityscapes-shaped dataset.  Search space: configs/_dynamic_/model_samplers/ar50to101v2.py.
norm_cfg = dict(type='DynSyncBN', requires_grad=True, group_size=1)
model = dict(
    type='DynamicEncoderDecoder',
    backbone=dict(type='DynamicResNet', in_channels=3, stem_width=[32, 32, 64], body_depth=[4, 6, 29, 4],
                  body_width=[80, 160, 320, 640], num_stages=4, out_indices=(0, 1, 2, 3), deep_stem=True,
                  strides=(1, 2, 1, 1), dilations=(1, 1, 2, 4), contract_dilation=True,
                  conv_cfg=dict(type='DynConv2d'), norm_cfg=norm_cfg, style='pytorch'),
    decode_head=dict(type='DynamicFCNHead', conv_cfg=dict(type='DynConv2d'), in_channels=2560, in_index=3, channels=512,
                     num_convs=2, concat_input=True, dropout_ratio=0.1, num_classes=19,
                     norm_cfg=dict(type='SyncBN', requires_grad=True), align_corners=False,
                     loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)),
    auxiliary_head=dict(type='DynamicFCNHead', conv_cfg=dict(type='DynConv2d'), in_channels=1280, in_index=2,
                        channels=256, num_convs=1, concat_input=False, dropout_ratio=0.1, num_classes=19,
                        norm_cfg=dict(type='SyncBN', requires_grad=True), align_corners=False,
                        loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.4)))
train_cfg = dict()
test_cfg = dict(mode='whole')

stem_width_range = dict(key='arch.backbone.stem.width', start=[16, 16, 32], end=[32, 32, 64], step=[8, 8, 16], ascending=True)
body_width_range = dict(key='arch.backbone.body.width', start=[48, 96, 192, 384], end=[80, 160, 320, 640],
                        step=[16, 32, 64, 128], ascending=True)
body_depth_range = dict(key='arch.backbone.body.depth', start=[2, 2, 5, 2], end=[4, 6, 29, 4], step=[1, 2, 2, 1])
MAX = {'name': 'MAX', 'arch.backbone.stem.width': [32, 32, 64], 'arch.backbone.body.width': [80, 160, 320, 640],
       'arch.backbone.body.depth': [4, 6, 29, 4]}
MIN = {'name': 'MIN', 'arch.backbone.stem.width': [16, 16, 32], 'arch.backbone.body.width': [48, 96, 192, 384],
       'arch.backbone.body.depth': [2, 2, 5, 2]}
R50 = {'name': 'R50', 'arch.backbone.stem.width': [32, 32, 64], 'arch.backbone.body.width': [64, 128, 256, 512],
       'arch.backbone.body.depth': [3, 4, 6, 3]}
R101 = {'name': 'R101', 'arch.backbone.stem.width': [32, 32, 64], 'arch.backbone.body.width': [64, 128, 256, 512],
        'arch.backbone.body.depth': [3, 4, 23, 3]}
random_subnet = dict(type='composite', model_samplers=[dict(type='range', **stem_width_range),
                                                        dict(type='range', **body_width_range),
                                                        dict(type='range', **body_depth_range)])
sandwich = True
max_net, min_net, sample_subnet_num = MAX, MIN, 2
train_sampler = None          # built from (max_net, min_net, random_subnet) because sandwich = True
val_sampler = dict(type='anchor', anchors=[R50, R101])
flops_sampler = dict(type='anchor', anchors=[MAX, MIN, R50, R101])   # tools/count_flops.py

data = dict(samples_per_gpu=2, workers_per_gpu=2,
            train=dict(type='SyntheticSegDataset', size=(512, 1024), num_classes=19, length=256),
            val=dict(type='SyntheticSegDataset', size=(512, 1024), num_classes=19, length=8),
            test=dict(type='SyntheticSegDataset', size=(1024, 2048), num_classes=19, length=8))
log_config = dict(interval=10, hooks=[dict(type='TextLoggerHook', by_epoch=False)])
dist_params = dict(backend='nccl')
load_from = None
resume_from = None
workflow = [('train', 1)]
optimizer = dict(type='SGD', lr=0.01, momentum=0.9, weight_decay=0.0005)
optimizer_config = dict()
lr_config = dict(policy='poly', power=0.9, min_lr=0.0001, by_epoch=False)
runner = dict(type='IterBasedRunner', max_iters=80000)
checkpoint_config = dict(by_epoch=False, interval=8000)
evaluation = dict(interval=8000, metric='mIoU')
