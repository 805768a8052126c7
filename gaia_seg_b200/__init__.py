"""gaia_seg_b200 -- B200-native (sm_100a) implementation of the GAIA-seg supernet train / eval hot path.

Host side: a Python mirror of the mmseg / gaiavision operator API the reference's hot path is written against
(build_segmentor, DynamicConv2d / DynamicBatchNorm2d, DynamicBottleneck, DynamicResNet, DynamicFCNHead,
DynamicEncoderDecoder, manipulate_arch + samplers, train_segmentor / test loops).  Device side:
libgaiaseg_b200.so, hand-written CUDA kernels behind the C ABI of include/gaiaseg_b200.h.  No CPU fallback.
"""
from . import _lib
from ._lib import GsError
from .core import (CONV_LAYERS, NORM_LAYERS, BatchNorm2d, DynamicBatchNorm2d, DynamicBottleneck, DynamicConv2d,
                   DynamicConvModule, DynamicMixin, DynamicSyncBatchNorm, Registry, SyncBatchNorm2d, build_conv_layer,
                   build_from_cfg, build_norm_layer)
from .backbone import BACKBONES, DynamicResLayer, DynamicResNet
from .heads import HEADS, LOSSES, CrossEntropyLoss, DynamicFCNHead, build_loss
from .psp_head import DynamicPPM, DynamicPSPHead
from .aspp_head import DynamicASPPHead, DynamicASPPModule
from .segmentor import (SEGMENTORS, DynamicEncoderDecoder, EncoderDecoder, build_backbone, build_head,
                        build_segmentor)
from .model_space import (MODEL_SAMPLERS, ManipulateArchHook, ModelSpaceManager, broadcast_object,
                          build_model_sampler, fold_dict, sandwich_sampler_cfg, unfold_dict)
from .runner import (Config, FlatParams, GraphedTrainStep, GsDataParallel, GsSGD, IterBasedRunner, build_optimizer, build_runner,
                     get_dist_info, init_dist, load_checkpoint, reserve_activation_pool, save_checkpoint, scatter_batch)
from .complexity import conv_macs, get_model_complexity_info
from .apis import (DATASETS, CrossArchEvalHook, DistCrossArchEvalHook, SyntheticSegDataset, build_dataloader,
                   build_dataset, multi_gpu_test, set_random_seed, single_gpu_test, train_segmentor)

__version__ = '0.1.0'
