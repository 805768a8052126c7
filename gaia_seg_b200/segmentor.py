"""DynamicEncoderDecoder -- the segmentor of the hot path.

Mirrors gaiaseg/models/segmentors/dynamic_encoder_decoder.py:8-42 on top of [EXT] mmseg EncoderDecoder
(train / inference semantics restated in-tree at gaiaseg/models/segmentors/dynamic_distiller.py:252-262,
461-521): `extract_feat`, `encode_decode`, `forward_train` (loss keys `decode.*`, `aux.*`), `_parse_losses`,
`train_step`, `forward(return_loss=...)`, `simple_test` -> list[np.ndarray int64 HxW]`, `manipulate_arch`.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as F_gs
from .backbone import BACKBONES
from .core import DynamicMixin, Registry, build_from_cfg
from .heads import HEADS

SEGMENTORS = Registry('segmentor')
NECKS = Registry('neck')


def build_backbone(cfg):
    return build_from_cfg(cfg, BACKBONES)


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


def build_neck(cfg):
    return build_from_cfg(cfg, NECKS)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    """mmseg.models.build_segmentor (tools/train_supernet.py:174, tools/test_supernet.py:184)."""
    return build_from_cfg(cfg, SEGMENTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))


class LogVars(OrderedDict):
    """log_vars of `_parse_losses`.  mmseg calls `.item()` on every entry every iteration (5 host syncs);
    here the values stay on the device in ONE stacked tensor and are only read when somebody looks."""

    def __init__(self, names, stacked):
        super().__init__((n, None) for n in names)
        self._stacked = stacked
        self._done = False

    def fresh(self):
        """A new view of the same device tensor (used after a CUDA-graph replay refreshed its values)."""
        return LogVars(list(super().keys()), self._stacked)

    def _materialise(self):
        if not self._done:
            vals = self._stacked.tolist()
            for k, v in zip(list(super().keys()), vals):
                super().__setitem__(k, v)
            self._done = True

    def __getitem__(self, k):
        self._materialise()
        return super().__getitem__(k)

    def items(self):
        self._materialise()
        return super().items()

    def values(self):
        self._materialise()
        return super().values()

    def get(self, k, default=None):
        self._materialise()
        return super().get(k, default)


class EncoderDecoder(nn.Module):
    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None):
        super().__init__()
        self.backbone = build_backbone(backbone)
        if neck is not None:
            self.neck = build_neck(neck)
        self.decode_head = build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        if auxiliary_head is not None:
            if isinstance(auxiliary_head, (list, tuple)):
                self.auxiliary_head = nn.ModuleList([build_head(c) for c in auxiliary_head])
            else:
                self.auxiliary_head = build_head(auxiliary_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.init_weights(pretrained=pretrained)

    @property
    def with_neck(self):
        return hasattr(self, 'neck') and self.neck is not None

    @property
    def with_auxiliary_head(self):
        return hasattr(self, 'auxiliary_head') and self.auxiliary_head is not None

    def init_weights(self, pretrained=None):
        self.backbone.init_weights(pretrained=pretrained)
        self.decode_head.init_weights()
        if self.with_auxiliary_head:
            heads = self.auxiliary_head if isinstance(self.auxiliary_head, nn.ModuleList) else [self.auxiliary_head]
            for h in heads:
                h.init_weights()

    def extract_feat(self, img):
        x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
        return x

    # ------------------------------------------------------------------ training
    def forward_train(self, img, img_metas, gt_semantic_seg):
        x = self.extract_feat(img)
        losses = dict()
        for k, v in self.decode_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg).items():
            losses[f'decode.{k}'] = v
        if self.with_auxiliary_head:
            if isinstance(self.auxiliary_head, nn.ModuleList):
                for idx, h in enumerate(self.auxiliary_head):
                    for k, v in h.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg).items():
                        losses[f'aux_{idx}.{k}'] = v
            else:
                for k, v in self.auxiliary_head.forward_train(x, img_metas, gt_semantic_seg, self.train_cfg).items():
                    losses[f'aux.{k}'] = v
        return losses

    log_vars_reduce = True   # GraphedTrainStep turns this off while capturing and reduces after the replay

    @staticmethod
    def _parse_losses(losses, reduce=True):
        names, vals = [], []
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                v = value.mean()
            elif isinstance(value, list):
                v = sum(_v.mean() for _v in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
            names.append(name)
            vals.append(v)
        loss = sum(v for n, v in zip(names, vals) if 'loss' in n)
        names.append('loss')
        vals.append(loss)
        stacked = torch.stack([v.detach().float() for v in vals])
        if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            ex = F_gs.PeerExchange.get(None)
            if ex:      # capturable peer-memory exchange (the same kernel as the SyncBN statistics)
                s64 = stacked.double()
                ex.all_reduce(s64)
                stacked = (s64 / dist.get_world_size()).float()
            else:
                stacked = stacked / dist.get_world_size()
                dist.all_reduce(stacked)
        return loss, LogVars(names, stacked)

    def train_step(self, data_batch, optimizer=None, **kwargs):
        losses = self(**data_batch)
        loss, log_vars = self._parse_losses(losses, self.log_vars_reduce)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data_batch['img_metas']))

    def val_step(self, data_batch, **kwargs):
        return self(**data_batch, **kwargs)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    # ------------------------------------------------------------------ inference
    def forward_test(self, imgs, img_metas, **kwargs):
        for var, name in [(imgs, 'imgs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError(f'{name} must be a list, but got {type(var)}')
        if len(imgs) != len(img_metas):
            raise ValueError(f'num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})')
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        raise NotImplementedError('multi-scale / flip augmentation test is outside the hot path')

    def encode_decode_lowres(self, img, img_metas):
        x = self.extract_feat(img)
        return self.decode_head.forward_test(x, img_metas, self.test_cfg)

    def encode_decode(self, img, img_metas):
        """Logits resized to the input size (fp32 [N, K, H, W], NHWC memory)."""
        out = self.encode_decode_lowres(img, img_metas)
        return F_gs.upsample_bilinear_f32(out, img.shape[2:])

    def simple_test(self, img, img_meta, rescale=True):
        """whole-image inference -> list of int64 HxW label maps.  soft-max is monotone, so
        resize -> softmax -> argmax is computed as ONE fused resize+argmax kernel on the low-res logits."""
        mode = (self.test_cfg or {}).get('mode', 'whole')
        if mode != 'whole':
            raise NotImplementedError("test_cfg.mode='slide' is not used by the GAIA-seg seg config "
                                      '(configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:53)')
        ori_shape = img_meta[0]['ori_shape']
        assert all(m['ori_shape'] == ori_shape for m in img_meta)
        logits = self.encode_decode_lowres(img, img_meta)
        in_size = tuple(img.shape[2:])
        out_size = tuple(ori_shape[:2]) if rescale else in_size
        if out_size != in_size:
            logits = F_gs.upsample_bilinear_f32(logits, in_size)   # reference resizes twice
        seg_pred = F_gs.upsample_argmax(logits, out_size)
        if img_meta[0].get('flip', False):
            dims = (2,) if img_meta[0].get('flip_direction', 'horizontal') == 'horizontal' else (1,)
            seg_pred = seg_pred.flip(dims=dims)
        return list(seg_pred.cpu().numpy())


@SEGMENTORS.register_module()
class DynamicEncoderDecoder(EncoderDecoder, DynamicMixin):
    search_space = {'backbone', 'decode_head', 'neck', 'auxiliary_head'}

    def manipulate_backbone(self, arch_meta):
        self.backbone.manipulate_arch(arch_meta)

    def manipulate_decode_head(self, arch_meta):
        pass

    def manipulate_neck(self, arch_meta):
        pass

    def manipulate_auxiliary_head(self, arch_meta):
        pass
