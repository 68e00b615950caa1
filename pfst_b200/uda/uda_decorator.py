"""UDADecorator — drop-in for rsiseg/models/uda/uda_decorator.py:29-103: wraps the
student segmentor and delegates the segmentor API to it. `cfg['model']` may be an
mmcv-style dict (built through pfst_b200.registry / an mmcv registry), an
nn.Module instance, or a zero-argument factory returning one."""
from __future__ import annotations

from copy import deepcopy

import torch
import torch.nn as nn

from ..registry import build_segmentor
from .log_ledger import ledger_for


def get_module(module):
    """uda_decorator.py:13-26 — unwrap a (Distributed)DataParallel-style wrapper."""
    if module is not None and hasattr(module, "module") and isinstance(module.module, nn.Module) \
            and type(module).__name__.endswith("DataParallel"):
        return module.module
    return module


def build_model(spec):
    if isinstance(spec, nn.Module):
        return deepcopy(spec)
    if callable(spec):
        return spec()
    return build_segmentor(deepcopy(spec))


def _model_attr(spec, model, key, default=None):
    if isinstance(spec, dict) and key in spec:
        return spec[key]
    return getattr(model, key, default)


class UDADecorator(nn.Module):

    def __init__(self, **cfg):
        super().__init__()
        spec = cfg['model']
        self.model = build_model(spec)
        self.train_cfg = _model_attr(spec, self.model, 'train_cfg')
        self.test_cfg = _model_attr(spec, self.model, 'test_cfg')
        if isinstance(spec, dict) and 'decode_head' in spec:
            self.num_classes = spec['decode_head']['num_classes']
        else:
            self.num_classes = getattr(self.model, 'num_classes', None)

    def get_model(self):
        return get_module(self.model)

    def extract_feat(self, img):
        return self.get_model().extract_feat(img)

    def encode_decode(self, img, img_metas):
        return self.get_model().encode_decode(img, img_metas)

    def forward_train(self, img, img_metas, gt_semantic_seg, target_img, target_img_metas, return_feat=False):
        return self.get_model().forward_train(img, img_metas, gt_semantic_seg, return_feat=return_feat)

    def inference(self, img, img_meta, rescale):
        return self.get_model().inference(img, img_meta, rescale)

    def simple_test(self, img, img_meta, rescale=True):
        return self.get_model().simple_test(img, img_meta, rescale)

    def aug_test(self, imgs, img_metas, rescale=True):
        return self.get_model().aug_test(imgs, img_metas, rescale)

    def forward(self, *args, return_loss=True, **kwargs):
        """BaseSegmentor.forward (rsiseg/models/segmentors/base.py:101-115)."""
        if return_loss:
            return self.forward_train(*args, **kwargs)
        return self.get_model().forward_test(*args, **kwargs)

    @staticmethod
    def _parse_losses(losses):
        """BaseSegmentor._parse_losses, rsiseg/models/segmentors/base.py:177-222 — same keys,
        ordering and values (`loss` = left-to-right fp32 sum of the entries whose key contains
        'loss'; log values = mean over ranks). The reference's per-variable `.item()` syncs and
        per-variable all-reduces are gone: one single-thread gather kernel per call writes the
        scalars into a persistent device ledger, ONE all-reduce per iteration reduces the row, and
        the returned log values are `LazyScalar`s that copy the ledger to the host once, when a
        value is first read (uda/log_ledger.py)."""
        dev = None
        for value in losses.values():
            t = value[0] if isinstance(value, list) and value else value
            if isinstance(t, torch.Tensor):
                dev = t.device
                break
        return ledger_for(dev if dev is not None else "cpu").parse(losses)
