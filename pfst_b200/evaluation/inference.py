"""Test-time inference arithmetic of the reference's EncoderDecoder, kept on the GPU.

Mirrors rsiseg/models/segmentors/encoder_decoder.py:
  slide_inference :220-263   sliding windows with overlap: `preds += F.pad(crop_logit)`, count matrix, division
  inference       :283-324   slide / whole, soft-max, horizontal / vertical flip
  simple_test     :326-353   arg-max, `.cpu().numpy()`, one map per image
The network pass itself (`encode_decode`) is the caller's: every function takes it as a callable
`encode_decode(img, img_meta) -> (seg_logit, states)` — bind `model.encode_decode` — so a maintainer's
EncoderDecoder subclass overrides its three methods with one-line calls.

What changes: a window adds into its own region only (`pfst_slide_add`), the count matrix is two
vectors (the windows are a product of row and column intervals), division and flips are one pass
(`pfst_slide_finalize`), and soft-max + arg-max (+ the confusion matrix) is the fused evaluation kernel.
Results are bit-identical to the reference methods (tests/test_gpu_slide.py).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Callable

import numpy as np
import torch
import torch.nn.functional as F

from .. import _lib, ops
from .._lib import PfstError
from .metrics import seg_argmax


def _cfg(test_cfg, key):
    return test_cfg[key] if isinstance(test_cfg, dict) else getattr(test_cfg, key)


def window_grid(h_img: int, w_img: int, crop_size, stride):
    """The (y1, y2, x1, x2) windows of slide_inference in its loop order (encoder_decoder.py:227-241)
    and the per-row / per-column window counts: count_mat[y, x] = cnt_y[y] * cnt_x[x]."""
    h_stride, w_stride = stride
    h_crop, w_crop = crop_size
    h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
    rows, cols = [], []
    for h_idx in range(h_grids):
        y2 = min(h_idx * h_stride + h_crop, h_img)
        rows.append((max(y2 - h_crop, 0), y2))
    for w_idx in range(w_grids):
        x2 = min(w_idx * w_stride + w_crop, w_img)
        cols.append((max(x2 - w_crop, 0), x2))
    cnt_y, cnt_x = np.zeros(h_img, dtype=np.float32), np.zeros(w_img, dtype=np.float32)
    for y1, y2 in rows:
        cnt_y[y1:y2] += 1
    for x1, x2 in cols:
        cnt_x[x1:x2] += 1
    return [(y1, y2, x1, x2) for (y1, y2) in rows for (x1, x2) in cols], cnt_y, cnt_x


def _resize(x, size, align_corners):
    # rsiseg/ops/wrappers.py resize == F.interpolate (the reference's own library call)
    return F.interpolate(x, size=tuple(int(v) for v in size), mode='bilinear', align_corners=align_corners)


def slide_logits(encode_decode: Callable, img: torch.Tensor, img_meta, crop_size, stride, num_classes: int,
                 flip_h: bool = False, flip_v: bool = False) -> torch.Tensor:
    """Window-averaged logits (B, num_classes, H, W), optionally flipped: slide_inference without the
    rescale, flips of `inference` folded into the same pass."""
    if not img.is_cuda:
        raise PfstError("slide inference runs on CUDA tensors only (no CPU fallback)")
    B, _, h_img, w_img = img.shape
    windows, cnt_y, cnt_x = window_grid(h_img, w_img, crop_size, stride)
    assert (cnt_y == 0).sum() == 0 and (cnt_x == 0).sum() == 0            # encoder_decoder.py:250
    preds = torch.zeros((B, num_classes, h_img, w_img), dtype=torch.float32, device=img.device)
    for y1, y2, x1, x2 in windows:
        crop_seg_logit, _ = encode_decode(img[:, :, y1:y2, x1:x2], img_meta)
        crop_seg_logit = crop_seg_logit.float().contiguous()
        if tuple(crop_seg_logit.shape) != (B, num_classes, y2 - y1, x2 - x1):
            raise PfstError(f"encode_decode returned {tuple(crop_seg_logit.shape)} for a "
                            f"{(B, num_classes, y2 - y1, x2 - x1)} window")
        _lib.call("pfst_slide_add", preds.data_ptr(), crop_seg_logit.data_ptr(), B, num_classes, h_img, w_img,
                  y1, x1, y2 - y1, x2 - x1, ops._stream())
    dy = torch.from_numpy(cnt_y).to(img.device, non_blocking=True)
    dx = torch.from_numpy(cnt_x).to(img.device, non_blocking=True)
    out = torch.empty_like(preds) if (flip_h or flip_v) else preds
    _lib.call("pfst_slide_finalize", preds.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, num_classes, h_img, w_img,
              int(flip_h), int(flip_v), out.data_ptr(), ops._stream())
    return out


def slide_inference(encode_decode: Callable, img, img_meta, rescale: bool, *, crop_size, stride, num_classes: int,
                    align_corners: bool = False) -> torch.Tensor:
    """EncoderDecoder.slide_inference (encoder_decoder.py:220-263) -> preds (B, C, H, W) raw logits."""
    preds = slide_logits(encode_decode, img, img_meta, crop_size, stride, num_classes)
    if rescale:
        preds = _resize(preds, img_meta[0]['ori_shape'][:2], align_corners)
    return preds


def _flips(img_meta):
    flip_h = flip_v = False
    if img_meta[0]['flip']:
        direction = img_meta[0]['flip_direction']
        for d in (direction if type(direction) == list else [direction]):
            assert d in ['horizontal', 'vertical']
            if d == 'horizontal':
                flip_h = not flip_h
            else:
                flip_v = not flip_v
    return flip_h, flip_v


def inference_logits(encode_decode: Callable, img, img_meta, rescale: bool, test_cfg, num_classes: int,
                     align_corners: bool = False):
    """`inference` (encoder_decoder.py:283-324) up to, not including, the soft-max: logits that the fused
    evaluation kernels consume (soft-max is per pixel, so flipping before it is the same map). -> (logits, states)."""
    mode = _cfg(test_cfg, 'mode')
    assert mode in ['slide', 'whole']
    ori_shape = img_meta[0]['ori_shape']
    assert all(_['ori_shape'] == ori_shape for _ in img_meta)
    flip_h, flip_v = _flips(img_meta)
    states = {}
    if mode == 'slide':
        same = (not rescale) or tuple(ori_shape[:2]) == tuple(img.shape[2:])
        logits = slide_logits(encode_decode, img, img_meta, _cfg(test_cfg, 'crop_size'), _cfg(test_cfg, 'stride'),
                              num_classes, flip_h and same, flip_v and same)
        if not same:
            logits = _resize(logits, ori_shape[:2], align_corners)
            flip_h_late, flip_v_late = flip_h, flip_v
        else:
            flip_h_late = flip_v_late = False
    else:
        logits, states = encode_decode(img, img_meta)
        logits = logits.float()
        if rescale:
            logits = _resize(logits, ori_shape[:2], align_corners)
        flip_h_late, flip_v_late = flip_h, flip_v
    if flip_h_late or flip_v_late:
        logits = logits.contiguous()
        out = torch.empty_like(logits)
        Bn, Cn, Hn, Wn = logits.shape
        _lib.call("pfst_slide_finalize", logits.data_ptr(), None, None, Bn, Cn, Hn, Wn, int(flip_h_late),
                  int(flip_v_late), out.data_ptr(), ops._stream())
        logits = out
    return logits.contiguous(), states


def inference(encode_decode: Callable, img, img_meta, rescale: bool, test_cfg, num_classes: int,
              align_corners: bool = False):
    """Drop-in for EncoderDecoder.inference: -> (soft-max output, states). Compatibility entry point — the
    evaluation path proper never materialises the soft-max (`simple_test` / `pre_eval_logits` below)."""
    logits, states = inference_logits(encode_decode, img, img_meta, rescale, test_cfg, num_classes, align_corners)
    return F.softmax(logits, dim=1), states


def simple_test(encode_decode: Callable, img, img_meta, rescale: bool = True, *, test_cfg, num_classes: int,
                align_corners: bool = False):
    """EncoderDecoder.simple_test (encoder_decoder.py:326-353): list of per-image int64 arg-max maps
    (numpy) and the per-image states list."""
    logits, states = inference_logits(encode_decode, img, img_meta, rescale, test_cfg, num_classes, align_corners)
    seg_pred = list(seg_argmax(logits).cpu().numpy())
    state_list = []
    for idx in range(len(seg_pred)):
        cur_state = {}
        if 'feats' in states:
            cur_state['feats'] = [x[idx].cpu() for x in states['feats']]
        if 'decoded_features' in states:
            cur_state['decoded_feats'] = states['decoded_features'].cpu()
        if 'seg_logits' in states:
            cur_state['seg_logits'] = states['seg_logits'][idx].cpu()
        state_list.append(cur_state)
    return seg_pred, state_list


def aug_test(encode_decode: Callable, imgs, img_metas, rescale: bool = True, *, test_cfg, num_classes: int,
             align_corners: bool = False):
    """EncoderDecoder.aug_test (encoder_decoder.py:355-373): the soft-max outputs of the augmented
    inferences are summed, divided by their number and arg-maxed. Here every augmentation's logits go
    through `pfst_softmax_accum` (soft-max and accumulation in one pass: the per-augmentation soft-max
    tensors never exist) and `pfst_div_argmax`. -> (list of int64 numpy maps, {})."""
    assert rescale
    acc = None
    for i in range(len(imgs)):
        logits, _ = inference_logits(encode_decode, imgs[i], img_metas[i], rescale, test_cfg, num_classes,
                                     align_corners)
        if acc is None:
            acc = torch.empty_like(logits)
        elif acc.shape != logits.shape:
            raise PfstError(f"aug_test: augmentation {i} gives {tuple(logits.shape)}, the first {tuple(acc.shape)}")
        B, Cn, H, W = logits.shape
        _lib.call("pfst_softmax_accum", logits.data_ptr(), acc.data_ptr(), B, Cn, H * W, int(i == 0), ops._stream())
    B, Cn, H, W = acc.shape
    pred = torch.empty((B, H, W), dtype=torch.int64, device=acc.device)
    _lib.call("pfst_div_argmax", acc.data_ptr(), B, Cn, H * W, float(len(imgs)), pred.data_ptr(), ops._stream())
    return list(pred.cpu().numpy()), {}


def make_test_cfg(mode='whole', crop_size=None, stride=None):
    return SimpleNamespace(mode=mode, crop_size=crop_size, stride=stride)
