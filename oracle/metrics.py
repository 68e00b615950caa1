"""mIoU evaluation — restates rsiseg/core/evaluation/metrics.py:26-86, 296-395 and
the integer confusion matrix of tools/confusion_matrix.py:46-65 /
tests/test_metrics.py:9-28."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


def areas(pred: np.ndarray, label: np.ndarray, num_classes: int, ignore_index: int,
          label_map=None, reduce_zero_label: bool = False):
    """metrics.py:56-86 -> (intersect, union, pred_area, label_area), float32 (C,)."""
    pred = torch.from_numpy(np.ascontiguousarray(pred))
    label = torch.from_numpy(np.array(label))  # copy: remapping below is in place
    if label_map:
        for old, new in label_map.items():
            label[label == old] = new
    if reduce_zero_label:
        label[label == 0] = 255
        label = label - 1
        label[label == 254] = 255
    keep = label != ignore_index
    pred, label = pred[keep], label[keep]
    hit = pred[pred == label]
    h = lambda v: torch.histc(v.float(), bins=num_classes, min=0, max=num_classes - 1)
    a_i, a_p, a_l = h(hit), h(pred), h(label)
    return a_i, a_p + a_l - a_i, a_p, a_l


def confusion(pred: np.ndarray, label: np.ndarray, num_classes: int, ignore_index: int) -> np.ndarray:
    """tests/test_metrics.py:9-28: bincount(n*gt+pred) over non-ignored pixels."""
    keep = label != ignore_index
    idx = num_classes * label[keep].astype(np.int64) + pred[keep].astype(np.int64)
    return np.bincount(idx, minlength=num_classes ** 2).reshape(num_classes, num_classes)


def f_score(precision, recall, beta=1):
    # metrics.py:9-23
    return (1 + beta ** 2) * (precision * recall) / ((beta ** 2 * precision) + recall)


def metrics_from_areas(a_i, a_u, a_p, a_l, metrics=("mIoU",), nan_to_num=None, beta=1):
    """metrics.py:333-395."""
    if isinstance(metrics, str):
        metrics = [metrics]
    if not set(metrics) <= {"mIoU", "mDice", "mFscore"}:
        raise KeyError(f"metrics {metrics} is not supported")
    out = OrderedDict({"aAcc": a_i.sum() / a_l.sum()})
    for m in metrics:
        if m == "mIoU":
            out["IoU"] = a_i / a_u
            out["Acc"] = a_i / a_l
        elif m == "mDice":
            out["Dice"] = 2 * a_i / (a_p + a_l)
            out["Acc"] = a_i / a_l
        else:
            prec, rec = a_i / a_p, a_i / a_l
            out["Fscore"] = torch.tensor([f_score(x[0], x[1], beta) for x in zip(prec, rec)])
            out["Precision"] = prec
            out["Recall"] = rec
    out = {k: v.numpy() for k, v in out.items()}
    if nan_to_num is not None:
        out = OrderedDict({k: np.nan_to_num(v, nan=nan_to_num) for k, v in out.items()})
    return out


def pre_eval_sum(per_image):
    """metrics.py:315-323: python `sum` of the per-image float32 vectors, in order."""
    cols = tuple(zip(*per_image))
    return tuple(sum(c) for c in cols)


def total_areas(preds, labels, num_classes, ignore_index, label_map=None, reduce_zero_label=False):
    """metrics.py:89-129: float64 running totals."""
    tot = [torch.zeros((num_classes,), dtype=torch.float64) for _ in range(4)]
    for p, l in zip(preds, labels):
        for t, a in zip(tot, areas(p, l, num_classes, ignore_index, label_map, reduce_zero_label)):
            t += a
    return tuple(tot)


def eval_metrics(preds, labels, num_classes, ignore_index, metrics=("mIoU",), nan_to_num=None,
                 label_map=None, reduce_zero_label=False, beta=1):
    """metrics.py:257-293."""
    return metrics_from_areas(*total_areas(preds, labels, num_classes, ignore_index, label_map,
                                           reduce_zero_label), metrics, nan_to_num, beta)


def seg_argmax(seg_logits: torch.Tensor) -> torch.Tensor:
    """EncoderDecoder.inference + simple_test, rsiseg/models/segmentors/encoder_decoder.py:311
    (`output = F.softmax(seg_logit, dim=1)`) and :332 (`seg_pred = seg_logit.argmax(dim=1)`;
    `seg_logit` there IS the softmax output). (N,C,H,W) -> int64 (N,H,W)."""
    return torch.nn.functional.softmax(seg_logits, dim=1).argmax(dim=1)


def pre_eval(seg_logits: torch.Tensor, gt_seg_maps, num_classes: int, ignore_index: int,
             label_map=None, reduce_zero_label: bool = False):
    """simple_test (encoder_decoder.py:329-338: arg-max, `.cpu().numpy()`, `list(seg_pred)`) followed
    by dataset.pre_eval (rsiseg/datasets/custom.py:644-682: one intersect_and_union per image)."""
    preds = list(seg_argmax(seg_logits).cpu().numpy())
    return [areas(p, np.asarray(g), num_classes, ignore_index, label_map, reduce_zero_label)
            for p, g in zip(preds, gt_seg_maps)]


def slide_inference(encode_decode, img: torch.Tensor, img_meta, rescale: bool, crop_size, stride, num_classes: int,
                    align_corners: bool = False) -> torch.Tensor:
    """EncoderDecoder.slide_inference, rsiseg/models/segmentors/encoder_decoder.py:220-263, statement for
    statement (`encode_decode(img, img_meta) -> (seg_logit, states)` is the network pass)."""
    F = torch.nn.functional
    h_stride, w_stride = stride
    h_crop, w_crop = crop_size
    batch_size, _, h_img, w_img = img.size()
    h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
    preds = img.new_zeros((batch_size, num_classes, h_img, w_img))
    count_mat = img.new_zeros((batch_size, 1, h_img, w_img))
    for h_idx in range(h_grids):
        for w_idx in range(w_grids):
            y1 = h_idx * h_stride
            x1 = w_idx * w_stride
            y2 = min(y1 + h_crop, h_img)
            x2 = min(x1 + w_crop, w_img)
            y1 = max(y2 - h_crop, 0)
            x1 = max(x2 - w_crop, 0)
            crop_seg_logit, _ = encode_decode(img[:, :, y1:y2, x1:x2], img_meta)
            preds += F.pad(crop_seg_logit, (int(x1), int(preds.shape[3] - x2), int(y1), int(preds.shape[2] - y2)))
            count_mat[:, :, y1:y2, x1:x2] += 1
    assert (count_mat == 0).sum() == 0
    preds = preds / count_mat
    if rescale:
        preds = F.interpolate(preds, size=tuple(img_meta[0]['ori_shape'][:2]), mode='bilinear',
                              align_corners=align_corners)
    return preds


def inference(encode_decode, img, img_meta, rescale: bool, mode: str, crop_size=None, stride=None,
              num_classes: int = 0, align_corners: bool = False):
    """EncoderDecoder.inference (+ whole_inference), encoder_decoder.py:265-324 -> (soft-max output, states)."""
    F = torch.nn.functional
    assert mode in ['slide', 'whole']
    ori_shape = img_meta[0]['ori_shape']
    assert all(_['ori_shape'] == ori_shape for _ in img_meta)
    if mode == 'slide':
        seg_logit = slide_inference(encode_decode, img, img_meta, rescale, crop_size, stride, num_classes,
                                    align_corners)
        states = {}
    else:
        seg_logit, states = encode_decode(img, img_meta)
        if rescale:
            seg_logit = F.interpolate(seg_logit, size=tuple(ori_shape[:2]), mode='bilinear',
                                      align_corners=align_corners)
    output = F.softmax(seg_logit, dim=1)
    if img_meta[0]['flip']:
        flip_direction = img_meta[0]['flip_direction']
        if type(flip_direction) != list:
            flip_direction = [flip_direction]
        for direction_ in flip_direction:
            assert direction_ in ['horizontal', 'vertical']
            if direction_ == 'horizontal':
                output = output.flip(dims=(3, ))
            elif direction_ == 'vertical':
                output = output.flip(dims=(2, ))
    return output, states


def aug_test(encode_decode, imgs, img_metas, rescale: bool, mode: str, crop_size=None, stride=None,
             num_classes: int = 0, align_corners: bool = False):
    """EncoderDecoder.aug_test, encoder_decoder.py:355-373 -> (list of int64 numpy maps, {})."""
    assert rescale
    seg_logit, _ = inference(encode_decode, imgs[0], img_metas[0], rescale, mode, crop_size, stride, num_classes,
                             align_corners)
    for i in range(1, len(imgs)):
        cur_seg_logit, _ = inference(encode_decode, imgs[i], img_metas[i], rescale, mode, crop_size, stride,
                                     num_classes, align_corners)
        seg_logit += cur_seg_logit
    seg_logit /= len(imgs)
    seg_pred = seg_logit.argmax(dim=1)
    return list(seg_pred.cpu().numpy()), {}
