"""mIoU evaluation — drop-in for rsiseg/core/evaluation/metrics.py.

Same public functions and return types as the reference (`intersect_and_union`
returns four float32 CPU tensors of C entries, `eval_metrics` an OrderedDict of
numpy arrays, ...). The per-pixel work (the reference's three CPU `torch.histc`
calls per image, metrics.py:74-86) runs in csrc/confusion.cu as one integer
confusion-matrix pass; the handful of per-class divisions stays on the host.

Extra, for sweeps that keep predictions on the GPU:
`confusion_matrix`, `intersect_and_union_batch`, `ConfusionMeter` (+ NCCL all-reduce).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import numpy as np
import torch

from .. import ops
from .._lib import PfstError


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise PfstError("pfst_b200.evaluation needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device_map(x, what: str) -> torch.Tensor:
    if isinstance(x, str):
        if what == "pred":
            x = np.load(x)
        else:
            from PIL import Image  # the reference uses mmcv.imread(flag='unchanged', backend='pillow')
            x = np.array(Image.open(x))
    if isinstance(x, np.ndarray):
        if x.dtype not in (np.uint8, np.int32, np.int64):
            x = x.astype(np.int64)
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what}: expected ndarray, str or torch.Tensor")
    if x.dtype not in (torch.uint8, torch.int32, torch.int64):
        x = x.long()
    if not x.is_cuda:
        x = x.to(_device(), non_blocking=True)
    return x.contiguous()


def _label_lut(label_map, device) -> Optional[torch.Tensor]:
    """metrics.py:66-68 applies `label[label == old] = new` sequentially, in dict order."""
    if not label_map:
        return None
    lut = np.arange(256, dtype=np.int64)
    for old, new in label_map.items():
        if not (0 <= new <= 255):
            raise PfstError("label_map target outside [0,255]")
        lut[lut == old] = new
    return torch.from_numpy(lut.astype(np.uint8)).to(device)


def _areas_from_conf(conf: torch.Tensor, C: int):
    """(…,C+1,C+1) int64 -> intersect, union, pred, label (…,C) int64 (histc semantics)."""
    inter = torch.diagonal(conf, dim1=-2, dim2=-1)[..., :C]
    label_area = conf.sum(-1)[..., :C]
    pred_area = conf.sum(-2)[..., :C]
    return inter, pred_area + label_area - inter, pred_area, label_area


def confusion_matrix(pred_label, label, num_classes: int, ignore_index: int = 255, label_map=None,
                     reduce_zero_label: bool = False) -> torch.Tensor:
    """Integer confusion matrix (rows = ground truth, cols = prediction) as
    tools/confusion_matrix.py:46-65 / tests/test_metrics.py:9-28 -> int64 (C,C) on device."""
    p, l = _to_device_map(pred_label, "pred"), _to_device_map(label, "label")
    conf = ops.confusion_accum(p, l, num_classes, ignore_index, reduce_zero_label, _label_lut(label_map, p.device))
    return conf[0, :num_classes, :num_classes]


def intersect_and_union(pred_label, label, num_classes, ignore_index, label_map=dict(),
                        reduce_zero_label=False):
    """metrics.py:26-86 — four float32 (C,) CPU tensors for ONE image."""
    p, l = _to_device_map(pred_label, "pred"), _to_device_map(label, "label")
    if p.shape != l.shape:
        raise ValueError("pred_label and label differ in shape")
    conf = ops.confusion_accum(p.reshape(1, -1), l.reshape(1, -1), num_classes, ignore_index,
                               reduce_zero_label, _label_lut(label_map, p.device))
    out = [a[0].cpu().to(torch.float32) for a in _areas_from_conf(conf, num_classes)]
    return out[0], out[1], out[2], out[3]


def intersect_and_union_batch(preds, labels, num_classes, ignore_index, label_map=dict(),
                              reduce_zero_label=False) -> torch.Tensor:
    """N maps in one launch -> int64 (N,4,C) on device (intersect, union, pred, label)."""
    p, l = _to_device_map(preds, "pred"), _to_device_map(labels, "label")
    conf = ops.confusion_accum(p, l, num_classes, ignore_index, reduce_zero_label,
                               _label_lut(label_map, p.device), per_image=True)
    return torch.stack(_areas_from_conf(conf, num_classes), dim=1)


def seg_argmax(seg_logits: torch.Tensor, dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """`F.softmax(seg_logit, dim=1).argmax(dim=1)` of EncoderDecoder.inference/simple_test
    (encoder_decoder.py:311,332) in one pass, kept on the device. (N,C,H,W) fp32 -> (N,H,W)."""
    return ops.argmax_confusion(seg_logits, None, return_pred=dtype)[1]


def pre_eval_logits(seg_logits: torch.Tensor, gt_seg_maps, num_classes: int, ignore_index: int,
                    label_map=dict(), reduce_zero_label: bool = False):
    """simple_test + dataset.pre_eval for logits that are still on the GPU
    (encoder_decoder.py:329-338 + custom.py:644-682): the arg-max map never exists in memory and
    never crosses PCIe; one launch and one (N,(C+1)^2) int64 read for the whole batch.
    gt_seg_maps: (N,H,W) tensor/array or a list of N (H,W) maps. -> list of N tuples of four
    float32 (C,) CPU tensors, exactly what `pre_eval` returns per image."""
    if isinstance(gt_seg_maps, (list, tuple)):
        gt_seg_maps = torch.stack([_to_device_map(g, "label") for g in gt_seg_maps])
    l = _to_device_map(gt_seg_maps, "label")
    conf, _ = ops.argmax_confusion(seg_logits, l, num_classes, ignore_index, reduce_zero_label,
                                   _label_lut(label_map, seg_logits.device), per_image=True)
    areas = [a.cpu().to(torch.float32) for a in _areas_from_conf(conf, num_classes)]
    return [tuple(a[i] for a in areas) for i in range(conf.shape[0])]


def total_intersect_and_union(results, gt_seg_maps, num_classes, ignore_index, label_map=dict(),
                              reduce_zero_label=False):
    """metrics.py:89-129 — float64 totals over a list of maps."""
    tot = None
    lut = None
    for res, gt in zip(results, gt_seg_maps):
        p, l = _to_device_map(res, "pred"), _to_device_map(gt, "label")
        if lut is None:
            lut = _label_lut(label_map, p.device)
        tot = ops.confusion_accum(p.reshape(1, -1), l.reshape(1, -1), num_classes, ignore_index,
                                  reduce_zero_label, lut, out=tot)
    if tot is None:
        z = torch.zeros((num_classes,), dtype=torch.float64)
        return z, z.clone(), z.clone(), z.clone()
    a = [x[0].cpu().to(torch.float64) for x in _areas_from_conf(tot, num_classes)]
    return a[0], a[1], a[2], a[3]


def f_score(precision, recall, beta=1):
    """metrics.py:9-23."""
    return (1 + beta ** 2) * (precision * recall) / ((beta ** 2 * precision) + recall)


def total_area_to_metrics(total_area_intersect, total_area_union, total_area_pred_label,
                          total_area_label, metrics=['mIoU'], nan_to_num=None, beta=1):
    """metrics.py:333-395 (host arithmetic on C-vectors)."""
    if isinstance(metrics, str):
        metrics = [metrics]
    allowed = ['mIoU', 'mDice', 'mFscore']
    if not set(metrics).issubset(set(allowed)):
        raise KeyError('metrics {} is not supported'.format(metrics))
    ret = OrderedDict({'aAcc': total_area_intersect.sum() / total_area_label.sum()})
    for metric in metrics:
        if metric == 'mIoU':
            ret['IoU'] = total_area_intersect / total_area_union
            ret['Acc'] = total_area_intersect / total_area_label
        elif metric == 'mDice':
            ret['Dice'] = 2 * total_area_intersect / (total_area_pred_label + total_area_label)
            ret['Acc'] = total_area_intersect / total_area_label
        elif metric == 'mFscore':
            precision = total_area_intersect / total_area_pred_label
            recall = total_area_intersect / total_area_label
            ret['Fscore'] = torch.tensor([f_score(x[0], x[1], beta) for x in zip(precision, recall)])
            ret['Precision'] = precision
            ret['Recall'] = recall
    ret = {k: v.numpy() for k, v in ret.items()}
    if nan_to_num is not None:
        ret = OrderedDict({k: np.nan_to_num(v, nan=nan_to_num) for k, v in ret.items()})
    return ret


def pre_eval_to_metrics(pre_eval_results, metrics=['mIoU'], nan_to_num=None, beta=1):
    """metrics.py:296-330 — python `sum` of the per-image float32 vectors, in list order
    (kept in float32 exactly like the reference, including its rounding beyond 2^24)."""
    cols = tuple(zip(*pre_eval_results))
    assert len(cols) == 4
    return total_area_to_metrics(sum(cols[0]), sum(cols[1]), sum(cols[2]), sum(cols[3]), metrics,
                                 nan_to_num, beta)


def eval_metrics(results, gt_seg_maps, num_classes, ignore_index, metrics=['mIoU'], nan_to_num=None,
                 label_map=dict(), reduce_zero_label=False, beta=1):
    """metrics.py:257-293."""
    tot = total_intersect_and_union(results, gt_seg_maps, num_classes, ignore_index, label_map,
                                    reduce_zero_label)
    return total_area_to_metrics(*tot, metrics, nan_to_num, beta)


def mean_iou(results, gt_seg_maps, num_classes, ignore_index, nan_to_num=None, label_map=dict(),
             reduce_zero_label=False):
    return eval_metrics(results, gt_seg_maps, num_classes, ignore_index, ['mIoU'], nan_to_num,
                        label_map, reduce_zero_label)


def mean_dice(results, gt_seg_maps, num_classes, ignore_index, nan_to_num=None, label_map=dict(),
              reduce_zero_label=False):
    return eval_metrics(results, gt_seg_maps, num_classes, ignore_index, ['mDice'], nan_to_num,
                        label_map, reduce_zero_label)


def mean_fscore(results, gt_seg_maps, num_classes, ignore_index, nan_to_num=None, label_map=dict(),
                reduce_zero_label=False, beta=1):
    return eval_metrics(results, gt_seg_maps, num_classes, ignore_index, ['mFscore'], nan_to_num,
                        label_map, reduce_zero_label, beta)


def shard_range(n_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of an evaluation sweep for `rank` (sizes differ by at most one);
    the confusion matrix is the only quantity reduced across ranks afterwards."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ConfusionMeter:
    """Running (C+1)x(C+1) int64 confusion matrix kept on the device; `all_reduce()`
    sums it across ranks (NCCL) — the only cross-image quantity of evaluation."""

    def __init__(self, num_classes: int, ignore_index: int = 255, reduce_zero_label: bool = False,
                 label_map=None, device: Optional[torch.device] = None):
        self.C = int(num_classes)
        self.ignore_index = ignore_index
        self.reduce_zero_label = reduce_zero_label
        self.device = device or _device()
        self._lut = _label_lut(label_map, self.device)
        self.conf = torch.zeros((1, self.C + 1, self.C + 1), dtype=torch.int64, device=self.device)

    def update(self, preds: torch.Tensor, labels: torch.Tensor) -> None:
        ops.confusion_accum(preds, labels, self.C, self.ignore_index, self.reduce_zero_label, self._lut,
                            out=self.conf)

    def update_logits(self, seg_logits: torch.Tensor, labels: torch.Tensor) -> None:
        """Same as `update(softmax(seg_logits,1).argmax(1), labels)` without materialising the
        prediction map (fused arg-max + confusion kernel)."""
        ops.argmax_confusion(seg_logits, labels, self.C, self.ignore_index, self.reduce_zero_label,
                             self._lut, out=self.conf)

    def all_reduce(self, group=None) -> None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.conf, op=dist.ReduceOp.SUM, group=group)

    def matrix(self) -> torch.Tensor:
        return self.conf[0, :self.C, :self.C]

    def areas(self):
        return tuple(a[0].cpu().to(torch.float64) for a in _areas_from_conf(self.conf, self.C))

    def metrics(self, metrics=['mIoU'], nan_to_num=None, beta=1):
        return total_area_to_metrics(*self.areas(), metrics, nan_to_num, beta)
