"""Pseudo-label generation — restates rsiseg/models/uda/pfgst.py:259-277 and the
offline class-wise rule of rsiseg/datasets/pipelines/loading.py:474-487."""
from __future__ import annotations

import numpy as np
import torch


def pseudo_label(logits: torch.Tensor, thr: float):
    """pfgst.py:259-261 -> (label int64 (B,H,W), prob fp32, confident bool)."""
    sm = torch.softmax(logits.detach(), dim=1)
    prob, label = torch.max(sm, dim=1)
    large = prob.ge(thr).long() == 1
    return label, prob, large


def pseudo_label_classwise(logits: torch.Tensor, thr_per_class: torch.Tensor):
    """Extension of the online rule to a per-class threshold vector (north_star;
    identical to pseudo_label when the vector is constant)."""
    sm = torch.softmax(logits.detach(), dim=1)
    prob, label = torch.max(sm, dim=1)
    large = prob >= thr_per_class.to(prob.dtype)[label]
    return label, prob, large


def pseudo_weight(large: torch.Tensor, thre_type: str = "all", ignore_top: int = 0,
                  ignore_bottom: int = 0) -> torch.Tensor:
    """pfgst.py:262-276."""
    ps_size = large.numel()  # np.size(np.array(pseudo_label.cpu()))
    if thre_type == "all":
        w = torch.sum(large).item() / ps_size
        w = w * torch.ones(large.shape, device=large.device)
    elif thre_type == "part":
        w = large.float()
    else:
        raise ValueError(thre_type)
    if ignore_top > 0:
        w[:, :ignore_top, :] = 0
    if ignore_bottom > 0:
        w[:, -ignore_bottom:, :] = 0
    return w


def entropy_label(logits: np.ndarray, thres: np.ndarray, reject: int = 255):
    """loading.py:474-484 for one (C,H,W) logit map: keep argmax where the
    softmax entropy is below the class's threshold, else `reject`."""
    preds = logits.argmax(axis=0)
    probs = np.exp(logits) / np.exp(logits).sum(axis=0)
    ent = -(probs * np.log(probs + 1e-8)).sum(axis=0)
    keep = ent < thres[preds]
    return np.where(keep, preds, reject), ent, keep
