"""CPU oracle (TEST INFRASTRUCTURE) for the offline class-wise pseudo-label thresholds:
PseudoLabelingHookV4._cal_threshold, rsiseg/core/hook/pseudo_labeling_hookv4.py:173-205,
restated operator for operator. The random subset comes from the caller's numpy stream exactly
as in the reference (`np.random.permutation(num_samples)[:int(num_samples * sample_ratio) - 1]`)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def cal_threshold(seg_logits: torch.Tensor, sample_ratio: float, cls_thre_ratios, rng=np.random):
    B, num_classes, H, W = seg_logits.shape
    seg_logits = seg_logits.permute(0, 2, 3, 1).contiguous().view(-1, num_classes)      # :176
    num_samples, _ = seg_logits.shape
    idx = rng.permutation(num_samples)[:int(num_samples * sample_ratio) - 1]            # :180
    seg_logits = seg_logits[idx, :]
    prob_maps = F.softmax(seg_logits, dim=1)                                            # :185
    pred_maps = prob_maps.argmax(dim=1)
    ent_maps = (- prob_maps * torch.log(prob_maps)).sum(dim=1)
    thre_map = {f'thre@{r}': [] for r in cls_thre_ratios}                               # :189-191
    for cls in range(num_classes):
        if (pred_maps == cls).sum() == 0:
            for r in cls_thre_ratios:
                thre_map[f'thre@{r}'].append(0)                                         # :196-198
        else:
            sorted_map = np.sort(ent_maps[pred_maps == cls].reshape(-1))                # :200
            for r in cls_thre_ratios:
                thre_map[f'thre@{r}'].append(sorted_map[int(len(sorted_map) * r)])      # :202-204
    return thre_map
