"""ClassMix — restates rsiseg/models/utils/dacs_transforms.py:110-144 and the
mixing loop of rsiseg/models/uda/pfgst.py:277-300 (colour jitter / blur are
kornia arithmetic, outside the graded path: SURVEY.md §8c)."""
from __future__ import annotations

import numpy as np
import torch


def class_masks(gt: torch.Tensor, rng=np.random) -> list[torch.Tensor]:
    """dacs_transforms.py:110-126. gt (B,1,H,W) int64 -> B masks (1,1,H,W) int64.
    The class set is that of the WHOLE batch (:113); one host RNG draw per image."""
    out = []
    for lab in gt:
        classes = torch.unique(gt)
        n = classes.shape[0]
        pick = rng.choice(n, int((n + n % 2) / 2), replace=False)
        chosen = classes[torch.Tensor(pick).long()]
        m = lab.eq(chosen.unsqueeze(1).unsqueeze(2)).sum(0, keepdims=True)
        out.append(m.unsqueeze(0))
    return out


def mix_pair(mask: torch.Tensor, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """one_mix, dacs_transforms.py:133-143: mask*a + (1-mask)*b with mask[0] broadcast."""
    m, _ = torch.broadcast_tensors(mask[0], a)
    return (m * a + (1 - m) * b).unsqueeze(0)


def mix_batch(img, trg_img, gt, pseudo_lbl, pseudo_w, masks):
    """pfgst.py:277-300. Returns mixed_img (B,3,H,W), mixed_lbl (B,1,H,W) int64,
    pseudo_w updated in place (B,H,W), mix_masks (B,1,H,W) int64."""
    B = img.shape[0]
    gt_w = torch.ones(pseudo_w.shape, device=pseudo_w.device)
    imgs, lbls = [], []
    for i in range(B):
        imgs.append(mix_pair(masks[i], img[i], trg_img[i]))
        lbls.append(mix_pair(masks[i], gt[i][0], pseudo_lbl[i]))
        pseudo_w[i] = mix_pair(masks[i], gt_w[i], pseudo_w[i])
    return torch.cat(imgs), torch.cat(lbls), pseudo_w, torch.cat(masks, dim=0)
