"""CPU oracle (TEST INFRASTRUCTURE) for the rest of the offline class-wise pseudo-labelling path
(SURVEY.md §8f rank 4), restated operator for operator:

  cal_loc_dis            PseudoLabelingHookV4._cal_loc_dis   rsiseg/core/hook/pseudo_labeling_hookv4.py:208-230
  cal_sigmas             PseudoLabelingHookV4._cal_sigmas    rsiseg/core/hook/pseudo_labeling_hookv4.py:232-277
  loader_pseudo_labels   LoadAnnotationsPseudoLabelsV2.__call__ (the label rule)
                                                             rsiseg/datasets/pipelines/loading.py:474-487

Random subsets come from the caller's numpy stream exactly as in the reference. Pinned against the
reference methods compiled from their source (tests/test_oracle_pins.py) and by
tests/golden/offline_labels.npz."""
from __future__ import annotations

import numpy as np
import torch


def cal_loc_dis(feats, kernel_size: int, dilations) -> dict:
    """feats: list of (C,H,W) tensors (one per level) -> {'level{l}_dila@{d}': (1,H,W,k*k) fp32}:
    squared L2 distance between every pixel and its k x k dilated neighbours, zero padding."""
    if type(dilations) != list:
        dilations = [dilations]
    loc_dis = dict()
    for level, feat in enumerate(feats):
        C, H, W = feat.shape
        feat = feat.unsqueeze(0)
        for dila in dilations:
            unfold = torch.nn.Unfold(kernel_size=kernel_size, padding=kernel_size // 2 * dila, dilation=dila)   # :220
            unf = unfold(feat).view(1, -1, kernel_size ** 2, H, W).permute(0, 1, 3, 4, 2)                       # :224
            loc_dis[f'level{level}_dila@{dila}'] = ((unf - feat.unsqueeze(4)) ** 2).sum(dim=1)                  # :226
    return loc_dis


def cal_sigmas(loc_dis_list, feat_level, dilations, mean_sims, sample_ratio: float, rng=np.random) -> dict:
    """Bisection of sigma on [0, 1000] until mean(exp(-d / sigma^2)) over a random subset of the
    pixels reaches `mean_sim` (:262-275); returns the left end like the reference."""
    if type(dilations) != list:
        dilations = [dilations]
    if type(mean_sims) != list:
        mean_sims = [mean_sims]
    loc_dis_tensor = dict()
    for level in feat_level:
        for dila in dilations:
            cur = torch.cat([ld[f'level{level}_dila@{dila}'] for ld in loc_dis_list], dim=0)                    # :250
            B, H, W, C = cur.shape
            cur = cur.view(-1, C)
            num_samples = cur.shape[0]
            idx = rng.permutation(num_samples)[:int(num_samples * sample_ratio) - 1]                            # :255
            loc_dis_tensor[f'level{level}_dila@{dila}'] = cur[idx, :]
    sigmas = dict()
    for key, dis in loc_dis_tensor.items():
        for mean_sim in mean_sims:
            left, right = 0, 1000
            while abs(left - right) > 1e-6:                                                                     # :265
                sigma = (left + right) / 2
                sim_feat = torch.exp(- dis / sigma ** 2)
                if sim_feat.mean() < mean_sim:
                    left = sigma
                else:
                    right = sigma
            sigmas[f'{key}_mean@{mean_sim}'] = left
    return sigmas


def loader_pseudo_labels(logits: np.ndarray, thres: np.ndarray, reduce_zero_label: bool = False) -> np.ndarray:
    """logits (C,H,W) float32 as stored in the h5 file, thres (C,) -> uint8 (H,W) labels with 255 for
    rejected pixels (loading.py:474-487: plain exp without max subtraction, log(p + 1e-8))."""
    preds = logits.argmax(axis=0)
    probs = np.exp(logits) / np.exp(logits).sum(axis=0)
    ent_map = - (probs * np.log(probs + 1e-8)).sum(axis=0)
    thre_map = thres[preds]
    mask = ent_map < thre_map
    pse_labels = np.where(mask, preds, 255)
    if reduce_zero_label:
        pse_labels[pse_labels == 0] = 255
        pse_labels = pse_labels - 1
        pse_labels[pse_labels == 254] = 255
    return pse_labels.astype(np.uint8)
