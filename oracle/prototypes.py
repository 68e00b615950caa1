"""Class prototypes (P1-P3) — north_star extension; the reference has NO code for it
(SURVEY.md §0): PARITY UNPINNED. The only anchor in the reference is
PFGST.masked_feat_dist (rsiseg/models/uda/pfgst.py:168-177), which `proto_dist_loss`
calls with f2 = mu[label]. Sums are accumulated in float64 so that the CUDA fp32
path can be judged against a higher-precision answer (1e-5 relative)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def resample_labels(labels: torch.Tensor, size) -> torch.Tensor:
    """(B,H,W) int64 -> (B,h,w) by the nearest rule of pfgst_loss.py:62."""
    return F.interpolate(labels.float().unsqueeze(1), size=size, mode="nearest").long().squeeze(1)


def proto_accumulate(feats: torch.Tensor, labels: torch.Tensor, num_classes: int, conf=None, conf_thr=0.0):
    """sums (C,D) float64, counts (C,) int64 over pixels with 0 <= label < C (and conf >= thr)."""
    B, D, h, w = feats.shape
    lbl = resample_labels(labels, (h, w))
    valid = (lbl >= 0) & (lbl < num_classes)
    if conf is not None:
        c = F.interpolate(conf.unsqueeze(1), size=(h, w), mode="nearest").squeeze(1)
        valid &= c >= conf_thr
    f = feats.permute(0, 2, 3, 1).reshape(-1, D)
    v = valid.reshape(-1)
    idx = lbl.reshape(-1)[v]
    sums = torch.zeros((num_classes, D), dtype=torch.float64).index_add_(0, idx, f[v].double())
    counts = torch.bincount(idx, minlength=num_classes)
    return sums, counts


def proto_finalize(sums, counts, mu_prev=None, seen_prev=None, a32=None, b32=None):
    """mu_c = sums_c / max(cnt_c, 1); classes seen before are EMA-updated with the E2 rule
    fl(fl(a32*mu_prev) + fl(b32*mean)); classes without pixels keep their previous value."""
    C, D = sums.shape
    mean = (sums / counts.clamp(min=1).unsqueeze(1).double()).float()
    has = counts > 0
    if mu_prev is None:
        mu_prev = torch.zeros((C, D))
        seen_prev = torch.zeros(C, dtype=torch.bool)
    mu = mu_prev.clone()
    fresh = has & ~seen_prev
    old = has & seen_prev
    mu[fresh] = mean[fresh]
    if old.any():
        a = torch.tensor(a32, dtype=torch.float32)
        b = torch.tensor(b32, dtype=torch.float32)
        mu[old] = a * mu_prev[old] + b * mean[old]
    return mu, seen_prev | has


def masked_feat_dist(f1, f2, mask=None):
    """PFGST.masked_feat_dist, pfgst.py:168-177."""
    d = torch.norm(f1 - f2, dim=1, p=2)
    if mask is not None:
        d = d[mask.squeeze(1)]
    return torch.mean(d)


def proto_dist_loss(feats, labels, mu, seen):
    """mean over valid pixels of ||f_n - mu[y_n]||_2; valid = label in range and prototype seen."""
    B, D, h, w = feats.shape
    C = mu.shape[0]
    lbl = resample_labels(labels, (h, w))
    valid = (lbl >= 0) & (lbl < C)
    valid &= seen[lbl.clamp(0, C - 1)]
    target = mu[lbl.clamp(0, C - 1)].permute(0, 3, 1, 2)
    return masked_feat_dist(feats, target, valid.unsqueeze(1)), valid


def proto_dist_all(feats, mu):
    """(B,C,h,w): ||f_n - mu_c||_2 for every class."""
    diff = feats.unsqueeze(1) - mu.view(1, mu.shape[0], mu.shape[1], 1, 1)
    return torch.norm(diff, dim=2, p=2)
