"""PFGST auxiliary loss — restates rsiseg/models/losses/pfgst_loss.py:44-234 for
src_loss_type='mean_std', feat_level=None, no proj_net, src_perc=None: the shipped
configuration (sim_type='cosine', cross_prob_type='trg', detach_unfold=True) and the
options sim_type='gaussian' (:189-191), cross_prob_type='ema' (:161-178),
detach_unfold=False, src_loss_type='margin'/'margin2' (:117-133) and top_k=None (:218-220). Test infrastructure: only tests/, smoke() and bench.py's CPU legs use it.

Written as free functions over the same ATen operator sequence as the reference
(nn.Unfold / F.interpolate / F.cosine_similarity / topk / boolean gathers) so that
results are bit-identical to it on the same torch build and its wall time is a
fair CPU baseline. Device-agnostic (the reference hard-codes .cuda() at :225-226).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class LossCfg:
    top_k: int | None = 3
    dilation: int = 2
    kernel_size: int = 3
    weights: dict = field(default_factory=lambda: {"src_pos": 0.1, "src_neg": 0.1, "sim_pos": 0.1,
                                                   "sim_neg": 0.1, "src_pos_std": 0.1, "src_neg_std": 0.1})
    detach_unfold: bool = True
    downscale: float | None = 0.5
    sim_type: str = "cosine"
    sigma: float = 30.0
    cross_prob_type: str = "trg"
    src_loss_type: str = "mean_std"
    margin: tuple = (0.5, 0.5)


def _unfold(x: torch.Tensor, cfg: LossCfg) -> torch.Tensor:
    # pfgst_loss.py:29-31 — zero padding k//2*dilation
    return F.unfold(x, kernel_size=cfg.kernel_size, padding=cfg.kernel_size // 2 * cfg.dilation,
                    dilation=cfg.dilation)


def neighbourhood_cosine(x: torch.Tensor, size, cfg: LossCfg):
    """get_sim_feat, pfgst_loss.py:181-201 -> (feats, sim (B,k*k,H,W)); cosine (:193-196) or
    gaussian (:189-191) by cfg.sim_type."""
    B, ch = x.shape[:2]
    k2 = cfg.kernel_size ** 2
    feats = F.interpolate(x, size=size, mode="nearest")
    unf = _unfold(feats, cfg).view(B, ch, k2, size[0], size[1]).permute(0, 1, 3, 4, 2)
    if cfg.sim_type == "gaussian":
        dis = ((unf - feats.unsqueeze(4)) ** 2).sum(dim=1)
        return feats, torch.exp(-dis / cfg.sigma ** 2).permute(0, 3, 1, 2)
    if cfg.sim_type != "cosine":
        raise ValueError()
    sim = F.cosine_similarity(unf, feats.unsqueeze(4), dim=1)
    return feats, sim.permute(0, 3, 1, 2)


def cross_prob_diag(logits: torch.Tensor, cfg: LossCfg) -> torch.Tensor:
    """get_cross_prob_map_diag, pfgst_loss.py:142-159 -> (B,C,H,W,k*k)."""
    B, C, H, W = logits.shape
    k2 = cfg.kernel_size ** 2
    p = F.softmax(logits, dim=1)
    q = _unfold(p, cfg)
    if cfg.detach_unfold:
        q = q.detach()
    q = q.view(B, -1, k2, H, W).permute(0, 1, 3, 4, 2)
    return p.unsqueeze(4).repeat(1, 1, 1, 1, k2) * q


def cross_prob_diag_ema(logits_trg: torch.Tensor, logits_ema: torch.Tensor, cfg: LossCfg) -> torch.Tensor:
    """get_cross_prob_map_diag_ema, pfgst_loss.py:161-178 -> (B,C,H,W,k*k): q from the teacher's logits."""
    B, C, H, W = logits_trg.shape
    k2 = cfg.kernel_size ** 2
    p = F.softmax(logits_trg, dim=1)
    q = _unfold(F.softmax(logits_ema, dim=1), cfg).view(B, -1, k2, H, W).permute(0, 1, 3, 4, 2)
    return p.unsqueeze(4).repeat(1, 1, 1, 1, k2) * q


def consistency_losses(sim: torch.Tensor, cross: torch.Tensor, mask: torch.Tensor, cfg: LossCfg):
    """get_sim_losses, pfgst_loss.py:203-234 (top_k branch, ignore_mask given)."""
    cp = cross.sum(dim=1).permute(0, 3, 1, 2)
    cn = 1 - cp
    if cfg.top_k is not None:
        _, imax = torch.topk(sim, cfg.top_k + 1, dim=1)
        _, imin = torch.topk(sim, cfg.top_k, dim=1, largest=False)
        loc_pos = torch.gather(sim, 1, imax) * (-torch.gather(cp, 1, imax))
        loc_neg = (1 - torch.gather(sim, 1, imin)) * (-torch.gather(cn, 1, imin))
    else:                                    # pfgst_loss.py:218-220
        loc_pos = sim * (-cp)
        loc_neg = (1 - sim) * (-cn)
    l_pos = torch.zeros(1, device=sim.device)
    l_neg = torch.zeros(1, device=sim.device)
    if mask.sum() > 1:
        l_pos = loc_pos[mask.repeat(1, loc_pos.shape[1], 1, 1)].mean()
        l_neg = loc_neg[mask.repeat(1, loc_neg.shape[1], 1, 1)].mean()
    return l_pos, l_neg


def pfgst_loss(tensors: dict, cfg: LossCfg) -> dict:
    """PFGSTLoss.forward, pfgst_loss.py:44-140. Keys used: logits_trg, gt_src, x_ema,
    x_src, img_trg, mix_masks. Returns the six loss_* scalars + 'vis|density_sim_feat'."""
    logits_trg = tensors["logits_trg"]
    gt_src, x_ema, x_src = tensors["gt_src"], tensors["x_ema"], tensors["x_src"]
    k2 = cfg.kernel_size ** 2
    if cfg.downscale is not None:
        logits_trg = F.interpolate(logits_trg, scale_factor=(cfg.downscale, cfg.downscale))
        x_ema = F.interpolate(x_ema, size=logits_trg.shape[2:])
        x_src = F.interpolate(x_src, size=logits_trg.shape[2:])
    B, C, H, W = logits_trg.shape
    gt = F.interpolate(gt_src.float(), size=(H, W), mode="nearest")
    valid_src = gt != 255

    trg = F.interpolate((1 - tensors["mix_masks"]).float(), size=(H, W), mode="nearest") > 0.5
    unf_trg = _unfold(trg.float(), cfg).view(-1, k2, H, W).long()
    trg_eroded = unf_trg.sum(dim=1).unsqueeze(1) == k2

    if cfg.cross_prob_type == "ema":
        cross = cross_prob_diag_ema(logits_trg, tensors["logits_ema"], cfg)
    else:
        cross = cross_prob_diag(logits_trg, cfg)
    _, sim_ema = neighbourhood_cosine(x_ema, (H, W), cfg)
    _, sim_src = neighbourhood_cosine(x_src, (H, W), cfg)

    unf_gt = _unfold(gt.float(), cfg).view(-1, k2, H, W).long()
    rep_gt = gt.repeat(1, k2, 1, 1)
    pos_pair = unf_gt == rep_gt
    neg_pair = unf_gt != rep_gt
    keep = valid_src.repeat(1, k2, 1, 1)
    pos = sim_src[pos_pair & keep]
    neg = sim_src[neg_pair & keep]

    l_pos, l_neg = consistency_losses(sim_ema, cross, valid_src & trg_eroded, cfg)
    w = cfg.weights
    if cfg.src_loss_type in ("margin", "margin2"):           # pfgst_loss.py:117-133
        hp, hn = F.relu(cfg.margin[0] - pos), F.relu(neg - cfg.margin[1])
        if cfg.src_loss_type == "margin2":
            hp, hn = hp ** 2, hn ** 2
        return {
            "loss_src_pos": hp.mean() * w["src_pos"],
            "loss_src_neg": hn.mean() * w["src_neg"],
            "loss_sim_pos": l_pos * w["sim_pos"],
            "loss_sim_neg": l_neg * w["sim_neg"],
            "vis|density_sim_feat": (tensors.get("img_trg"), 1 - sim_ema.mean(dim=1).detach().unsqueeze(1),
                                     trg_eroded),
        }
    return {
        "loss_src_pos_mean": -pos.mean() * w["src_pos"],
        "loss_src_neg_mean": neg.mean() * w["src_neg"],
        "loss_src_pos_std": pos.std() * w["src_pos_std"],
        "loss_src_neg_std": neg.std() * w["src_neg_std"],
        "loss_sim_pos": l_pos * w["sim_pos"],
        "loss_sim_neg": l_neg * w["sim_neg"],
        "vis|density_sim_feat": (tensors.get("img_trg"), 1 - sim_ema.mean(dim=1).detach().unsqueeze(1),
                                 trg_eroded),
    }


LOSS_KEYS = ("loss_src_pos_mean", "loss_src_neg_mean", "loss_src_pos_std", "loss_src_neg_std",
             "loss_sim_pos", "loss_sim_neg")
