"""One whole self-training hot-path step on the CPU, composed from the oracle pieces in
the order of PFGST.forward_train (rsiseg/models/uda/pfgst.py:203-344) plus the
north_star prototype extension. Used as the checker in tests/smoke and as the timed
CPU baseline of bench.py (`cpu_baseline`, `--impl reference`)."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import ema as oema, mixing as omix, pfgst_loss as OL, prototypes as OP, pseudo as opl


def hot_path_step(it, teacher, student, inp, C, alpha=0.999, thr=0.98, loss_cfg=None, proto_state=None,
                  proto_weight=0.1, rng=np.random, timings=None, peer_protos=None):
    """inp: dict(img, target_img_strong_aug, gt, ema_logits, logits_trg, x_src, x_ema) CPU tensors.
    Returns dict of outputs; `timings` (dict) receives per-phase seconds. `peer_protos`: list of
    (sums, counts) accumulated by the other ranks (what the NCCL all-reduce contributes)."""
    loss_cfg = loss_cfg or OL.LossCfg()
    t0 = time.perf_counter()
    if it == 0:
        oema.ema_init(teacher, student)
    else:
        oema.ema_update(teacher, student, it, alpha)
    t1 = time.perf_counter()
    label, prob, large = opl.pseudo_label(inp["ema_logits"], thr)
    weight = opl.pseudo_weight(large, "all")
    masks = omix.class_masks(inp["gt"], rng)
    mixed_img, mixed_lbl, weight, mix_masks = omix.mix_batch(inp["img"], inp["target_img_strong_aug"], inp["gt"],
                                                             label, weight, masks)
    t2 = time.perf_counter()
    logits = inp["logits_trg"].clone().requires_grad_(True)
    x_src = inp["x_src"].clone().requires_grad_(True)
    res = OL.pfgst_loss(dict(logits_trg=logits, gt_src=inp["gt"], x_ema=inp["x_ema"], x_src=x_src, img_trg=None,
                             mix_masks=mix_masks), loss_cfg)
    total = sum(res[k] for k in OL.LOSS_KEYS)
    # prototypes (float64 sums -> fp32 prototypes), distance loss on the source features
    sums, counts = OP.proto_accumulate(inp["x_ema"], label, C)
    for ps, pc in (peer_protos or []):          # multi-rank: the all-reduce adds the other ranks' sums / counts
        sums, counts = sums + ps, counts + pc
    mu_prev, seen_prev, pit = proto_state if proto_state is not None else (None, None, 0)
    a = oema.alpha_teacher(max(pit, 1), alpha)
    mu, seen = OP.proto_finalize(sums, counts, mu_prev, seen_prev, float(np.float32(a)), float(np.float32(1 - a)))
    ploss, _ = OP.proto_dist_loss(x_src, inp["gt"][:, 0], mu, seen)
    (total + proto_weight * ploss).backward()
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(ema=t1 - t0, pseudo_mix=t2 - t1, loss_proto=t3 - t2)
    return dict(losses=torch.stack([res[k].detach().reshape(()) for k in OL.LOSS_KEYS]), proto_loss=ploss.detach(),
                pseudo_label=label, pseudo_conf=prob, large=large, mixed_img=mixed_img, mixed_lbl=mixed_lbl,
                pseudo_weight=weight, mix_masks=mix_masks, grad_x_src=x_src.grad, grad_logits_trg=logits.grad,
                mu=mu, proto_state=(mu, seen, pit + 1))
