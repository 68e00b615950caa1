"""SelfTrainingStep — the whole per-iteration hot path as one sync-light launch
sequence (SURVEY.md §3.2 steps ①②⑤⑥⑦⑨⑩ + north_star P1-P3), given the outputs of the
three network passes. It issues the same kernels as the PFGST drop-in
(pfst_b200/uda/pfgst.py) but calls the forward/backward kernels directly instead of
through autograd, so that no PyTorch kernel sits between them:

  main stream:  presence(gt) -> D2H 36 B -> [event]
                EMA update (1 launch, all tensors)                    E2
                pseudo_label(ema_logits)                              S1/S2
                [host: wait event, np.random.choice per image, H2D]   M1
                class_mix                                             M2
                neigh_dots(x_ema, x_src)                              L2
                proto_accum(x_ema, pseudo_label) -> NCCL all-reduce   P1 (async)
                pfgst_loss_fwd                                        L1,L3-L6
                proto_finalize, proto_dist_fwd(x_src, gt)             P2,P3
                pfgst_loss_bwd, neigh_grad, proto_dist_bwd(+=)        backward

The only host round trip is the 36-byte class-presence read, hidden behind the EMA
kernel. Used by bench.py and __graft_entry__.smoke(); a trainer that owns its
autograd graph can call it in place of the aux-loss section of forward_train.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from .prototypes import PrototypeBank
from .utils.dacs_transforms import ClassMixPlan

W6_DEFAULT = (0.1, 0.1, 0.1, 0.1, 0.1, 0.1)   # src_pos, src_neg, src_pos_std, src_neg_std, sim_pos, sim_neg


class SelfTrainingStep:
    # launches of THIS library's kernels per run() (EMA 1, presence 1, pseudo-label 1, mix 1,
    # dots 1, proto accum 1, loss fwd 1, finalize 1, dist fwd 1, loss bwd 1, neigh grad 1,
    # dist bwd 1); memsets and NCCL are not counted
    KERNEL_LAUNCHES = 12

    def __init__(self, teacher_params, student_params, num_classes: int, feat_dim: int, device,
                 alpha: float = 0.999, pseudo_threshold: float = 0.98, dilation: int = 2, top_k: int = 3,
                 downscale: Optional[float] = 0.5, weights6=W6_DEFAULT, proto_weight: float = 0.1,
                 max_batch: int = 64, group=None):
        self.device = torch.device(device)
        self.alpha, self.thr = alpha, pseudo_threshold
        self.dilation, self.top_k, self.downscale = dilation, top_k, downscale
        self.w6 = tuple(float(v) for v in weights6)
        self.proto_weight = float(proto_weight)
        self.C, self.D = num_classes, feat_dim
        self.table = ops.EmaTable(teacher_params, student_params)
        self.plan = ClassMixPlan(self.device, max_batch=max_batch)
        self.bank = PrototypeBank(num_classes, feat_dim, self.device, alpha=alpha, group=group)
        self.gout = torch.ones(6, dtype=torch.float32, device=self.device)
        self.gproto = torch.full((1,), self.proto_weight, dtype=torch.float32, device=self.device)
        self.ema_events = None    # optional (start, end) CUDA events around the EMA launch

    def run(self, it: int, img, trg_img, gt, ema_logits, logits_trg, x_src, x_ema, rng=np.random):
        B = img.shape[0]
        # M1 part 1: presence bits + tiny D2H, overlapped with the kernels below
        self.plan.start(gt)
        # E2
        if self.ema_events is not None:
            self.ema_events[0].record()
        if it == 0:
            self.table.update(0.0, 1.0, mode=1)
        else:
            self.table.update(*ops.ema_coeffs(it, self.alpha))
        if self.ema_events is not None:
            self.ema_events[1].record()
        # S1/S2
        label, conf, count, _ = ops.pseudo_label(ema_logits, self.thr)
        # M1 part 2 (host) + M2
        chosen = self.plan.choose(rng)
        mixed_img, mixed_lbl, weight, mix_mask = ops.class_mix(gt, chosen, img, trg_img, label, count=count,
                                                               ps_size=label.numel())
        # L2
        geo = ops.LossGeometry(logits_trg.shape, x_src.shape, gt.shape, self.downscale, self.dilation)
        dots, ks = ops.neigh_dots(x_ema, x_src, geo.dilation // geo.up)
        # P1 (+ all-reduce in flight while the loss statistics run)
        self.bank.accumulate(x_ema, label)
        work = self.bank.all_reduce()
        # L1, L3-L6
        losses, stats, density, eroded = ops.pfgst_loss_fwd(dots, ks, geo, logits_trg, gt, mix_mask, self.top_k,
                                                            self.w6, want_vis=False)
        # P2, P3
        mu = self.bank.finalize(work)
        Bf, D, h, w = x_src.shape
        lab3 = gt.reshape(B, gt.shape[-2], gt.shape[-1])
        dist = torch.empty((Bf, h, w), dtype=torch.float32, device=self.device)
        acc = torch.empty(4, dtype=torch.float64, device=self.device)
        ploss = torch.empty(1, dtype=torch.float32, device=self.device)
        _lib.call("pfst_proto_dist_fwd", x_src.data_ptr(), Bf, D, h, w, lab3.data_ptr(), lab3.shape[-2],
                  lab3.shape[-1], mu.data_ptr(), self.bank.seen.data_ptr(), self.C, dist.data_ptr(),
                  acc.data_ptr(), ploss.data_ptr(), ops._stream())
        # backward of (sum of the six losses + proto_weight * proto loss)
        coef, grad_logits = ops.pfgst_loss_bwd(dots, ks, geo, logits_trg, gt, mix_mask, self.top_k, self.w6, stats,
                                               self.gout)
        grad_x = ops.neigh_grad(x_src, coef, geo.dilation // geo.up)
        _lib.call("pfst_proto_dist_bwd", x_src.data_ptr(), Bf, D, h, w, lab3.data_ptr(), lab3.shape[-2],
                  lab3.shape[-1], mu.data_ptr(), self.bank.seen.data_ptr(), self.C, dist.data_ptr(),
                  acc.data_ptr(), self.gproto.data_ptr(), grad_x.data_ptr(), 1, ops._stream())
        return dict(losses=losses, proto_loss=ploss, pseudo_label=label, pseudo_conf=conf, count=count,
                    mixed_img=mixed_img, mixed_lbl=mixed_lbl, pseudo_weight=weight, mix_masks=mix_mask,
                    grad_x_src=grad_x, grad_logits_trg=grad_logits, mu=mu)


def algorithmic_bytes(B: int, C: int, H: int, W: int, D: int, h: int, w: int, n_params: int) -> dict:
    """Algorithmic HBM bytes per step and per kernel."""
    P, p = B * H * W, B * h * w
    return {
        "ema": 12 * n_params,
        "pseudo_label": (4 * C + 12) * P,
        "class_presence": 8 * P,
        "class_mix": 80 * P,                 # thre_type='all': the incoming weight is a scalar
        "neigh_dots": 2 * 4 * D * p,
        "proto_accum": 4 * D * p + 8 * P // 64,
        "loss_maps": (2 * 5 * 4 + 9 * 4 * 2) * p,
        "proto_dist_fwd": 4 * D * p,
        "neigh_grad": 2 * 4 * D * p,
        "proto_dist_bwd": 3 * 4 * D * p,
    }
