"""CPU oracle (TEST INFRASTRUCTURE, see oracle/__init__.py) for the decode-head loss:
BaseDecodeHead.losses (rsiseg/models/decode_heads/decode_head.py:249-283) with the
configured CrossEntropyLoss (rsiseg/models/losses/cross_entropy_loss.py:12-65 +
losses/utils.py:48-79, reduction='mean', avg_non_ignore=False) and accuracy
(rsiseg/models/losses/accuracy.py:6-59). Same ATen operator sequence as the reference.
Pinned live against those reference functions (tests/test_oracle_pins.py) and through
tests/golden/weighted_ce.npz."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def accuracy(pred, target, ignore_index=None):
    """accuracy.py:6-59 for topk=1, thresh=None."""
    if pred.size(0) == 0:
        return pred.new_tensor(0.)
    pred_value, pred_label = pred.topk(1, dim=1)
    pred_label = pred_label.transpose(0, 1)
    correct = pred_label.eq(target.unsqueeze(0).expand_as(pred_label))
    if ignore_index is not None:
        correct = correct[:, target != ignore_index]
    eps = torch.finfo(torch.float32).eps
    correct_k = correct[:1].reshape(-1).float().sum(0, keepdim=True) + eps
    if ignore_index is not None:
        total_num = target[target != ignore_index].numel() + eps
    else:
        total_num = target.numel() + eps
    return correct_k.mul_(100.0 / total_num)


def decode_head_losses(seg_logit, seg_label, seg_weight=None, class_weight=None, ignore_index=255,
                       loss_weight=1.0, align_corners=False):
    """-> (loss_ce 0-dim, acc_seg (1,), up-sampled logits). seg_logit (B,C,lh,lw) float,
    seg_label (B,1,H,W) int64, seg_weight (B,H,W) or None."""
    up = F.interpolate(seg_logit, seg_label.shape[2:], None, 'bilinear', align_corners)   # ops/wrappers.py:28
    label = seg_label.squeeze(1)                                                          # decode_head.py:261
    loss = F.cross_entropy(up, label, weight=class_weight, reduction='none',
                           ignore_index=ignore_index)                                     # cross_entropy_loss.py:45-50
    if seg_weight is not None:
        loss = loss * seg_weight.float()                                                  # utils.py:62-66
    loss = loss_weight * loss.mean()                                                      # utils.py:69-70, :233
    return loss, accuracy(up, label, ignore_index), up
