"""Offline class-wise pseudo-label thresholds — drop-in for PseudoLabelingHookV4._cal_threshold
(rsiseg/core/hook/pseudo_labeling_hookv4.py:173-205). The thresholds it
returns are what `ops.pseudo_label(..., thr_per_class=..., mode=1)` (the loading.py:474-487 rule:
keep a pixel iff its entropy is below the threshold of its predicted class) consumes.

The random subset is drawn on the host from the caller's numpy stream exactly as the reference
does; softmax / argmax / entropy and the per-class order statistics run on the GPU
(csrc/class_quantile.cu: radix select, no sort)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from ._lib import PfstError


def cal_threshold(seg_logits: torch.Tensor, sample_ratio: float, cls_thre_ratios, rng=np.random) -> dict:
    """-> {'thre@<ratio>': [C python floats]} like the reference (0 for a class without pixels)."""
    B, C, H, W = seg_logits.shape
    ops._dev(seg_logits, "seg_logits", torch.float32)
    num_samples = B * H * W
    idx = rng.permutation(num_samples)[:int(num_samples * sample_ratio) - 1]            # :180, same stream
    n, R = int(idx.shape[0]), len(cls_thre_ratios)
    if any(not (0.0 <= float(r) < 1.0) for r in cls_thre_ratios):
        raise PfstError("cls_thre_ratios must lie in [0, 1)")
    dev = seg_logits.device
    nbytes = int(_lib.load().pfst_class_quantile_ws_bytes(n, C, R))
    if nbytes <= 0:
        raise PfstError(f"cal_threshold: unsupported size (C={C}, {R} ratios)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    idx_d = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(dev)
    ratios = torch.tensor([float(r) for r in cls_thre_ratios], dtype=torch.float64, device=dev)
    out = torch.empty(C * R, dtype=torch.float32, device=dev)
    _lib.call("pfst_class_quantile", seg_logits.data_ptr(), B, C, H * W, idx_d.data_ptr(), n, ratios.data_ptr(), R,
              ws.data_ptr(), out.data_ptr(), ops._stream())
    thr = out.view(C, R).cpu().numpy()
    return {f'thre@{r}': [thr[c, j] for c in range(C)] for j, r in enumerate(cls_thre_ratios)}
