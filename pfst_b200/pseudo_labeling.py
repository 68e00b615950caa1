"""Offline class-wise pseudo-labelling — drop-ins for the arithmetic of PseudoLabelingHookV4
(rsiseg/core/hook/pseudo_labeling_hookv4.py) and of the loader that consumes its
output (rsiseg/datasets/pipelines/loading.py:474-487):

  cal_threshold          _cal_threshold :173-205   per-class entropy quantiles (radix select)
  cal_loc_dis            _cal_loc_dis   :208-230   squared distances to the 3x3 dilated neighbours
  cal_sigmas             _cal_sigmas    :232-277   bisection of sigma on mean(exp(-d/sigma^2))
  loader_pseudo_labels   loading.py:474-487        entropy < thres[argmax] ? argmax : 255 (uint8)

The h5 container the hook writes and the loader reads is not part of this package (h5py is an
I/O dependency of the reference, not arithmetic); these functions take and return tensors.

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


def cal_loc_dis(feats, kernel_size: int = 3, dilations=2) -> dict:
    """PseudoLabelingHookV4._cal_loc_dis: feats = list of (C,H,W) CUDA tensors (one per level) ->
    {'level{l}_dila@{d}': (1,H,W,9) fp32 CUDA tensor} (the reference returns CPU tensors)."""
    if kernel_size != 3:
        raise PfstError("cal_loc_dis: only kernel_size=3 (every shipped sim_feat_cfg)")
    if type(dilations) != list:
        dilations = [dilations]
    out = dict()
    for level, feat in enumerate(feats):
        for dila in dilations:
            out[f'level{level}_dila@{dila}'] = loc_dis_batch(feat.unsqueeze(0), dila)
    return out


def loc_dis_batch(feats: torch.Tensor, dilation: int) -> torch.Tensor:
    """(B,C,H,W) fp32 -> (B,H,W,9) fp32, one launch for the whole batch."""
    feats = feats.contiguous()
    B, C, H, W = feats.shape
    out = torch.empty((B, H, W, 9), dtype=torch.float32, device=feats.device)
    _lib.call("pfst_loc_dis", ops._dev(feats, "feats", torch.float32), B, C, H, W, int(dilation), out.data_ptr(),
              ops._stream())
    return out


def cal_sigmas(loc_dis_list, feat_level, dilations, mean_sims, sample_ratio: float, rng=np.random) -> dict:
    """PseudoLabelingHookV4._cal_sigmas -> {'level{l}_dila@{d}_mean@{m}': sigma (python float)}. The
    random subsets are drawn from the caller's numpy stream in the reference's order; every search
    runs on the device without a host sync (30 launches), one read-back at the end."""
    if type(dilations) != list:
        dilations = [dilations]
    if type(mean_sims) != list:
        mean_sims = [mean_sims]
    subsets = dict()
    for level in feat_level:
        for dila in dilations:
            key = f'level{level}_dila@{dila}'
            cur = torch.cat([ld[key] for ld in loc_dis_list], dim=0) if len(loc_dis_list) > 1 else loc_dis_list[0][key]
            cur = cur.contiguous()
            rowf = cur.shape[-1]
            num_samples = cur.numel() // rowf
            idx = rng.permutation(num_samples)[:int(num_samples * sample_ratio) - 1]          # :255, same stream
            n = int(idx.shape[0])
            if n < 1:
                raise PfstError("cal_sigmas: empty sample")
            idx_d = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(cur.device)
            sub = torch.empty((n, rowf), dtype=torch.float32, device=cur.device)
            _lib.call("pfst_gather_rows", ops._dev(cur, "loc_dis", torch.float32), idx_d.data_ptr(), n, rowf,
                      sub.data_ptr(), ops._stream())
            subsets[key] = sub
    names, states = [], []
    for key, sub in subsets.items():
        for mean_sim in mean_sims:
            st = torch.empty(4, dtype=torch.float64, device=sub.device)
            _lib.call("pfst_sigma_bisect", sub.data_ptr(), sub.numel(), float(mean_sim), 0.0, 1000.0, 30,
                      st.data_ptr(), ops._stream())
            names.append(f'{key}_mean@{mean_sim}')
            states.append(st)
    lefts = torch.stack(states)[:, 0].cpu().tolist() if states else []
    return dict(zip(names, lefts))


def loader_pseudo_labels(seg_logits: torch.Tensor, thres, reduce_zero_label: bool = False) -> torch.Tensor:
    """The label rule of LoadAnnotationsPseudoLabelsV2.__call__ (loading.py:474-487) as a batch op:
    seg_logits (N,C,H,W) or (C,H,W) fp32 CUDA, thres: C thresholds ('thre@<ratio>' of cal_threshold) ->
    uint8 labels (N,H,W) / (H,W), 255 = rejected, on the device."""
    single = seg_logits.dim() == 3
    x = (seg_logits.unsqueeze(0) if single else seg_logits).contiguous()
    N, C, H, W = x.shape
    thr = torch.as_tensor(np.asarray(thres, dtype=np.float32)).to(x.device) if not isinstance(thres, torch.Tensor) \
        else thres.to(device=x.device, dtype=torch.float32).contiguous()
    if thr.numel() != C:
        raise ValueError("thres must have one entry per class")
    out = torch.empty((N, H, W), dtype=torch.uint8, device=x.device)
    _lib.call("pfst_loader_pseudo_labels", ops._dev(x, "seg_logits", torch.float32), N, C, H * W, thr.data_ptr(),
              1 if reduce_zero_label else 0, out.data_ptr(), ops._stream())
    return out[0] if single else out
