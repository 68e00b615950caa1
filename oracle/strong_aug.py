"""Gaussian blur of the mixed image — the `gaussian_blur` step of strong_transform,
rsiseg/models/utils/dacs_transforms.py:88-107.

PARITY UNPINNED. The reference's own lines (restated verbatim below) only draw sigma and size
the kernel; the arithmetic is `kornia.filters.GaussianBlur2d(kernel_size, (sigma, sigma))`, a
third-party dependency that is NOT under /root/reference, NOT installed in this image and NOT
version-pinned by the reference (requirements.sh:1 `pip install kornia`). What follows restates
kornia's published algorithm (kornia 0.6/0.7 `filters.gaussian_blur2d`, defaults
border_type='reflect'):
    x      = arange(k) - k // 2                      (+0.5 for even k)
    gauss  = exp(-x^2 / (2 sigma^2));  kernel1d = gauss / gauss.sum()
    kernel2d = kernel1d_y[:, None] @ kernel1d_x[None, :]
    out    = conv2d(pad(input, (k_x//2, k_x//2, k_y//2, k_y//2), mode='reflect'), kernel2d) per channel
(kornia's `separable=True` variant applies kernel1d_x then kernel1d_y; both differ from the 2-D
form by fp32 rounding only). Anchors: the reference call site for kernel size and the sigma draw
(numpy global stream, one `uniform(0.15, 1.15)` per image, after the ClassMix draws), and
mathematical properties tested in tests/ (partition of unity, sigma -> 0 identity, an fp64
evaluation of the same formula).

THIS IS TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def kernel_size(n: int) -> int:
    """dacs_transforms.py:94-101: int(floor(ceil(0.1 n) - 0.5 + ceil(0.1 n) % 2)) (always odd)."""
    return int(np.floor(np.ceil(0.1 * n) - 0.5 + np.ceil(0.1 * n) % 2))


def draw_sigma(rng=np.random) -> float:
    """dacs_transforms.py:93."""
    return rng.uniform(0.15, 1.15)


def gaussian_kernel1d(k: int, sigma: float, dtype=torch.float32) -> torch.Tensor:
    x = torch.arange(k, dtype=dtype) - k // 2
    if k % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_blur2d(data: torch.Tensor, ksize: tuple[int, int], sigma: tuple[float, float],
                    dtype=None) -> torch.Tensor:
    """(N,C,H,W) -> (N,C,H,W); ksize = (k_y, k_x), sigma = (sigma_y, sigma_x)."""
    dt = dtype or data.dtype
    ky, kx = ksize
    k2 = gaussian_kernel1d(ky, sigma[0], dt)[:, None] * gaussian_kernel1d(kx, sigma[1], dt)[None, :]
    N, C, H, W = data.shape
    x = F.pad(data.to(dt), (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect")
    out = F.conv2d(x.reshape(N * C, 1, *x.shape[2:]), k2[None, None])
    return out.reshape(N, C, H, W)


def gaussian_blur(blur: float, data: torch.Tensor, rng=np.random, dtype=None):
    """dacs_transforms.py:88-107 for one (1,3,H,W) mixed image -> (data, sigma or None)."""
    if data is None or data.shape[1] != 3 or not (blur > 0.5):
        return data, None
    sigma = draw_sigma(rng)
    ks = (kernel_size(data.shape[2]), kernel_size(data.shape[3]))
    return gaussian_blur2d(data, ks, (sigma, sigma), dtype), sigma
