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


# --------------------------------------------------------------------------------------------
# StrongAugmentation — the data-pipeline photometric distortion that produces
# `target_img_strong_aug` (rsiseg/datasets/pipelines/transforms.py:1062-1145, used by every shipped
# dataset config, e.g. configs/_base_/datasets/pots_irrg2vaih_irrg.py:39).
#
# The class's own arithmetic (`convert`, the hue shift, the draw order) is restated from the
# reference lines. Its colour-space conversions are `mmcv.bgr2hsv` / `mmcv.hsv2bgr`, i.e.
# `cv2.cvtColor(img, cv2.COLOR_BGR2HSV / COLOR_HSV2BGR)` on uint8 images — third-party code that is
# not under /root/reference and not version-pinned by it, but IS installed in this image
# (opencv-python 4.13.0): the two restatements below are PINNED bit-exactly against it over all
# 2^24 BGR triples and all 180*256*256 HSV triples (tests/test_oracle_pins.py), and the whole class
# against the reference class compiled from its source with cv2 standing in for mmcv.
#
# What the brute force established about cv2 4.13 (x86-64, AVX2 dispatch), uint8 images:
#   * BGR->HSV is OpenCV's integer algorithm (12-bit fixed-point division tables), layout-free;
#   * HSV->BGR is computed in fp32: s,v scaled by 1/255, h by 6/180, sector = floor(h),
#     tab = {v, v*(1-s), v*fma(-s,f,1), v*fma(-s,1-f,1)}, result*255 converted to uint8 by
#     TRUNCATION inside the 32-pixel SIMD blocks of a row and by ROUND-HALF-EVEN in the scalar tail
#     (the last W % 32 pixels of every row). The rule is part of the observable behaviour of the
#     reference on such a host; `simd` below is that block width (0 = whole rows vectorised).
_HSV_SHIFT = 12
_SDIV = np.zeros(256, np.int64)
_HDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    _SDIV[_i] = int(np.rint((255 << _HSV_SHIFT) / (1.0 * _i)))
    _HDIV[_i] = int(np.rint((180 << _HSV_SHIFT) / (6.0 * _i)))
_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def bgr2hsv_u8(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(img, cv2.COLOR_BGR2HSV) for uint8 (...,3) images (H in [0,180))."""
    b, g, r = (img[..., k].astype(np.int64) for k in range(3))
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    vr = np.where(v == r, -1, 0)
    vg = np.where(v == g, -1, 0)
    s = (diff * _SDIV[v] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))))
    h = (h * _HDIV[diff] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def _fma32(a, b, c):
    # fused multiply-add of float32 operands: the product is exact in float64
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def hsv2bgr_u8(hsv: np.ndarray, simd: int = 32) -> np.ndarray:
    """cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR) for uint8 (H,W,3) images with H < 180."""
    f32 = np.float32
    h = hsv[..., 0].astype(f32) * f32(6.0 / 180.0)
    s = hsv[..., 1].astype(f32) * f32(1.0 / 255.0)
    v = hsv[..., 2].astype(f32) * f32(1.0 / 255.0)
    sector = np.floor(h).astype(np.int64)
    fr = h - sector.astype(f32)
    one = f32(1.0)
    tab = np.stack([v, v * (one - s), v * _fma32(-s, fr, one), v * _fma32(-s, one - fr, one)], -1)
    bgr = np.take_along_axis(tab, _SECTOR[np.clip(sector, 0, 5)], -1)
    bgr = np.where((hsv[..., 1] == 0)[..., None], v[..., None], bgr) * f32(255.0)
    W = hsv.shape[-2]
    tail = np.arange(W) >= (W - W % simd if simd else W)           # scalar-tail columns of every row
    out = np.where(tail[:, None], np.rint(bgr), np.trunc(bgr))
    return np.clip(out, 0, 255).astype(np.uint8)


def convert_u8(img: np.ndarray, alpha=1, beta=0) -> np.ndarray:
    """StrongAugmentation.convert, transforms.py:1075-1079."""
    img = img.astype(np.float32) * alpha + beta
    img = np.clip(img, 0, 255)
    return img.astype(np.uint8)


# op codes shared with the CUDA kernel: (code, p0, p1)
OP_CONVERT, OP_SATURATION, OP_HUE = 1, 2, 3


def draw_strong_aug(rng=np.random, brightness_delta=32, contrast_range=(0.5, 1.5),
                    saturation_range=(0.5, 1.5), hue_delta=18):
    """The random draws of StrongAugmentation.__call__ in the reference's order
    (transforms.py:1081-1141) -> list of (code, p0, p1) in application order."""
    ops = []

    def contrast():
        if rng.randint(2):
            ops.append((OP_CONVERT, rng.uniform(contrast_range[0], contrast_range[1]), 0))

    if rng.randint(2):                                                     # brightness :1081-1089
        ops.append((OP_CONVERT, 1, rng.uniform(-brightness_delta, brightness_delta)))
    mode = rng.randint(2)                                                  # :1133
    if mode == 1:
        contrast()
    if rng.randint(2):                                                     # saturation :1100-1110
        ops.append((OP_SATURATION, rng.uniform(saturation_range[0], saturation_range[1]), 0))
    if rng.randint(2):                                                     # hue :1112-1121
        ops.append((OP_HUE, rng.randint(-hue_delta, hue_delta), 0))
    if mode == 0:
        contrast()
    return ops


def apply_strong_aug(img: np.ndarray, ops, simd: int = 32) -> np.ndarray:
    """Applies the drawn distortions to a uint8 (H,W,3) BGR image, step by step as the reference."""
    for code, p0, p1 in ops:
        if code == OP_CONVERT:
            img = convert_u8(img, alpha=p0, beta=p1)
        elif code == OP_SATURATION:
            hsv = bgr2hsv_u8(img)
            hsv[:, :, 1] = convert_u8(hsv[:, :, 1], alpha=p0)
            img = hsv2bgr_u8(hsv, simd)
        elif code == OP_HUE:
            hsv = bgr2hsv_u8(img)
            hsv[:, :, 0] = (hsv[:, :, 0].astype(int) + int(p0)) % 180
            img = hsv2bgr_u8(hsv, simd)
        else:
            raise ValueError(code)
    return img


# --------------------------------------------------------------------------------------------
# Colour jitter of the mixed image — the `color_jitter` step of strong_transform,
# rsiseg/models/utils/dacs_transforms.py:56-85.
#
# PARITY UNPINNED. The reference's own lines (restated in `color_jitter` below) only decide whether
# the jitter runs and (de)normalise the image; the sampler and the arithmetic are
# `kornia.augmentation.ColorJitter(brightness=s, contrast=s, saturation=s, hue=s)` — third-party, not
# under /root/reference, not installed here, not version-pinned by the reference. What follows
# restates kornia's published implementation as of the 0.6 series (the releases current when the
# reference was written; `ColorJitter` of that series = additive brightness, multiplicative contrast):
#   generator  (kornia/augmentation/random_generator/_2d/color_jitter.py): factors drawn in the order
#              brightness, contrast, hue, saturation as low + (high - low) * torch.rand(B), then
#              order = torch.randperm(4); ranges: brightness/contrast/saturation [max(0, 1-s), 1+s]
#              (brightness capped at 2), hue [-s, s] (capped at +-0.5);
#   transforms (kornia/enhance/adjust.py, kornia/color/hsv.py), applied in `order`:
#              0 brightness: clamp(x + (f - 1), 0, 1)      1 contrast: clamp(x * f, 0, 1)
#              2 saturation: hsv, s = clamp(s * f, 0, 1)   3 hue: hsv, h = fmod(h + 2 pi f, 2 pi)
# Whether the torch RNG stream of an actual kornia install is consumed identically is NOT
# established; the arithmetic given the factors is what the CUDA kernel is tested against.
import math


def rgb_to_hsv(image: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """kornia.color.rgb_to_hsv: (...,3,H,W) in [0,1] -> h in [0, 2 pi), s, v."""
    max_rgb, argmax_rgb = image.max(-3)
    min_rgb = image.min(-3)[0]
    deltac = max_rgb - min_rgb
    v = max_rgb
    s = deltac / (max_rgb + eps)
    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = torch.unbind((max_rgb.unsqueeze(-3) - image), dim=-3)
    h1 = bc - gc
    h2 = (rc - bc) + 2.0 * deltac
    h3 = (gc - rc) + 4.0 * deltac
    h = torch.stack((h1, h2, h3), dim=-3) / deltac.unsqueeze(-3)
    h = torch.gather(h, dim=-3, index=argmax_rgb.unsqueeze(-3)).squeeze(-3)
    h = (h / 6.0) % 1.0
    h = 2.0 * math.pi * h
    return torch.stack((h, s, v), dim=-3)


def hsv_to_rgb(image: torch.Tensor) -> torch.Tensor:
    """kornia.color.hsv_to_rgb."""
    h = image[..., 0, :, :] / (2 * math.pi)
    s = image[..., 1, :, :]
    v = image[..., 2, :, :]
    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    one = torch.tensor(1.0, dtype=image.dtype)
    p = v * (one - s)
    q = v * (one - f * s)
    t = v * (one - (one - f) * s)
    hi = hi.long()
    indices = torch.stack([hi, hi + 6, hi + 12], dim=-3)
    out = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    return torch.gather(out, -3, indices)


def jitter_ranges(s):
    """kornia `_range_bound` for ColorJitter(brightness=s, contrast=s, saturation=s, hue=s) or a dict."""
    if not isinstance(s, dict):
        s = dict(brightness=s, contrast=s, saturation=s, hue=s)
    b, c, sa, h = (float(s.get(k, 0.0)) for k in ("brightness", "contrast", "saturation", "hue"))
    clamp = lambda lo, hi, a, bnd: (min(max(lo, a), bnd), min(max(hi, a), bnd))
    return dict(brightness=clamp(1 - b, 1 + b, 0.0, 2.0), contrast=clamp(1 - c, 1 + c, 0.0, float("inf")),
                saturation=clamp(1 - sa, 1 + sa, 0.0, float("inf")), hue=clamp(-h, h, -0.5, 0.5))


def draw_jitter(s, batch: int = 1, generator=None):
    """ColorJitterGenerator.forward: brightness, contrast, hue, saturation factors, then the order."""
    r = jitter_ranges(s)
    u = lambda lo_hi: lo_hi[0] + (lo_hi[1] - lo_hi[0]) * torch.rand((batch,), generator=generator)
    out = dict(brightness=u(r["brightness"]), contrast=u(r["contrast"]), hue=u(r["hue"]),
               saturation=u(r["saturation"]))
    out["order"] = torch.randperm(4, generator=generator)
    return out


def apply_jitter(data: torch.Tensor, params) -> torch.Tensor:
    """ColorJitter.apply_transform on (N,3,H,W) images in [0,1]; per-image factors, one shared order."""
    f = lambda k: params[k].to(data.dtype).view(-1, 1, 1, 1)

    def saturation(img):
        hsv = rgb_to_hsv(img)
        s_out = torch.clamp(hsv[:, 1:2] * f("saturation"), min=0, max=1)
        return hsv_to_rgb(torch.cat([hsv[:, 0:1], s_out, hsv[:, 2:3]], dim=1))

    def hue(img):
        hsv = rgb_to_hsv(img)
        h_out = torch.fmod(hsv[:, 0:1] + f("hue") * 2 * math.pi, 2 * math.pi)
        return hsv_to_rgb(torch.cat([h_out, hsv[:, 1:2], hsv[:, 2:3]], dim=1))

    transforms = [lambda img: torch.clamp(img + (f("brightness") - 1), min=0.0, max=1.0),
                  lambda img: torch.clamp(img * f("contrast"), min=0.0, max=1.0),
                  saturation, hue]
    out = data
    for idx in params["order"].tolist():
        out = transforms[idx](out)
    return out


def color_jitter(color_jitter, mean, std, data=None, target=None, s=.25, p=.2, denorm_type='mean_std',
                 generator=None):
    """dacs_transforms.py:56-85 for one (1,3,H,W) mixed image; mean/std (1,3,1,1). -> (data, target, params)."""
    params = None
    if data is not None and data.shape[1] == 3 and color_jitter > p:
        if denorm_type not in ('mean_std', 'none'):
            raise ValueError('No such denorm type!')
        data = data.clone()
        if denorm_type == 'mean_std':
            data.mul_(std).add_(mean).div_(255.0)                       # denorm_ :48-49
        params = draw_jitter(s, data.shape[0], generator)
        data = apply_jitter(data, params)
        if denorm_type == 'mean_std':
            data.mul_(255.0).sub_(mean).div_(std)                       # renorm_ :52-53
    return data, target, params
