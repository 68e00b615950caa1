"""Synthetic inputs of SURVEY.md §8(d) — shared by tests/ and bench.py.

All generators are seeded (1234 + rank by convention) and run on the CPU so that
the oracle and the CUDA path see identical bytes.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class Workload:
    name: str
    B: int          # per-GPU batch
    C: int          # classes
    H: int
    W: int
    D: int = 512    # decoder feature channels (sep_aspp_head.py:91-92)
    downscale: float = 0.5
    dilation: int = 2

    @property
    def h(self):    # decoded features, stride 8
        return self.H // 8

    @property
    def lh(self):   # student logits, stride 4
        return self.H // 4

    @property
    def g(self):    # loss grid
        return int(self.lh * self.downscale)


WORKLOADS = {
    # BASELINE.json configs[0..3]
    "cfg1": Workload("potsdam2vaihingen_B2_C6_512", 2, 6, 512, 512),
    "cfg2": Workload("potsdam2vaihingen_B8_C6_512", 8, 6, 512, 512),
    "cfg3": Workload("inria_B4_C2_1024", 4, 2, 1024, 1024),
    "cfg4": Workload("seasonnet_B64_C33_120", 64, 33, 120, 120, downscale=1.0),
    # small cases for quick parity
    "tiny": Workload("tiny_B2_C6_128", 2, 6, 128, 128, D=32),
    "tiny33": Workload("tiny_B3_C33_48", 3, 33, 48, 48, D=16, downscale=1.0),
}


@dataclass(frozen=True)
class EvalWorkload:
    name: str
    maps: int       # label maps in the whole sweep (sharded over the ranks)
    C: int
    H: int
    W: int


EVAL_WORKLOADS = {
    # BASELINE.json configs[4]
    "cfg5": EvalWorkload("miou_eval_sweep_10000x1024x1024_C6", 10000, 6, 1024, 1024),
    "tiny_eval": EvalWorkload("tiny_eval_8x64x64_C6", 8, 6, 64, 64),
}


def blocky_labels(B: int, H: int, W: int, C: int, gen: torch.Generator, ignore_frac: float = 0.05,
                  min_rect: int = 8, max_rect: int = 128) -> torch.Tensor:
    """(B,1,H,W) int64: random rectangles of random classes, 255 border padding."""
    lab = torch.randint(0, C, (B, 1, 1, 1), generator=gen).expand(B, 1, H, W).clone()
    n_rect = max(4, (H * W) // (48 * 48))
    hi = max(min(max_rect, H, W), min_rect + 1)
    for b in range(B):
        ys = torch.randint(0, H, (n_rect,), generator=gen).tolist()
        xs = torch.randint(0, W, (n_rect,), generator=gen).tolist()
        hs = torch.randint(min_rect, hi, (n_rect,), generator=gen).tolist()
        ws = torch.randint(min_rect, hi, (n_rect,), generator=gen).tolist()
        cs = torch.randint(0, C, (n_rect,), generator=gen).tolist()
        for y, x, hh, ww, c in zip(ys, xs, hs, ws, cs):
            lab[b, 0, y:y + hh, x:x + ww] = c
    pad = int(round(min(H, W) * ignore_frac / 2))
    if pad > 0:
        lab[:, :, :pad, :] = 255
        lab[:, :, -pad:, :] = 255
        lab[:, :, :, :pad] = 255
        lab[:, :, :, -pad:] = 255
    return lab


def teacher_logits(B: int, C: int, H: int, W: int, gen: torch.Generator) -> torch.Tensor:
    """4*randn + 6*onehot(blocky) so that roughly half the pixels exceed 0.98."""
    lab = blocky_labels(B, H, W, C, gen, ignore_frac=0.0)[:, 0]
    x = 4.0 * torch.randn((B, C, H, W), generator=gen)
    x.scatter_add_(1, lab.unsqueeze(1), torch.full((B, 1, H, W), 6.0))
    return x


def step_inputs(wl: Workload, seed: int = 1234) -> dict:
    """Everything one self-training step consumes between the network passes."""
    g = torch.Generator().manual_seed(seed)
    B, C, H, W, D = wl.B, wl.C, wl.H, wl.W, wl.D
    return dict(
        img=torch.randn((B, 3, H, W), generator=g),
        target_img_strong_aug=torch.randn((B, 3, H, W), generator=g),
        gt=blocky_labels(B, H, W, C, g),
        ema_logits=teacher_logits(B, C, H, W, g),
        logits_trg=2.0 * torch.randn((B, C, wl.lh, wl.lh * W // H), generator=g),
        x_src=torch.relu(torch.randn((B, D, wl.h, wl.h * W // H), generator=g)),
        x_ema=torch.relu(torch.randn((B, D, wl.h, wl.h * W // H), generator=g)),
    )


def deeplab_r50_param_shapes(num_classes: int) -> list[tuple[int, ...]]:
    """Parameter shapes (nn.Module.parameters() order is irrelevant for the EMA) of the
    DeepLabV3+ R50-D8 segmentor the reference trains: ResNetV1c-50 backbone,
    DepthwiseSeparableASPPHead, FCN auxiliary head (configs/_base_/models/
    deeplabv3plus_r50-d8.py:3-44). Conv+BN layers have no conv bias. 214 tensors,
    43 579 868 elements at C=6."""
    shapes: list[tuple[int, ...]] = []

    def conv_bn(cout, cin, k):
        shapes.extend([(cout, cin, k, k), (cout,), (cout,)])

    # deep stem 3->32->32->64
    conv_bn(32, 3, 3); conv_bn(32, 32, 3); conv_bn(64, 32, 3)
    inplanes = 64
    for planes, blocks in ((64, 3), (128, 4), (256, 6), (512, 3)):
        for i in range(blocks):
            conv_bn(planes, inplanes, 1)
            conv_bn(planes, planes, 3)
            conv_bn(planes * 4, planes, 1)
            if i == 0:
                conv_bn(planes * 4, inplanes, 1)  # downsample branch
            inplanes = planes * 4
    # DepthwiseSeparableASPPHead (sep_aspp_head.py:43-77, aspp_head.py:64-92)
    conv_bn(512, 2048, 1)                       # image pool branch
    conv_bn(512, 2048, 1)                       # aspp 1x1
    for _ in range(3):                          # aspp depthwise-separable, dilations 12/24/36
        conv_bn(2048, 1, 3)                     # depthwise
        conv_bn(512, 2048, 1)                   # pointwise
    conv_bn(512, 5 * 512, 3)                    # bottleneck 2560->512 3x3
    conv_bn(48, 256, 1)                         # c1 bottleneck
    conv_bn(512 + 48, 1, 3); conv_bn(512, 512 + 48, 1)   # sep bottleneck 1
    conv_bn(512, 1, 3); conv_bn(512, 512, 1)             # sep bottleneck 2
    shapes.extend([(num_classes, 512, 1, 1), (num_classes,)])  # conv_seg
    # FCN auxiliary head on C4 (1024 -> 256)
    conv_bn(256, 1024, 3)
    shapes.extend([(num_classes, 256, 1, 1), (num_classes,)])
    return shapes


def model_params(num_classes: int, gen: torch.Generator, scale: float = 0.02) -> list[torch.Tensor]:
    return [scale * torch.randn(s, generator=gen) for s in deeplab_r50_param_shapes(num_classes)]


def eval_maps(n: int, H: int, W: int, C: int, seed: int, ignore_frac: float = 0.03):
    """cfg5: uniform-random pred (int64) / gt (uint8) with `ignore_frac` of gt = 255."""
    rs = np.random.RandomState(seed)
    pred = rs.randint(0, C, size=(n, H, W)).astype(np.int64)
    gt = rs.randint(0, C, size=(n, H, W)).astype(np.uint8)
    gt[rs.random_sample((n, H, W)) < ignore_frac] = 255
    return pred, gt
