"""StrongAugmentation — drop-in for the pipeline step of the same name,
rsiseg/datasets/pipelines/transforms.py:1061-1155 (registered in PIPELINES, used by every shipped
dataset config to produce `img_strong_aug`, the teacher-side view `target_img_strong_aug`).

Same constructor arguments, same `__call__(results)` contract (adds `results['img_strong_aug']`,
appends it to `results['img_fields']`), same consumption of the global numpy random stream
(`from numpy import random`, :4): `draw()` makes the reference's draws in the reference's order, the
arithmetic runs in ONE launch of csrc/strong_aug.cu, bit-identical to the reference's numpy + cv2
result (oracle/strong_aug.py is pinned against cv2 over every colour and against the reference class).

`apply_batch` is the form a GPU data path uses: N uint8 images already on the device, N draws.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import PfstError
from .registry import PIPELINES


@PIPELINES.register_module()
class StrongAugmentation(object):

    def __init__(self, brightness_delta=32, contrast_range=(0.5, 1.5), saturation_range=(0.5, 1.5),
                 hue_delta=18, simd_width=32):
        self.brightness_delta = brightness_delta
        self.contrast_lower, self.contrast_upper = contrast_range
        self.saturation_lower, self.saturation_upper = saturation_range
        self.hue_delta = hue_delta
        # block width of cv2's vectorised HSV->BGR on the host the reference runs on (AVX2: 32 pixels);
        # the last W % simd_width pixels of every row are rounded instead of truncated there
        self.simd_width = simd_width

    def draw(self, rng=np.random):
        """transforms.py:1081-1141: the draws of brightness, mode, [contrast], saturation, hue,
        [contrast] -> list of (code, p0, p1) in application order."""
        ops_ = []

        def contrast():
            if rng.randint(2):
                ops_.append((ops.SA_CONVERT, rng.uniform(self.contrast_lower, self.contrast_upper), 0))

        if rng.randint(2):
            ops_.append((ops.SA_CONVERT, 1, rng.uniform(-self.brightness_delta, self.brightness_delta)))
        mode = rng.randint(2)
        if mode == 1:
            contrast()
        if rng.randint(2):
            ops_.append((ops.SA_SATURATION, rng.uniform(self.saturation_lower, self.saturation_upper), 0))
        if rng.randint(2):
            ops_.append((ops.SA_HUE, rng.randint(-self.hue_delta, self.hue_delta), 0))
        if mode == 0:
            contrast()
        return ops_

    def apply_batch(self, imgs: torch.Tensor, op_lists, out=None) -> torch.Tensor:
        """imgs: (N,H,W,3) uint8 CUDA tensor (BGR, HWC, as the pipeline holds images)."""
        return ops.photometric_u8(imgs, op_lists, self.simd_width, out)

    def __call__(self, results):
        img = results['img']
        op_list = self.draw()
        if isinstance(img, np.ndarray):
            if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
                raise PfstError("StrongAugmentation expects a uint8 (H,W,3) image")
            if torch.cuda._is_in_bad_fork():
                # mmcv / torch data-loader workers are forked AFTER the trainer initialised CUDA: the
                # child cannot create a context ("Cannot re-initialize CUDA in forked subprocess")
                raise PfstError("StrongAugmentation (B200 path) was called in a forked data-loader worker, where "
                                "CUDA cannot be initialised. Use workers_per_gpu=0 or the 'spawn' start method, or "
                                "keep this step out of the worker pipeline: call aug.draw() per image there (same "
                                "numpy stream as the reference) and aug.apply_batch(uint8 images on the GPU, draws) "
                                "once per batch before normalisation (INTEGRATION.md §2d)")
            if not torch.cuda.is_available():
                raise PfstError("pfst_b200.pipelines needs a CUDA device (no CPU fallback)")
            if op_list:
                dev = torch.from_numpy(np.ascontiguousarray(img)).cuda().unsqueeze(0)
                img = self.apply_batch(dev, [op_list])[0].cpu().numpy()
            # no distortion drawn: the reference returns the very same array (:1143)
        else:
            if op_list:
                img = self.apply_batch(img.unsqueeze(0).contiguous(), [op_list])[0]
        results['img_strong_aug'] = img
        results['img_fields'].append('img_strong_aug')
        return results

    def __repr__(self):
        repr_str = self.__class__.__name__
        repr_str += (f'(brightness_delta={self.brightness_delta}, '
                     f'contrast_range=({self.contrast_lower}, '
                     f'{self.contrast_upper}), '
                     f'saturation_range=({self.saturation_lower}, '
                     f'{self.saturation_upper}), '
                     f'hue_delta={self.hue_delta})')
        return repr_str
