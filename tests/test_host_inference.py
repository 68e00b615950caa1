"""Host logic of the test-time inference mirror (pfst_b200/evaluation/inference.py) that needs no GPU:
the window grid of slide_inference and its separable count matrix, flip parsing, option errors."""
import numpy as np
import pytest
import torch

from pfst_b200 import evaluation as E
from pfst_b200._lib import PfstError
from pfst_b200.evaluation.inference import _flips
from tests.golden.make_golden import slide_meta, synthetic_encode_decode


@pytest.mark.parametrize("h,w,crop,stride", [(24, 40, (16, 16), (8, 12)), (10, 14, (16, 12), (8, 8)),
                                             (67, 93, (32, 40), (21, 17)), (512, 512, (256, 256), (171, 171)),
                                             (5, 5, (5, 5), (5, 5))])
def test_window_grid_is_the_reference_loop_and_counts_are_separable(h, w, crop, stride):
    wins, cy, cx = E.window_grid(h, w, crop, stride)
    # encoder_decoder.py:227-241, statement for statement
    h_stride, w_stride = stride
    h_crop, w_crop = crop
    h_grids = max(h - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(w - w_crop + w_stride - 1, 0) // w_stride + 1
    want, cm = [], np.zeros((h, w), dtype=np.float32)
    for h_idx in range(h_grids):
        for w_idx in range(w_grids):
            y1, x1 = h_idx * h_stride, w_idx * w_stride
            y2, x2 = min(y1 + h_crop, h), min(x1 + w_crop, w)
            y1, x1 = max(y2 - h_crop, 0), max(x2 - w_crop, 0)
            want.append((y1, y2, x1, x2))
            cm[y1:y2, x1:x2] += 1
    assert wins == want
    assert np.array_equal(cm, np.outer(cy, cx)) and (cm > 0).all()


def test_flip_directions_follow_the_reference_loop():
    assert _flips(slide_meta(False, None, (4, 4), 1)) == (False, False)
    assert _flips(slide_meta(True, "horizontal", (4, 4), 1)) == (True, False)
    assert _flips(slide_meta(True, ["vertical", "horizontal"], (4, 4), 1)) == (True, True)
    assert _flips(slide_meta(True, ["horizontal", "horizontal"], (4, 4), 1)) == (False, False)   # flipped twice
    with pytest.raises(AssertionError):
        _flips(slide_meta(True, "diagonal", (4, 4), 1))


def test_cpu_tensors_are_rejected_without_touching_a_device():
    enc = synthetic_encode_decode(3)
    with pytest.raises(PfstError):
        E.slide_logits(enc, torch.zeros(1, 3, 8, 8), slide_meta(False, None, (8, 8), 1), (4, 4), (4, 4), 3)
    with pytest.raises(AssertionError):
        E.inference_logits(enc, torch.zeros(1, 3, 8, 8), slide_meta(False, None, (8, 8), 1), True,
                           dict(mode="tiled", crop_size=None, stride=None), 3)
