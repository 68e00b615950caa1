"""Offline class-wise thresholds: GPU radix-select vs the CPU restatement
of PseudoLabelingHookV4._cal_threshold, same numpy stream. The thresholds are order statistics of
fp32 entropies: 1e-5 relative; empty classes give exactly 0."""
import numpy as np
import pytest
import torch

from oracle import class_thresholds as OT
from pfst_b200 import ops
from pfst_b200.pseudo_labeling import cal_threshold
from pfst_b200.synthetic import teacher_logits

pytestmark = pytest.mark.gpu
RATIOS = [0.1, 0.3, 0.5, 0.8, 0.95]


@pytest.mark.parametrize("B,C,H,W,sample", [(2, 6, 64, 64, 0.5), (1, 33, 40, 40, 1.0), (3, 2, 33, 17, 0.25),
                                             (2, 6, 256, 256, 0.1)])
def test_thresholds_match_reference_rule(cuda, B, C, H, W, sample):
    g = torch.Generator().manual_seed(B * 100 + C)
    logits = teacher_logits(B, C, H, W, g)
    want = OT.cal_threshold(logits, sample, RATIOS, np.random.RandomState(7))
    got = cal_threshold(logits.to(cuda), sample, RATIOS, np.random.RandomState(7))
    assert list(got) == list(want)
    for k in want:
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        assert np.all(np.abs(a - b) <= 1e-5 * np.abs(b) + 1e-9), (k, a, b)


def test_empty_class_gives_zero_and_feeds_the_entropy_rule(cuda):
    g = torch.Generator().manual_seed(3)
    logits = teacher_logits(1, 5, 32, 32, g)
    logits[:, 4] = -50.0                                     # class 4 is never predicted
    want = OT.cal_threshold(logits, 1.0, [0.5], np.random.RandomState(1))
    got = cal_threshold(logits.to(cuda), 1.0, [0.5], np.random.RandomState(1))
    assert got['thre@0.5'][4] == 0 == want['thre@0.5'][4]
    thr = torch.tensor(np.asarray(got['thre@0.5'], dtype=np.float32)).to(cuda)
    label, conf, count, _ = ops.pseudo_label(logits.to(cuda), thr_per_class=thr, mode=1, reject_label=255)
    kept = (label != 255).float().mean().item()
    assert 0.3 < kept < 0.7                                  # about half of the pixels pass a median threshold
