"""Test-time sliding-window / flip inference (encoder_decoder.py:220-353) on the GPU: the window
accumulation, count division and flips of `pfst_slide_add` / `pfst_slide_finalize` and the fused
soft-max arg-max, against the fixture the reference methods wrote and the CPU oracle. The averaged
logits are bit-exact; the soft-max output is compared at 1e-6 (torch's CPU and CUDA soft-max kernels
differ in the last bit) and the arg-max maps exactly outside 1e-6 near-ties."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import metrics as om
from pfst_b200 import evaluation as E
from pfst_b200._lib import PfstError
from tests.golden.make_golden import slide_cases, slide_meta, synthetic_encode_decode

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _same_labels(pred, want_pred, prob):
    """arg-max maps equal wherever the two largest soft-max values are more than 1e-6 apart."""
    top2 = np.sort(prob, axis=1)[:, -2:]
    safe = (top2[:, 1] - top2[:, 0]) > 1e-6
    return np.array_equal(pred[safe], want_pred[safe])


def _inputs(case):
    name, B, C, H, W, mode, crop, stride, flip, direction, ori, seed = case
    g = torch.Generator().manual_seed(900 + seed)
    return torch.randn((B, 3, H, W), generator=g), slide_meta(flip, direction, ori, B), synthetic_encode_decode(C)


@pytest.mark.parametrize("case", slide_cases(), ids=[c[0] for c in slide_cases()])
def test_matches_reference_golden(cuda, case):
    name, B, C, H, W, mode, crop, stride = case[:8]
    z = np.load(G / "slide_inference.npz")
    img, meta, enc = _inputs(case)
    cfg = E.make_test_cfg(mode, crop, stride)
    out, _ = E.inference(enc, img.to(cuda), meta, True, cfg, C)
    assert np.allclose(out.cpu().numpy(), z[f"{name}_output"], rtol=0, atol=1e-6)
    pred, states = E.simple_test(enc, img.to(cuda), meta, True, test_cfg=cfg, num_classes=C)
    assert len(pred) == B and pred[0].dtype == np.int64 and states == [{} for _ in range(B)]
    assert _same_labels(np.stack(pred), z[f"{name}_pred"].astype(np.int64), z[f"{name}_output"])
    if mode == "slide":
        raw = E.slide_inference(enc, img.to(cuda), meta, False, crop_size=crop, stride=stride, num_classes=C)
        assert np.array_equal(raw.cpu().numpy(), z[f"{name}_slide"])


@pytest.mark.parametrize("H,W,crop,stride,direction", [(96, 128, (64, 64), (32, 48), None),
                                                       (67, 93, (32, 40), (21, 17), ["horizontal"]),
                                                       (128, 128, (128, 128), (85, 85), "vertical"),
                                                       (50, 70, (64, 32), (40, 20), ["vertical", "horizontal"])])
def test_matches_oracle_on_unaligned_windows(cuda, H, W, crop, stride, direction):
    C, B = 6, 3
    g = torch.Generator().manual_seed(H * 1000 + W)
    img = torch.randn((B, 3, H, W), generator=g)
    enc = synthetic_encode_decode(C)
    meta = slide_meta(direction is not None, direction, (H, W), B)
    want, _ = om.inference(enc, img, meta, True, "slide", crop, stride, C)
    got, _ = E.inference(enc, img.to(cuda), meta, True, dict(mode="slide", crop_size=crop, stride=stride), C)
    assert torch.allclose(got.cpu(), want, rtol=0, atol=1e-6)
    # the averaged logits themselves are bit-exact (flip applied to the oracle's un-flipped slide output)
    raw_want = om.slide_inference(enc, img, meta, False, crop, stride, C)
    for d in (direction if isinstance(direction, list) else [direction] if direction else []):
        raw_want = raw_want.flip(dims=(3,) if d == "horizontal" else (2,))
    gt = torch.randint(0, C, (B, H, W), generator=g).to(torch.uint8)
    logits, _ = E.inference_logits(enc, img.to(cuda), meta, True, dict(mode="slide", crop_size=crop, stride=stride), C)
    assert torch.equal(logits.cpu(), raw_want)
    res = E.pre_eval_logits(logits, gt.to(cuda), C, 255)

    pred = want.argmax(dim=1).numpy()
    top2 = np.sort(want.numpy(), axis=1)[:, -2:]
    if ((top2[:, 1] - top2[:, 0]) > 1e-6).all():            # no near-tie: the per-image areas must be identical
        for i in range(B):
            a = om.areas(pred[i], gt[i].numpy(), C, 255, None, False)
            for x, y in zip(res[i], a):
                assert np.array_equal(np.asarray(x), np.asarray(y))


def test_window_grid_counts_equal_count_mat():
    for (h, w, crop, stride) in [(24, 40, (16, 16), (8, 12)), (10, 14, (16, 12), (8, 8)), (67, 93, (32, 40), (21, 17))]:
        wins, cy, cx = E.window_grid(h, w, crop, stride)
        cm = np.zeros((h, w), dtype=np.float32)
        for y1, y2, x1, x2 in wins:
            cm[y1:y2, x1:x2] += 1
        assert np.array_equal(cm, np.outer(cy, cx))


def test_cpu_tensors_are_rejected():
    enc = synthetic_encode_decode(3)
    with pytest.raises(PfstError):
        E.slide_logits(enc, torch.zeros(1, 3, 8, 8), slide_meta(False, None, (8, 8), 1), (4, 4), (4, 4), 3)


def test_aug_test_matches_reference_golden(cuda):
    """aug_test (encoder_decoder.py:355-373): soft-max accumulation over augmentations + arg-max, against the
    fixture the reference method wrote; labels equal wherever the two largest averaged probabilities are more
    than 1e-6 apart (CPU vs CUDA soft-max differ in the last bit)."""
    from tests.golden.make_golden import aug_cases, aug_inputs
    z = np.load(G / "aug_test.npz")
    for case in aug_cases():
        name, B, C, ori, mode, crop, stride, augs, seed = case
        imgs, metas = aug_inputs(case)
        enc = synthetic_encode_decode(C)
        pred, st = E.aug_test(enc, [im.to(cuda) for im in imgs], metas, True,
                              test_cfg=E.make_test_cfg(mode, crop, stride), num_classes=C)
        assert st == {} and len(pred) == B and pred[0].dtype == np.int64
        assert _same_labels(np.stack(pred), z[f"{name}_pred"].astype(np.int64), z[f"{name}_avg"]), name
        assert (np.stack(pred) == z[f"{name}_pred"]).mean() > 0.999


def test_softmax_accum_equals_torch_cuda_softmax(cuda):
    """pfst_softmax_accum restates torch's CUDA soft-max along dim 1: bit-identical sums on the device."""
    from pfst_b200 import _lib, ops
    g = torch.Generator().manual_seed(5)
    xs = [(3.0 * torch.randn((2, 7, 33, 21), generator=g)).to(cuda) for _ in range(3)]
    acc = torch.empty_like(xs[0])
    want = None
    for i, x in enumerate(xs):
        _lib.call("pfst_softmax_accum", x.data_ptr(), acc.data_ptr(), 2, 7, 33 * 21, int(i == 0), ops._stream())
        sm = torch.softmax(x, dim=1)
        want = sm if want is None else want + sm
    assert torch.equal(acc, want)
    pred = torch.empty((2, 33, 21), dtype=torch.int64, device=cuda)
    _lib.call("pfst_div_argmax", acc.data_ptr(), 2, 7, 33 * 21, 3.0, pred.data_ptr(), ops._stream())
    assert torch.equal(pred, (want / 3).argmax(dim=1))
