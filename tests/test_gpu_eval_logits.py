"""On-device evaluation input path: fused softmax-argmax + confusion matrix
vs the oracle restatement of simple_test + pre_eval (encoder_decoder.py:311,329-338;
custom.py:644-682; metrics.py:26-86).

Predictions are bit-exact except pixels whose two largest softmax outputs lie within 1e-6 of each
other (a 1-ulp expf difference between CPU and CUDA may move the arg-max there, as north_star
allows for the pseudo-labels); the confusion matrix is always integer-exact against the oracle
count of the kernel's OWN prediction map, and against the oracle end to end when no such pixel exists."""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from pfst_b200 import ops
from pfst_b200.evaluation import metrics as M
from pfst_b200.synthetic import blocky_labels, teacher_logits

pytestmark = pytest.mark.gpu


def _labels(N, H, W, C, seed, dtype=torch.uint8):
    g = torch.Generator().manual_seed(seed)
    return blocky_labels(N, H, W, C, g)[:, 0].to(dtype)


def _check(cuda, logits, gt, C, ignore=255, label_map=None, rzl=False, pred_dtype=torch.int64):
    N = logits.shape[0]
    pred_o = om.seg_argmax(logits)
    lut = M._label_lut(label_map, cuda)
    conf, pred = ops.argmax_confusion(logits.to(cuda), gt.to(cuda), C, ignore, rzl, lut, per_image=True,
                                      return_pred=pred_dtype)
    assert pred.dtype == pred_dtype and conf.shape == (N, C + 1, C + 1)
    pred = pred.cpu().long()
    near_tie = torch.zeros_like(pred_o, dtype=torch.bool)
    if C > 1:
        top2 = torch.softmax(logits, 1).topk(2, dim=1).values
        near_tie = (top2[:, 0] - top2[:, 1]) < 1e-6
    assert torch.equal(pred[~near_tie], pred_o[~near_tie])
    # the matrix is the exact count of the kernel's own predictions ...
    gt_np = gt.numpy()
    for i in range(N):
        want = om.areas(pred[i].numpy(), gt_np[i], C, ignore, label_map, rzl)
        got = [a[i].cpu().float() for a in M._areas_from_conf(conf, C)]
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # ... and the reference's result end to end when no pixel is a near tie
    if not bool(near_tie.any()):
        want = om.pre_eval(logits, list(gt_np), C, ignore, label_map, rzl)
        got = M.pre_eval_logits(logits.to(cuda), gt.to(cuda), C, ignore, label_map or dict(), rzl)
        assert len(got) == N
        for gi, wi in zip(got, want):
            for a, b in zip(gi, wi):
                assert a.dtype == torch.float32 and torch.equal(a, b)
    return conf, pred


@pytest.mark.parametrize("N,C,H,W", [(2, 6, 64, 64), (1, 2, 32, 48), (3, 33, 40, 40), (2, 19, 17, 23),
                                      (1, 6, 5, 7), (2, 64, 16, 16), (1, 1, 8, 8), (3, 8, 24, 20),
                                      (2, 3, 9, 12), (1, 150, 12, 12), (1, 255, 6, 10)])
def test_eval_logits_shapes(cuda, N, C, H, W):
    g = torch.Generator().manual_seed(1234 + C)
    _check(cuda, teacher_logits(N, C, H, W, g), _labels(N, H, W, C, 7 + C), C)


@pytest.mark.parametrize("ldt", [torch.uint8, torch.int32, torch.int64])
@pytest.mark.parametrize("pdt", [torch.uint8, torch.int64])
def test_eval_logits_dtypes(cuda, ldt, pdt):
    g = torch.Generator().manual_seed(3)
    _check(cuda, teacher_logits(2, 6, 48, 40, g), _labels(2, 48, 40, 6, 5, ldt), 6, pred_dtype=pdt)
    # odd plane size -> scalar kernel
    _check(cuda, teacher_logits(2, 6, 15, 15, g), _labels(2, 15, 15, 6, 6, ldt), 6, pred_dtype=pdt)


def test_eval_logits_label_map_reduce_zero_and_out_of_range(cuda):
    g = torch.Generator().manual_seed(11)
    logits = teacher_logits(2, 6, 32, 32, g)
    gt = torch.randint(0, 9, (2, 32, 32), generator=g).to(torch.uint8)   # 6..8 out of range
    gt[:, :3] = 255
    for lm, rz in ((None, False), ({7: 0, 6: 255}, False), (None, True), ({1: 2, 2: 3}, True)):
        _check(cuda, logits, gt, 6, 255, lm, rz)
    gt64 = gt.long()
    gt64[0, 5, :7] = -3
    gt64[1, 6, :7] = 100000
    _check(cuda, logits, gt64, 6)
    _check(cuda, logits, gt, 6, ignore=0)


def test_eval_logits_ties_nan_inf(cuda):
    # exact ties: lowest index wins
    t = torch.zeros((1, 6, 16, 16))
    t[:, 2] = 3.0; t[:, 4] = 3.0
    gt = _labels(1, 16, 16, 6, 1)
    _, pred = _check(cuda, t, gt, 6)
    assert (pred == 2).all()
    # tie decided in softmax space: class 3 has the strictly largest logit, exp(x-max) of class 0
    # rounds to 1.0 -> torch returns 0
    x = torch.zeros((1, 4, 4, 4))
    x[:, 3] = 1e-9
    x[:, 1] = -5.0
    pred_o = om.seg_argmax(x)
    _, pred = ops.argmax_confusion(x.to(cuda), None, return_pred=torch.int64)
    assert torch.equal(pred.cpu(), pred_o)
    # generic (C > 8) kernel, same rule
    x = torch.full((1, 12, 4, 4), -2.0)
    x[:, 9] = 1e-9
    x[:, 4] = 0.0
    pred_o = om.seg_argmax(x)
    _, pred = ops.argmax_confusion(x.to(cuda), None, return_pred=torch.uint8)
    assert torch.equal(pred.cpu().long(), pred_o)
    # NaN / inf propagate through the softmax: index 0
    g = torch.Generator().manual_seed(9)
    for C in (6, 12):
        x = torch.randn((1, C, 8, 8), generator=g)
        x[0, 3, 0, 0] = float("nan")
        x[0, 0, 0, 1] = float("inf")
        x[0, 4, 0, 2] = float("inf")
        x[0, :, 0, 3] = float("-inf")
        x[0, 5, 0, 4] = float("-inf")
        pred_o = om.seg_argmax(x)
        _, pred = ops.argmax_confusion(x.to(cuda), None, return_pred=torch.int64)
        assert torch.equal(pred.cpu(), pred_o)


def test_eval_logits_pure_noise_and_adversarial_labels(cuda):
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, 6, 64, 64), generator=g)
    gt = torch.randint(0, 6, (3, 64, 64), generator=g).to(torch.uint8)
    gt[torch.rand((3, 64, 64), generator=g) < 0.03] = 255
    _check(cuda, x, gt, 6)
    x = torch.randn((2, 33, 30, 30), generator=g)
    gt = torch.randint(0, 33, (2, 30, 30), generator=g).to(torch.uint8)
    _check(cuda, x, gt, 33)


def test_eval_logits_matches_two_kernel_path_and_meter(cuda):
    """fused kernel == pseudo-label arg-max kernel + confusion kernel, bit for bit, at cfg5's map
    size; accumulation into a running matrix; arg-max only mode; empty batch."""
    g = torch.Generator().manual_seed(1234)
    logits = teacher_logits(4, 6, 1024, 1024, g).to(cuda)
    gt = _labels(4, 1024, 1024, 6, 99).to(cuda)
    conf, pred = ops.argmax_confusion(logits, gt, 6, per_image=True, return_pred=torch.int64)
    lab, _, _, _ = ops.pseudo_label(logits, 0.0)
    assert torch.equal(pred, lab)
    assert torch.equal(conf, ops.confusion_accum(lab, gt, 6, per_image=True))
    assert torch.equal(M.seg_argmax(logits, torch.uint8).long(), lab)
    # size-independent invariants
    c = conf.cpu().numpy()
    g_np = gt.cpu().numpy()
    for i in range(4):
        assert c[i].sum() == int((g_np[i] != 255).sum())
        assert np.array_equal(c[i, :6, :6].sum(1), np.bincount(g_np[i][g_np[i] != 255], minlength=6))
    meter = M.ConfusionMeter(6, device=cuda)
    meter.update_logits(logits[:2], gt[:2])
    meter.update_logits(logits[2:], gt[2:])
    assert torch.equal(meter.conf[0], conf.sum(0))
    z, _ = ops.argmax_confusion(torch.zeros((0, 6, 8, 8), device=cuda),
                                torch.zeros((0, 8, 8), dtype=torch.uint8, device=cuda), 6)
    assert int(z.sum()) == 0
    with pytest.raises(ValueError):
        ops.argmax_confusion(logits, None)
    with pytest.raises(ValueError):
        ops.argmax_confusion(logits, gt[:2], 6)
