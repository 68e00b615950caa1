"""V1-V4 parity: confusion-matrix kernel and the metrics drop-in vs the oracle
(metrics.py:26-86, 296-395; tools/confusion_matrix.py:46-65). Integer-exact."""
import numpy as np
import pytest
import torch

from oracle import metrics as om
from pfst_b200 import ops
from pfst_b200.evaluation import metrics as M
from pfst_b200.synthetic import blocky_labels, eval_maps

pytestmark = pytest.mark.gpu


def _legacy_iou(results, gts, C, ignore):
    tot = np.zeros((C, C), dtype=np.float64)
    for r, g in zip(results, gts):
        tot += om.confusion(r, g, C, ignore)
    d = np.diag(tot)
    return d.sum() / tot.sum(), d / tot.sum(1), d / (tot.sum(1) + tot.sum(0) - d)


def test_reference_golden_test_metrics(cuda):
    """tests/test_metrics.py:86-143 of the reference, re-targeted at the CUDA drop-in."""
    rs = np.random.RandomState(0)
    C, ignore = 19, 255
    results = rs.randint(0, C, size=(10, 30, 30))
    label = rs.randint(0, C, size=(10, 30, 30))
    label[:, 2, 5:10] = ignore
    ret = M.eval_metrics(results, label, C, ignore, metrics=["mIoU", "mDice", "mFscore"])
    all_acc, acc, iou = _legacy_iou(results, label, C, ignore)
    assert ret["aAcc"] == all_acc
    assert np.allclose(ret["Acc"], acc) and np.allclose(ret["IoU"], iou)
    want = om.eval_metrics(results, label, C, ignore, metrics=["mIoU", "mDice", "mFscore"])
    for k in want:
        assert np.array_equal(ret[k], want[k], equal_nan=True), k
    # absent classes -> nan_to_num (test_metrics.py:145-203)
    results = rs.randint(0, 5, size=(10, 30, 30))
    label = rs.randint(0, 4, size=(10, 30, 30))
    ret = M.eval_metrics(results, label, C, ignore_index=255, metrics=["mDice", "mIoU", "mFscore"], nan_to_num=-1)
    for k in ("Acc", "IoU", "Dice", "Precision", "Recall", "Fscore"):
        assert ret[k][-1] == -1
    # histc 59-class regression (test_metrics.py:205-216)
    ret = M.eval_metrics(np.array([np.repeat(31, 59)]), np.array([np.arange(59)]), 59, 255, metrics="mIoU")
    assert not np.any(np.isnan(ret["IoU"]))


@pytest.mark.parametrize("C", [2, 6, 8, 9, 19, 33, 150, 255])
@pytest.mark.parametrize("pdt,ldt", [(np.int64, np.uint8), (np.uint8, np.uint8), (np.int64, np.int64),
                                     (np.int32, np.uint8)])
def test_intersect_and_union_dtypes_and_classes(cuda, C, pdt, ldt):
    rs = np.random.RandomState(C)
    hi = min(C + 3, 256) if pdt == np.uint8 else C + 3   # some predictions out of range
    pred = rs.randint(0, hi, size=(37, 53)).astype(pdt)
    lab = rs.randint(0, min(C + 2, 255), size=(37, 53)).astype(ldt)  # some labels out of range
    lab[rs.random_sample(lab.shape) < 0.05] = 255
    want = om.areas(pred.astype(np.int64), lab, C, 255)
    got = M.intersect_and_union(pred, lab, C, 255)
    for a, b in zip(got, want):
        assert a.dtype == torch.float32 and torch.equal(a, b)


def test_label_map_and_reduce_zero_label(cuda):
    rs = np.random.RandomState(3)
    pred = rs.randint(0, 6, size=(64, 64)).astype(np.int64)
    lab = rs.randint(0, 8, size=(64, 64)).astype(np.uint8)
    lab[:4] = 255
    for lm, rz in (({7: 0, 6: 255}, False), ({}, True), ({1: 2, 2: 3}, True)):
        want = om.areas(pred, lab, 6, 255, lm, rz)
        got = M.intersect_and_union(pred, lab, 6, 255, lm, rz)
        for a, b in zip(got, want):
            assert torch.equal(a, b), (lm, rz)


def test_confusion_matrix_integer_exact(cuda):
    for C, shape in ((6, (3, 128, 128)), (2, (2, 100, 60)), (33, (4, 30, 30))):
        pred, gt = eval_maps(shape[0], shape[1], shape[2], C, seed=C)
        want = om.confusion(pred, gt, C, 255)
        got = M.confusion_matrix(pred, gt, C, 255).cpu().numpy()
        assert got.dtype == np.int64 and np.array_equal(got, want)
    g = torch.Generator().manual_seed(1)
    gt = blocky_labels(2, 256, 256, 6, g)[:, 0]
    pred = blocky_labels(2, 256, 256, 6, g, ignore_frac=0)[:, 0]
    want = om.confusion(pred.numpy(), gt.numpy(), 6, 255)
    assert np.array_equal(M.confusion_matrix(pred, gt, 6, 255).cpu().numpy(), want)


def test_per_image_batch_and_pre_eval(cuda):
    pred, gt = eval_maps(7, 96, 80, 6, seed=11)
    batch = M.intersect_and_union_batch(pred, gt, 6, 255).cpu()
    per_image = []
    for i in range(7):
        want = om.areas(pred[i], gt[i], 6, 255)
        for k in range(4):
            assert torch.equal(batch[i, k].float(), want[k])
        per_image.append(M.intersect_and_union(pred[i], gt[i], 6, 255))
    got = M.pre_eval_to_metrics(per_image, ["mIoU", "mFscore"])
    want = om.metrics_from_areas(*om.pre_eval_sum([om.areas(pred[i], gt[i], 6, 255) for i in range(7)]),
                                 ["mIoU", "mFscore"])
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True)


def test_ragged_and_empty(cuda):
    pred = np.array([[1, 2, 3]], dtype=np.int64)
    lab = np.array([[1, 255, 3]], dtype=np.uint8)
    got = M.intersect_and_union(pred, lab, 4, 255)
    want = om.areas(pred, lab, 4, 255)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    z = ops.confusion_accum(torch.zeros((0, 4, 4), dtype=torch.int64, device=cuda),
                            torch.zeros((0, 4, 4), dtype=torch.uint8, device=cuda), 3)
    assert int(z.sum()) == 0
    # odd pixel counts / unaligned views go through the scalar path
    pred, gt = eval_maps(3, 33, 31, 6, seed=2)
    want = sum(om.confusion(pred[i], gt[i], 6, 255) for i in range(3))
    assert np.array_equal(M.confusion_matrix(pred, gt, 6, 255).cpu().numpy(), want)
    p = torch.from_numpy(pred).to(cuda).reshape(-1)[1:]
    l = torch.from_numpy(gt).to(cuda).reshape(-1)[1:]
    c = ops.confusion_accum(p.reshape(1, -1), l.reshape(1, -1), 6)[0, :6, :6].cpu().numpy()
    assert np.array_equal(c, om.confusion(pred.reshape(-1)[1:], gt.reshape(-1)[1:], 6, 255))


def test_full_size_maps_checksum(cuda):
    """BASELINE configs[4] shape (1024x1024 maps): totals must satisfy the size-independent
    invariants sum(conf) = #non-ignored pixels, row sums = label histogram."""
    pred, gt = eval_maps(8, 1024, 1024, 6, seed=1234)
    conf = ops.confusion_accum(torch.from_numpy(pred).to(cuda), torch.from_numpy(gt).to(cuda), 6,
                               per_image=True).cpu().numpy()
    assert conf.shape == (8, 7, 7)
    for i in range(8):
        assert conf[i].sum() == int((gt[i] != 255).sum())
        assert np.array_equal(conf[i, :6, :6].sum(1), np.bincount(gt[i][gt[i] != 255], minlength=6))
        assert np.array_equal(conf[i, :6, :6], om.confusion(pred[i], gt[i], 6, 255))
    meter = M.ConfusionMeter(6, device=cuda)
    meter.update(torch.from_numpy(pred).to(cuda), torch.from_numpy(gt).to(cuda))
    assert np.array_equal(meter.matrix().cpu().numpy(), conf[:, :6, :6].sum(0))
