"""StrongAugmentation on the GPU (transforms.py:1062-1145) vs the oracle, which is
pinned bit-exactly against cv2 4.13 over every colour and against the reference class compiled from
its source (tests/test_oracle_pins.py), and vs the golden fixture written by that reference class.
uint8 work: every comparison is bit-exact."""
import itertools
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import strong_aug as osa
from pfst_b200 import ops
from pfst_b200.pipelines import StrongAugmentation
from pfst_b200.registry import PIPELINES

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _run(cuda, imgs, op_lists, simd=32):
    d = torch.from_numpy(np.stack(imgs)).to(cuda)
    return ops.photometric_u8(d, op_lists, simd).cpu().numpy()


@pytest.mark.parametrize("H,W", [(64, 64), (33, 47), (120, 120), (16, 100), (7, 5), (96, 256)])
def test_every_single_distortion_and_the_pairs(cuda, H, W):
    rs = np.random.RandomState(H * 1000 + W)
    img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    img[0, :, :] = 0
    img[1, :, :] = 255
    img[2, :, :] = rs.randint(0, 256, (W, 1))                 # grey pixels: S == 0 branch
    singles = [(osa.OP_CONVERT, 1, 17.3), (osa.OP_CONVERT, 1, -31.9), (osa.OP_CONVERT, 1.47, 0),
               (osa.OP_CONVERT, 0.52, 0), (osa.OP_SATURATION, 1.5, 0), (osa.OP_SATURATION, 0.5, 0),
               (osa.OP_SATURATION, 1.0, 0), (osa.OP_HUE, 18, 0), (osa.OP_HUE, -18, 0), (osa.OP_HUE, 0, 0)]
    lists = [[s] for s in singles] + [list(p) for p in itertools.permutations(singles[1:9:2], 2)] + [[]]
    got = _run(cuda, [img] * len(lists), lists)
    for i, ol in enumerate(lists):
        assert np.array_equal(got[i], osa.apply_strong_aug(img.copy(), ol)), ol


def test_random_draws_match_the_oracle_and_the_simd_rule(cuda):
    rs = np.random.RandomState(7)
    aug = StrongAugmentation()
    for simd in (32, 16, 0):
        imgs, lists = [], []
        for i in range(40):
            imgs.append(rs.randint(0, 256, (48, 120, 3)).astype(np.uint8))
            lists.append(aug.draw(rs))
        got = _run(cuda, imgs, lists, simd)
        for i in range(40):
            assert np.array_equal(got[i], osa.apply_strong_aug(imgs[i].copy(), lists[i], simd)), (simd, lists[i])


def test_golden_written_by_the_reference_class(cuda):
    from tests.golden.make_golden import strong_aug_cases, strong_aug_image
    z = np.load(G / "strong_aug.npz")
    aug = PIPELINES.build(dict(type="StrongAugmentation"))
    for seed, H, W in strong_aug_cases():
        np.random.seed(seed + 1000)
        res = aug(dict(img=strong_aug_image(seed, H, W), img_fields=['img']))
        assert np.random.random() == float(z[f"next_{seed}"]), seed          # same numpy stream consumption
        assert res['img_fields'] == ['img', 'img_strong_aug']
        assert res['img_strong_aug'].dtype == np.uint8
        assert np.array_equal(res['img_strong_aug'], z[f"out_{seed}"]), seed


def test_cv2_directly_when_present(cuda):
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(11)
    img = rs.randint(0, 256, (64, 96, 3)).astype(np.uint8)
    got = _run(cuda, [img, img], [[(osa.OP_SATURATION, 1.3, 0)], [(osa.OP_HUE, 11, 0)]])
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    hsv[:, :, 1] = osa.convert_u8(hsv[:, :, 1], alpha=1.3)
    assert np.array_equal(got[0], cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    hsv[:, :, 0] = (hsv[:, :, 0].astype(int) + 11) % 180
    assert np.array_equal(got[1], cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))


def test_batches_in_place_device_tensors_and_errors(cuda):
    rs = np.random.RandomState(3)
    aug = StrongAugmentation()
    imgs = [rs.randint(0, 256, (20, 36, 3)).astype(np.uint8) for _ in range(70)]      # > 64 images: two launches
    lists = [aug.draw(rs) for _ in range(70)]
    d = torch.from_numpy(np.stack(imgs)).to(cuda)
    got = aug.apply_batch(d, lists)
    for i in (0, 1, 63, 64, 69):
        assert np.array_equal(got[i].cpu().numpy(), osa.apply_strong_aug(imgs[i].copy(), lists[i]))
    same = aug.apply_batch(d, lists, out=d)                                            # in place
    assert same.data_ptr() == d.data_ptr() and torch.equal(d, got)
    odd = torch.from_numpy(np.stack(imgs[:3])).to(cuda)[:, :19, :35].contiguous()       # H*W % 4 != 0: byte path
    g2 = ops.photometric_u8(odd, lists[:3]).cpu().numpy()
    for i in range(3):
        assert np.array_equal(g2[i], osa.apply_strong_aug(imgs[i][:19, :35].copy(), lists[i]))
    np.random.seed(2)
    res = aug(dict(img=torch.from_numpy(imgs[5]).to(cuda), img_fields=['img']))        # device tensor in and out
    np.random.seed(2)
    want = osa.apply_strong_aug(imgs[5].copy(), osa.draw_strong_aug(np.random))
    assert res['img_strong_aug'].is_cuda and np.array_equal(res['img_strong_aug'].cpu().numpy(), want)
    with pytest.raises(ops.PfstError):
        ops.photometric_u8(torch.zeros((1, 4, 4, 3), dtype=torch.uint8), [[]])          # CPU tensor
    with pytest.raises(ops.PfstError):
        ops.photometric_u8(d[:1], [[(9, 1.0, 0)]])                                      # unknown op code
    with pytest.raises(ops.PfstError):
        ops.photometric_u8(d[:1], [[(osa.OP_HUE, 2.5, 0)]])                             # hue delta must be an integer
    with pytest.raises(ValueError):
        ops.photometric_u8(d[:2], [[]])
    with pytest.raises(ValueError):
        ops.photometric_u8(d[:1], [[(osa.OP_HUE, 1, 0)] * 5])
