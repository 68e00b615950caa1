"""M1/M2 parity: ClassMix kernels vs the oracle (dacs_transforms.py:110-144,
pfgst.py:277-300). Masks, labels bit-exact; images / weights bit-exact (m in {0,1})."""
import numpy as np
import pytest
import torch

from oracle import mixing as omix, pseudo as opl
from pfst_b200 import ops
from pfst_b200.synthetic import blocky_labels, teacher_logits
from pfst_b200.utils import dacs_transforms as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,H,W", [(2, 6, 64, 64), (3, 2, 32, 40), (2, 33, 40, 40), (1, 6, 7, 9)])
def test_get_class_masks_matches_reference_rng_stream(cuda, B, C, H, W):
    g = torch.Generator().manual_seed(1234)
    gt = blocky_labels(B, H, W, C, g, min_rect=2, max_rect=max(4, H // 2))
    np.random.seed(77)
    want = omix.class_masks(gt)
    state_after = np.random.get_state()[1][:4].copy()
    np.random.seed(77)
    got = T.get_class_masks(gt.to(cuda))
    assert np.array_equal(np.random.get_state()[1][:4], state_after)  # same number of draws
    assert len(got) == B
    for a, b in zip(got, want):
        assert a.dtype == torch.int64 and tuple(a.shape) == (1, 1, H, W)
        assert torch.equal(a.cpu(), b)


def test_class_presence_bits_and_range_flag(cuda):
    gt = torch.tensor([0, 5, 5, 31, 32, 63, 64, 200, 255, 255, 7], dtype=torch.int64)
    words = ops.class_presence(gt.to(cuda)).cpu().numpy().view(np.uint32)
    present = T._present_classes(words)
    assert present.tolist() == sorted(set(gt.tolist()))
    bad = ops.class_presence(torch.tensor([1, 256, 3], dtype=torch.int64, device=cuda)).cpu().numpy()
    assert bad[8] != 0
    with pytest.raises(ops.PfstError):
        T._present_classes(bad.view(np.uint32))
    neg = ops.class_presence(torch.tensor([-1], dtype=torch.int64, device=cuda)).cpu().numpy()
    assert neg[8] != 0


@pytest.mark.parametrize("B,C,H,W,thre", [(2, 6, 64, 64, "all"), (2, 6, 64, 64, "part"),
                                          (3, 33, 40, 40, "all"), (1, 2, 9, 11, "all")])
def test_fused_mix_matches_reference_loop(cuda, B, C, H, W, thre):
    g = torch.Generator().manual_seed(1234)
    gt = blocky_labels(B, H, W, C, g, min_rect=2, max_rect=max(4, H // 2))
    img = torch.randn((B, 3, H, W), generator=g)
    trg = torch.randn((B, 3, H, W), generator=g)
    img[0, 0, 0, 0] = -0.0
    trg[0, 1, 0, 1] = -0.0
    logits = teacher_logits(B, C, H, W, g)
    top, bottom = 2, 3

    # oracle: pseudo labels -> weights -> masks -> per-image loop
    lab_o, prob_o, large_o = opl.pseudo_label(logits, 0.98)
    # take the confident mask from the CUDA path so that a 1-ulp confidence difference at
    # the threshold cannot leak into this test (that tolerance is test_gpu_pseudo_label's job)
    lab, conf, count, wpart = ops.pseudo_label(logits.to(cuda), 0.98, want_part_weight=True)
    large = conf.cpu().ge(0.98)
    w_o = opl.pseudo_weight(large, thre, top, bottom)
    np.random.seed(5)
    masks = omix.class_masks(gt)
    mi_o, ml_o, mw_o, mm_o = omix.mix_batch(img, trg, gt, lab.cpu(), w_o.clone(), masks)

    np.random.seed(5)
    plan = T.ClassMixPlan(cuda, max_batch=B)
    plan.start(gt.to(cuda))
    chosen = plan.choose()
    mi, ml, mw, mm = T.class_mix_batch(
        img.to(cuda), trg.to(cuda), gt.to(cuda), lab, chosen, count=count, ps_size=B * H * W,
        weight_in=wpart if thre == "part" else None, ignore_top=top, ignore_bottom=bottom)
    assert torch.equal(mm.cpu(), mm_o) and mm.dtype == torch.int64
    assert torch.equal(ml.cpu(), ml_o) and tuple(ml.shape) == (B, 1, H, W)
    assert torch.equal(mi.cpu().view(torch.int32), mi_o.view(torch.int32))  # incl. signed zeros
    assert torch.equal(mw.cpu(), mw_o)


def test_one_mix_and_strong_transform_api(cuda):
    g = torch.Generator().manual_seed(3)
    H, W = 16, 24
    mask = (torch.rand((1, 1, H, W), generator=g) > 0.5).long()
    data = torch.randn((2, 3, H, W), generator=g)
    target = torch.randint(0, 6, (2, H, W), generator=g)
    d_o = omix.mix_pair(mask, data[0], data[1])
    t_o = omix.mix_pair(mask, target[0], target[1])
    d, t = T.one_mix(mask.to(cuda), data.to(cuda), target.to(cuda))
    assert torch.equal(d.cpu(), d_o) and tuple(d.shape) == (1, 3, H, W)
    assert torch.equal(t.cpu(), t_o) and tuple(t.shape) == (1, 1, H, W) and t.dtype == torch.int64
    param = dict(mix=mask.to(cuda), color_jitter=0.1, color_jitter_s=0.2, color_jitter_p=0.2, blur=0,
                 mean=None, std=None, denorm_type="mean_std")
    d2, t2 = T.strong_transform(param, data=data.to(cuda), target=target.to(cuda))
    assert torch.equal(d2.cpu(), d_o) and torch.equal(t2.cpu(), t_o)
    wpair = torch.stack((torch.ones(H, W), torch.full((H, W), 0.37)))
    _, w = T.strong_transform(param, target=wpair.to(cuda))
    assert torch.equal(w.cpu(), omix.mix_pair(mask, wpair[0], wpair[1]))
    # blur branch (dacs_transforms.py:88-107): one sigma from the global numpy stream, then kornia's
    # GaussianBlur2d (restated in oracle/strong_aug.py) on the mixed image
    from oracle import strong_aug as osa
    param["blur"] = 0.9
    np.random.seed(21)
    d3, _ = T.strong_transform(param, data=data.to(cuda))
    after = np.random.random()
    np.random.seed(21)
    want, sigma = osa.gaussian_blur(0.9, d_o)
    assert 0.15 <= sigma <= 1.15 and after == np.random.random()      # same stream consumption
    assert torch.allclose(d3.cpu(), want, rtol=1e-5, atol=1e-5 * float(d_o.abs().max()))
    # the colour-jitter branch is kornia's random sampler + arithmetic: loud failure, never a silent skip
    param["color_jitter"] = 0.5
    with pytest.raises(ops.PfstError):
        T.strong_transform(param, data=data.to(cuda))


def test_generate_class_mask(cuda):
    g = torch.Generator().manual_seed(4)
    label = torch.randint(0, 6, (1, 12, 12), generator=g)
    classes = torch.tensor([1, 4])
    want = label.eq(classes.unsqueeze(1).unsqueeze(2)).sum(0, keepdims=True)
    got = T.generate_class_mask(label.to(cuda), classes.to(cuda))
    assert torch.equal(got.cpu(), want)
