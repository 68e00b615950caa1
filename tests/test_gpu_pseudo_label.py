"""S1/S2 parity: fused softmax/argmax/threshold kernel vs the oracle (pfgst.py:259-277).
Labels bit-exact; confident mask bit-exact except pixels within 1e-6 of the threshold
(north_star); confidences within 1e-6 absolute (CPU Sleef expf vs CUDA expf)."""
import numpy as np
import pytest
import torch

from oracle import pseudo as opl
from pfst_b200 import ops
from pfst_b200.synthetic import teacher_logits

pytestmark = pytest.mark.gpu
THR = 0.98


def _check(cuda, logits, thr=THR):
    lab_o, prob_o, large_o = opl.pseudo_label(logits, thr)
    lab, conf, count, _ = ops.pseudo_label(logits.to(cuda), thr)
    lab, conf = lab.cpu(), conf.cpu()
    near_tie = torch.zeros_like(large_o)
    # oracle-side near-ties between the top two probabilities are the only pixels where a
    # 1-ulp expf difference may legitimately move the arg-max
    if logits.shape[1] > 1:
        top2 = torch.softmax(logits, 1).topk(2, dim=1).values
        near_tie = (top2[:, 0] - top2[:, 1]) < 1e-6
    assert torch.equal(lab[~near_tie], lab_o[~near_tie])
    finite = torch.isfinite(prob_o)
    assert torch.allclose(conf[finite], prob_o[finite], rtol=0, atol=1e-6)
    assert torch.equal(torch.isnan(conf), torch.isnan(prob_o))
    large = conf.ge(thr)
    safe = (prob_o - np.float32(thr)).abs() > 1e-6
    safe &= finite
    assert torch.equal(large[safe], large_o[safe])
    # the device-side count is the count of the kernel's own mask
    assert int(count.cpu()) == int(large.sum())
    return lab, conf, count


@pytest.mark.parametrize("B,C,H,W", [(2, 6, 64, 64), (1, 2, 32, 48), (3, 33, 40, 40), (2, 19, 17, 23),
                                      (1, 6, 5, 7), (2, 64, 16, 16), (1, 1, 8, 8)])
def test_pseudo_label_shapes(cuda, B, C, H, W):
    g = torch.Generator().manual_seed(1234)
    _check(cuda, teacher_logits(B, C, H, W, g))


def test_pseudo_label_pure_noise_and_ties(cuda):
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 6, 32, 32), generator=g)
    _check(cuda, x)
    # exact-tie planes: lowest index must win (torch.max semantics)
    t = torch.zeros((1, 6, 16, 16))
    t[:, 2] = 3.0; t[:, 4] = 3.0
    lab, conf, _ = _check(cuda, t)
    assert (lab == 2).all()
    allsame = torch.full((1, 5, 8, 8), -1.25)
    lab, conf, _ = _check(cuda, allsame)
    assert (lab == 0).all() and torch.allclose(conf, torch.full_like(conf, 0.2))


def test_pseudo_label_softmax_space_ties(cuda):
    # logits so close that exp(x-max) rounds to 1.0: the tie is decided in softmax space
    x = torch.zeros((1, 4, 4, 4))
    x[:, 3] = 1e-9            # strictly largest logit, but softmax ties with class 0
    x[:, 1] = -5.0
    lab_o, _, _ = opl.pseudo_label(x, THR)
    lab, _, _, _ = ops.pseudo_label(x.to(cuda), THR)
    assert torch.equal(lab.cpu(), lab_o)


def test_pseudo_label_nan_inf(cuda):
    g = torch.Generator().manual_seed(9)
    x = torch.randn((1, 6, 8, 8), generator=g)
    x[0, 3, 0, 0] = float("nan")
    x[0, 0, 0, 1] = float("inf")
    x[0, :, 0, 2] = float("-inf")
    x[0, 5, 0, 3] = float("-inf")
    lab_o, prob_o, large_o = opl.pseudo_label(x, THR)
    lab, conf, count, _ = ops.pseudo_label(x.to(cuda), THR)
    assert torch.equal(lab.cpu(), lab_o)
    assert torch.equal(torch.isnan(conf.cpu()), torch.isnan(prob_o))
    assert int(count.cpu()) == int(conf.cpu().ge(THR).sum())


def test_pseudo_weight_all_and_part(cuda):
    g = torch.Generator().manual_seed(1234)
    x = teacher_logits(2, 6, 64, 64, g)
    _, prob_o, large_o = opl.pseudo_label(x, THR)
    lab, conf, count, wpart = ops.pseudo_label(x.to(cuda), THR, want_part_weight=True)
    large = conf.cpu().ge(THR)
    assert torch.equal(wpart.cpu(), large.float())
    for top, bottom in ((0, 0), (5, 0), (0, 7), (3, 4)):
        w = ops.pseudo_weight_fill((2, 64, 64), count, 2 * 64 * 64, top, bottom).cpu()
        w_o = opl.pseudo_weight(large, "all", top, bottom)   # same mask -> must be bit-exact
        assert torch.equal(w, w_o)


def test_pseudo_label_classwise_threshold(cuda):
    g = torch.Generator().manual_seed(21)
    x = teacher_logits(2, 6, 32, 32, g)
    thr = torch.tensor([0.5, 0.9, 0.98, 0.6, 0.99, 0.7])
    lab_o, prob_o, large_o = opl.pseudo_label_classwise(x, thr)
    lab, conf, count, w = ops.pseudo_label(x.to(cuda), 0.0, thr_per_class=thr.to(cuda), want_part_weight=True)
    assert torch.equal(lab.cpu(), lab_o)
    safe = (prob_o - thr[lab_o]).abs() > 1e-6
    assert torch.equal(w.cpu().bool()[safe], large_o[safe])


def test_pseudo_label_entropy_mode(cuda):
    """offline class-wise rule, loading.py:474-487 (entropy < thr[pred], else 255)."""
    g = torch.Generator().manual_seed(22)
    x = teacher_logits(1, 6, 32, 32, g)
    thr = np.array([0.05, 0.2, 0.1, 0.3, 0.02, 0.15], dtype=np.float32)
    lab_o, ent, keep = opl.entropy_label(x[0].numpy(), thr)
    lab, conf, count, _ = ops.pseudo_label(x.to(cuda), 0.0, thr_per_class=torch.from_numpy(thr).to(cuda),
                                           mode=1, reject_label=255)
    safe = np.abs(ent - thr[x[0].numpy().argmax(0)]) > 1e-5
    assert np.array_equal(lab.cpu().numpy()[0][safe], lab_o[safe])


def test_pseudo_label_against_torch_cuda(cuda):
    """Secondary check: the same torch ops the reference runs, executed on the GPU."""
    g = torch.Generator().manual_seed(1234)
    x = teacher_logits(4, 6, 128, 128, g).to(cuda)
    lab_t, prob_t, large_t = opl.pseudo_label(x, THR)
    lab, conf, count, _ = ops.pseudo_label(x, THR)
    assert torch.equal(lab, lab_t)
    assert torch.allclose(conf, prob_t, rtol=0, atol=1e-6)
    safe = (prob_t - THR).abs() > 1e-6
    assert torch.equal(conf.ge(THR)[safe], large_t[safe])


def test_exp_split_is_bit_identical(cuda):
    """The kernel's hand-scheduled exp must equal CUDA expf bit for bit on d = x - max <= 0."""
    from pfst_b200 import _lib
    g = torch.Generator().manual_seed(0)
    xs = [-torch.rand(4_000_000, generator=g) * 110.0,            # whole useful range incl. underflow
          -torch.rand(2_000_000, generator=g) * 1e-3,              # near zero
          -torch.logspace(-45, 2.1, 200_000),                      # denormals .. -126
          torch.tensor([0.0, -0.0, float("-inf"), float("nan"), -87.3, -88.7, -103.9, -104.1, -126.0])]
    x = torch.cat(xs).to(cuda)
    bad = torch.zeros(1, dtype=torch.int64, device=cuda)
    _lib.call("pfst_selftest_exp", x.data_ptr(), x.numel(), bad.data_ptr(), ops._stream())
    assert int(bad.cpu()) == 0


def test_pseudo_label_bitwise_vs_torch_cuda_report(cuda):
    """Informational strictness: on the GPU the confidences should equal torch's own
    softmax->max bit for bit (same expf, same summation order); allow 1e-6 as specified."""
    g = torch.Generator().manual_seed(77)
    x = teacher_logits(2, 6, 256, 256, g).to(cuda)
    _, prob_t, _ = opl.pseudo_label(x, THR)
    _, conf, _, _ = ops.pseudo_label(x, THR)
    frac_equal = float((conf == prob_t).float().mean())
    print(f"bitwise-equal confidences vs torch CUDA: {frac_equal:.6f}")
    assert frac_equal > 0.99
