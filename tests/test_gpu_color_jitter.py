"""Strong-augmentation colour jitter (dacs_transforms.py:56-85) vs the oracle restatement of
kornia's 0.6-series ColorJitter (oracle/strong_aug.py — third-party sampler and arithmetic, PARITY UNPINNED).
Tolerance: 1e-5 of the data scale (the kernel repeats the torch expressions operation by operation).

First confirmed on a B200 in capture r1v (profiles/r1v_pytest_color_jitter.log). The module stays opt-in
(`kornia_aug='builtin'`) because its parity against kornia itself cannot be pinned."""
import itertools

import numpy as np
import pytest
import torch

from oracle import strong_aug as osa
from pfst_b200 import ops
from pfst_b200.utils import dacs_transforms as T

pytestmark = pytest.mark.gpu


def _params(factors, order):
    f = torch.tensor(factors, dtype=torch.float32)
    return dict(brightness=f[:, 0], contrast=f[:, 1], saturation=f[:, 2], hue=f[:, 3], order=torch.tensor(order))


@pytest.mark.parametrize("H,W", [(32, 32), (17, 23), (64, 48)])
def test_every_order_matches_the_oracle(cuda, H, W):
    g = torch.Generator().manual_seed(H + W)
    x = torch.rand((2, 3, H, W), generator=g)
    x[0, :, 0, :4] = 0.5                                   # grey pixels (delta == 0)
    x[0, :, 1, :4] = torch.tensor([1.0, 0.0, 0.0]).view(3, 1)
    fac = [(0.85, 1.15, 1.2, 0.17), (1.2, 0.8, 0.8, -0.2)]
    for order in itertools.permutations(range(4)):
        want = osa.apply_jitter(x, _params(fac, list(order)))
        got = ops.color_jitter(x.to(cuda), fac, [list(order)] * 2).cpu()
        assert float((got - want).abs().max()) <= 1e-5, order
    for k in range(4):                                     # single transforms
        want = osa.apply_jitter(x, _params(fac, [k]))
        got = ops.color_jitter(x.to(cuda), fac, [[k]] * 2).cpu()
        assert float((got - want).abs().max()) <= 1e-5, k
    assert torch.equal(ops.color_jitter(x.to(cuda), fac, [[], []]).cpu(), x)


def test_reference_call_site_with_denorm(cuda):
    g = torch.Generator().manual_seed(3)
    mean = torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1)
    std = torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1)
    img = (torch.rand((1, 3, 40, 56), generator=g) * 255 - mean) / std
    for seed in range(6):
        torch.manual_seed(seed)
        want, _, params = osa.color_jitter(0.9, mean, std, img.clone(), None, 0.2, 0.2, 'mean_std')
        after = torch.get_rng_state()
        torch.manual_seed(seed)
        got, _ = T.color_jitter(0.9, mean.to(cuda), std.to(cuda), img.to(cuda), None, 0.2, 0.2, 'mean_std')
        assert torch.equal(after, torch.get_rng_state())              # same torch CPU stream consumption
        assert float((got.cpu() - want).abs().max()) <= 1e-5 * float(img.abs().max()) * 4, seed
    # inactive branch: untouched, no draw
    torch.manual_seed(1)
    before = torch.get_rng_state()
    d = img.to(cuda)
    assert T.color_jitter(0.1, mean, std, d, None, 0.2, 0.2)[0] is d
    assert torch.equal(before, torch.get_rng_state())
    with pytest.raises(ValueError):
        T.color_jitter(0.9, mean, std, d, None, 0.2, 0.2, 'bogus')


def test_strong_transform_builtin_and_batches(cuda):
    g = torch.Generator().manual_seed(5)
    x = torch.rand((70, 3, 12, 20), generator=g)                          # > 64 images, HW % 4 == 0
    fac = [(0.8 + 0.005 * i, 1.0 + 0.002 * i, 1.1, 0.001 * i) for i in range(70)]
    orders = [list(np.random.RandomState(i).permutation(4)) for i in range(70)]
    got = ops.color_jitter(x.to(cuda), fac, orders).cpu()
    for i in (0, 63, 64, 69):
        want = osa.apply_jitter(x[i:i + 1], _params([fac[i]], orders[i]))
        assert float((got[i:i + 1] - want).abs().max()) <= 1e-5
    d = x[:2, :, :11, :19].contiguous().to(cuda)                          # odd plane: scalar path, in place
    want = osa.apply_jitter(d.cpu(), _params(fac[:2], [2, 3, 0, 1]))
    same = ops.color_jitter(d, fac[:2], [[2, 3, 0, 1]] * 2, out=d)
    assert same.data_ptr() == d.data_ptr() and float((d.cpu() - want).abs().max()) <= 1e-5
    mask = (torch.rand((1, 1, 12, 20), generator=g) > 0.5).long().to(cuda)
    param = dict(mix=mask, color_jitter=0.9, color_jitter_s=0.2, color_jitter_p=0.2, blur=0,
                 mean=torch.zeros(1, 3, 1, 1), std=torch.ones(1, 3, 1, 1) * 255, denorm_type="mean_std",
                 kornia_aug="builtin")
    out, _ = T.strong_transform(param, data=x[:2].to(cuda))
    assert tuple(out.shape) == (1, 3, 12, 20) and bool(torch.isfinite(out).all())
