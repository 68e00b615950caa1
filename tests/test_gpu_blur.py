"""Strong-augmentation Gaussian blur (dacs_transforms.py:88-107) vs the oracle
restatement of kornia's GaussianBlur2d (oracle/strong_aug.py — third-party arithmetic, PARITY
UNPINNED: kornia is neither vendored nor version-pinned by the reference).

Tolerance: |gpu - oracle| <= 1e-5 * max|input| (north_star's 1e-5 relative for fp32 quantities;
a blur output can cancel to ~0, so the scale is the input's), against the fp32 oracle AND an fp64
evaluation of the same formula."""
import numpy as np
import pytest
import torch

from oracle import strong_aug as osa
from pfst_b200 import ops
from pfst_b200.utils import dacs_transforms as T

pytestmark = pytest.mark.gpu


def _check(cuda, x, sigmas, ksize=None):
    N, C, H, W = x.shape
    ks = ksize or (osa.kernel_size(H), osa.kernel_size(W))
    got = ops.gaussian_blur(x.to(cuda), sigmas, ksize).cpu()
    tol = 1e-5 * float(x.abs().max())
    for i, s in enumerate(sigmas):
        w32 = osa.gaussian_blur2d(x[i:i + 1], ks, (s, s))
        w64 = osa.gaussian_blur2d(x[i:i + 1], ks, (s, s), torch.float64)
        assert float((got[i:i + 1] - w32).abs().max()) <= tol, (i, s)
        assert float((got[i:i + 1].double() - w64).abs().max()) <= tol, (i, s)
    return got


def test_kernel_size_rule():
    assert [ops.blur_kernel_size(n) for n in (120, 512, 1024, 100, 33, 16, 24)] == \
           [osa.kernel_size(n) for n in (120, 512, 1024, 100, 33, 16, 24)] == [11, 51, 103, 9, 3, 1, 3]


@pytest.mark.parametrize("N,C,H,W", [(2, 3, 128, 128), (1, 3, 512, 512), (2, 3, 100, 76), (1, 3, 33, 47),
                                      (3, 3, 120, 120), (1, 1, 64, 64), (2, 4, 70, 130)])
def test_blur_matches_oracle(cuda, N, C, H, W):
    g = torch.Generator().manual_seed(H * 7 + W)
    x = torch.randn((N, C, H, W), generator=g)
    sig = [0.15, 1.15, 0.6][:N]
    _check(cuda, x, sig)


def test_blur_reference_sigma_range_and_image_statistics(cuda):
    # de-normalised image-like data (large offsets): the tolerance scales with the input
    g = torch.Generator().manual_seed(3)
    x = torch.rand((4, 3, 256, 256), generator=g) * 255.0
    rs = np.random.RandomState(0)
    _check(cuda, x, [osa.draw_sigma(rs) for _ in range(4)])


def test_blur_cfg2_batch_at_the_largest_reference_sigma(cuda):
    # B=8 x 512^2, sigma at the top of the reference range: 15 of 51 taps, ~45 KB of shared memory
    # (the size at which the launch needs the opt-in above the 48 KB default)
    g = torch.Generator().manual_seed(8)
    x = torch.randn((8, 3, 512, 512), generator=g)
    sig = [1.15, 0.9, 0.15, 0.5, 1.0, 0.3, 0.75, 1.149]
    got = ops.gaussian_blur(x.to(cuda), sig).cpu()
    tol = 1e-5 * float(x.abs().max())
    for i in (0, 2, 7):
        want = osa.gaussian_blur2d(x[i:i + 1], (51, 51), (sig[i], sig[i]))
        assert float((got[i:i + 1] - want).abs().max()) <= tol, i


def test_blur_wide_sigma_uses_every_tap(cuda):
    # sigma large against the kernel: no tap is below the 2^-30 cut-off, reflect border fully used
    g = torch.Generator().manual_seed(4)
    x = torch.randn((2, 3, 96, 80), generator=g)
    _check(cuda, x, [3.0, 10.0], ksize=(21, 31))
    _check(cuda, x, [2.0, 0.5], ksize=(1, 9))
    _check(cuda, x, [40.0, 25.0], ksize=(95, 79))     # kernel almost as large as the image


def test_blur_properties(cuda):
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, 3, 192, 160), generator=g).to(cuda)
    sig = [0.3, 0.7, 1.1]
    y = ops.gaussian_blur(x, sig)
    # partition of unity: constants are preserved
    c = torch.full((3, 3, 192, 160), 2.5, device=cuda)
    assert float((ops.gaussian_blur(c, sig) - 2.5).abs().max()) <= 1e-5
    # sigma -> 0: identity (all weight on the centre tap)
    assert torch.allclose(ops.gaussian_blur(x, [1e-3] * 3), x, rtol=0, atol=1e-6)
    # linearity
    x2 = torch.randn((3, 3, 192, 160), generator=g).to(cuda)
    lhs = ops.gaussian_blur(2.0 * x + x2, sig)
    assert torch.allclose(lhs, 2.0 * y + ops.gaussian_blur(x2, sig), rtol=0, atol=2e-5)
    # images are independent: a batch equals its images blurred one by one, bit for bit
    for i in range(3):
        assert torch.equal(ops.gaussian_blur(x[i:i + 1].contiguous(), sig[i:i + 1]), y[i:i + 1])
    # mean is preserved up to the reflect border's effect on a flat-ish image; max never grows
    assert float(y.abs().max()) <= float(x.abs().max()) + 1e-5
    # more than 64 images: several launches
    big = torch.randn((70, 3, 24, 24), generator=g)
    sg = [0.2 + 0.01 * i for i in range(70)]
    got = ops.gaussian_blur(big.to(cuda), sg).cpu()
    for i in (0, 63, 64, 69):
        want = osa.gaussian_blur2d(big[i:i + 1], (3, 3), (sg[i], sg[i]))
        assert torch.allclose(got[i:i + 1], want, rtol=0, atol=1e-5 * float(big.abs().max()))


def test_blur_batch_helper_and_errors(cuda):
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 3, 64, 64), generator=g)
    np.random.seed(9)
    y = T.gaussian_blur_batch(0.8, x.to(cuda)).cpu()
    rs = np.random.RandomState(9)
    sig = [rs.uniform(0.15, 1.15) for _ in range(2)]
    assert np.random.random() == rs.random_sample()
    for i in range(2):
        want = osa.gaussian_blur2d(x[i:i + 1], (7, 7), (sig[i], sig[i]))
        assert torch.allclose(y[i:i + 1], want, rtol=0, atol=1e-5 * float(x.abs().max()))
    xc = x.to(cuda)
    assert T.gaussian_blur_batch(0.3, xc) is xc                     # inactive: untouched, no draw
    with pytest.raises(ops.PfstError):
        ops.gaussian_blur(x, [0.5, 0.5])                            # CPU tensor
    with pytest.raises(ops.PfstError):
        ops.gaussian_blur(xc, [0.5, 0.5], (4, 5))                   # even kernel
    with pytest.raises(ops.PfstError):
        ops.gaussian_blur(xc, [0.5, 0.5], (129, 5))                 # reflect pad >= size
    with pytest.raises(ops.PfstError):
        ops.gaussian_blur(xc, [0.5, -1.0])                          # bad sigma
    with pytest.raises(ValueError):
        ops.gaussian_blur(xc, [0.5])                                # one sigma per image
    with pytest.raises(ValueError):
        ops.gaussian_blur(xc, [0.5, 0.5], out=xc)                   # in place
