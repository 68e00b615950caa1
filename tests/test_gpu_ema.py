"""E1/E2 parity: CUDA multi-tensor EMA vs the oracle (pfgst.py:105-127). Bit-exact."""
import pytest
import torch

from oracle import ema as oema
from pfst_b200 import ops
from pfst_b200.synthetic import model_params

pytestmark = pytest.mark.gpu


def _ragged_params(gen):
    # scalar tensor, tiny vectors, odd sizes (scalar tail), exactly one chunk, > one chunk
    shapes = [(), (1,), (3,), (6,), (513,), (4096,), (4097,), (64, 3, 3, 3), (8191,), (70001,)]
    return [0.02 * torch.randn(s, generator=gen) for s in shapes]


@pytest.mark.parametrize("it", [1, 2, 10, 999, 5000])
def test_ema_bit_exact_ragged(cuda, it):
    g = torch.Generator().manual_seed(1234)
    student, teacher = _ragged_params(g), _ragged_params(g)
    t_dev = [t.clone().to(cuda) for t in teacher]
    s_dev = [s.clone().to(cuda) for s in student]
    oema.ema_update(teacher, student, it, 0.999)
    table = ops.EmaTable(t_dev, s_dev)
    table.update(*ops.ema_coeffs(it, 0.999))
    for i, (a, b) in enumerate(zip(t_dev, teacher)):
        assert torch.equal(a.cpu(), b), f"tensor {i} differs at iter {it}"


def test_ema_unaligned_views(cuda):
    g = torch.Generator().manual_seed(7)
    base_t, base_s = torch.randn(10007, generator=g), torch.randn(10007, generator=g)
    dt, ds = base_t.to(cuda), base_s.to(cuda)
    views = lambda x: [x[1:4098], x[4099:4100], x[5001:10007]]  # 4-byte aligned only
    ref_t = [v.clone() for v in views(base_t)]
    oema.ema_update(ref_t, views(base_s), 77, 0.999)
    ops.EmaTable(views(dt), views(ds)).update(*ops.ema_coeffs(77, 0.999))
    for a, b in zip(views(dt), ref_t):
        assert torch.equal(a.cpu(), b)


def test_ema_init_copy_is_bitwise(cuda):
    g = torch.Generator().manual_seed(3)
    student = _ragged_params(g)
    student[3][0] = float("nan"); student[4][1] = float("inf"); student[4][2] = -0.0
    s_dev = [s.to(cuda) for s in student]
    t_dev = [torch.full_like(s, 7.0) for s in s_dev]
    ops.EmaTable(t_dev, s_dev).update(0.0, 1.0, mode=1)
    for a, b in zip(t_dev, student):
        assert torch.equal(a.cpu().view(torch.int32) if a.dim() else a.cpu().reshape(1).view(torch.int32),
                           b.view(torch.int32) if b.dim() else b.reshape(1).view(torch.int32))


def test_ema_full_model_three_steps(cuda):
    """DeepLabV3+ R50-D8 parameter list (214 tensors, 43.58 M fp32), iters 1,2,3 chained."""
    g = torch.Generator().manual_seed(1234)
    student, teacher = model_params(6, g), model_params(6, g)
    assert len(student) == 214 and sum(p.numel() for p in student) == 43579868
    t_dev, s_dev = [t.to(cuda) for t in teacher], [s.to(cuda) for s in student]
    table = ops.EmaTable(t_dev, s_dev)
    for it in (1, 2, 3):
        oema.ema_update(teacher, student, it, 0.999)
        table.update(*ops.ema_coeffs(it, 0.999))
    bad = sum(int((a.cpu() != b).sum()) for a, b in zip(t_dev, teacher))
    assert bad == 0


def test_ema_flat(cuda):
    g = torch.Generator().manual_seed(11)
    e, p = torch.randn(1000003, generator=g), torch.randn(1000003, generator=g)
    de, dp = e.to(cuda), p.to(cuda)
    ref = [e.clone()]
    oema.ema_update(ref, [p], 5000, 0.999)
    ops.ema_update_flat(de, dp, *ops.ema_coeffs(5000, 0.999))
    assert torch.equal(de.cpu(), ref[0])


def test_ema_device_coefficients_and_bounded_grid(cuda):
    """pfst_ema_update_multi_dev (coefficients read from device memory: graph-capturable) and the
    bounded persistent grid give the same bits as the host-argument launch."""
    g = torch.Generator().manual_seed(11)
    shapes = [(3,), (4097,), (64, 3, 3, 3), (1,), (100003,)]
    stu = [(0.02 * torch.randn(s, generator=g)).to(cuda) for s in shapes]
    tea_a = [(0.02 * torch.randn(s, generator=g)).to(cuda) for s in shapes]
    tea_b = [t.clone() for t in tea_a]
    tea_c = [t.clone() for t in tea_a]
    a32, b32 = ops.ema_coeffs(37, 0.999)
    ops.EmaTable(tea_a, stu).update(a32, b32)
    ops.EmaTable(tea_b, stu).update(a32, b32, blocks_per_sm=2)
    coefs = torch.tensor([a32, b32], dtype=torch.float32, device=cuda)
    ops.EmaTable(tea_c, stu).update_dev(coefs, blocks_per_sm=3)
    for x, y, z in zip(tea_a, tea_b, tea_c):
        assert torch.equal(x, y) and torch.equal(x, z)
