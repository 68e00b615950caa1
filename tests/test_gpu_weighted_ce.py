"""Decode-head loss: the fused up-sample + weighted CE + accuracy +
gradient kernel against the CPU oracle and the golden vectors produced by the reference's
own resize / cross_entropy / accuracy (tests/golden/weighted_ce.npz). fp32: 1e-5 relative."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import weighted_ce as OC
from pfst_b200 import ops
from pfst_b200.losses import decode_head_losses, upsample_cross_entropy
from tests.golden.make_golden import ce_cases, ce_inputs

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
TOL = 1e-5


def _close(a, b, what):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert (a - b).abs().max() <= TOL * b.abs().max() + 1e-12, (what, float((a - b).abs().max()), float(b.abs().max()))


@pytest.mark.parametrize("case", ce_cases(), ids=[c[0] for c in ce_cases()])
def test_matches_reference_golden_and_oracle(cuda, case):
    name, B, C, lh, lw, scale, use_w, use_cw, lwt = case
    z = np.load(G / "weighted_ce.npz")
    logits, label, weight, cw = ce_inputs(name, B, C, lh, lw, scale, use_w, use_cw)
    zl = logits.to(cuda).requires_grad_(True)
    loss, acc = upsample_cross_entropy(zl, label.to(cuda), None if weight is None else weight.to(cuda),
                                       None if cw is None else cw.to(cuda), 255, lwt)
    loss.backward()
    _close(loss.detach(), z[f"{name}_loss"], "loss vs reference")
    _close(acc, z[f"{name}_acc"], "acc vs reference")
    _close(zl.grad, z[f"{name}_grad"], "grad vs reference")
    zo = logits.clone().requires_grad_(True)
    lo, ao, _ = OC.decode_head_losses(zo, label, weight, cw, 255, lwt)
    lo.backward()
    _close(loss.detach(), lo.detach(), "loss vs oracle")
    _close(acc, ao, "acc vs oracle")
    _close(zl.grad, zo.grad, "grad vs oracle")


def test_full_size_cfg2_and_upstream_gradient(cuda):
    g = torch.Generator().manual_seed(1)
    B, C, lh, H = 8, 6, 128, 512
    logits = 2.0 * torch.randn((B, C, lh, lh), generator=g)
    label = torch.randint(0, C, (B, 1, H, H), generator=g)
    label[:, :, :9] = 255
    weight = torch.rand((B, H, H), generator=g)
    zl = logits.to(cuda).requires_grad_(True)
    out = decode_head_losses(zl, label.to(cuda), weight.to(cuda), loss_weight=0.4)
    (3.0 * out["loss_ce"]).backward()
    zo = logits.clone().requires_grad_(True)
    lo, ao, _ = OC.decode_head_losses(zo, label, weight, None, 255, 0.4)
    (3.0 * lo).backward()
    _close(out["loss_ce"].detach(), lo.detach(), "loss")
    _close(out["acc_seg"], ao, "acc")
    _close(zl.grad, zo.grad, "grad")


def test_all_ignored_and_no_grad(cuda):
    logits = torch.randn((1, 4, 8, 8)).to(cuda)
    label = torch.full((1, 1, 32, 32), 255, dtype=torch.int64).to(cuda)
    loss, acc = upsample_cross_entropy(logits, label)
    lo, ao, _ = OC.decode_head_losses(logits.cpu(), label.cpu())
    assert float(loss) == 0.0 == float(lo)
    _close(acc, ao, "acc of an all-ignored map")


def test_unsupported_resampling_raises(cuda):
    logits = torch.randn((1, 4, 8, 8)).to(cuda)
    with pytest.raises(ops.PfstError):
        upsample_cross_entropy(logits, torch.zeros((1, 1, 30, 32), dtype=torch.int64).to(cuda))
    with pytest.raises(ops.PfstError):
        decode_head_losses(logits, torch.zeros((1, 1, 32, 32), dtype=torch.int64).to(cuda), align_corners=True)
