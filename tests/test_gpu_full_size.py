"""BASELINE.json configs at FULL size: one whole hot-path step of every training config
(cfg2 Potsdam->Vaihingen B=8, cfg3 Inria B=4 1024^2, cfg4 SeasonNet B=64 C=33) through the
CUDA-graph schedule against the CPU oracle on the same seeded inputs, plus size-independent
properties (checksums, selection identities, linearity) that do not need the oracle."""
import numpy as np
import pytest
import torch

from oracle import pfgst_loss as OL, step as ostep
from pfst_b200 import ops
from pfst_b200.prototypes import PrototypeBank
from pfst_b200.step import SelfTrainingStep
from pfst_b200.synthetic import WORKLOADS, step_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_full_size_step_matches_oracle(cuda, name):
    wl = WORKLOADS[name]
    host = step_inputs(wl, 1234)
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 3, 3, 3), (64,), (wl.C, 512, 1, 1), (wl.C,), (100003,)]
    student = [0.02 * torch.randn(s, generator=g) for s in shapes]
    teacher = [0.02 * torch.randn(s, generator=g) for s in shapes]
    d_student, d_teacher = [p.to(cuda) for p in student], [p.to(cuda) for p in teacher]
    inp = {k: v.to(cuda) for k, v in host.items()}
    down = wl.downscale if wl.downscale != 1.0 else None
    step = SelfTrainingStep(d_teacher, d_student, wl.C, wl.D, cuda, dilation=wl.dilation, downscale=down,
                            max_batch=max(wl.B, 64), graphs=True)
    cfg = OL.LossCfg(dilation=wl.dilation, downscale=down)
    it = 7
    np.random.seed(11)
    out = step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                   inp["logits_trg"], inp["x_src"], inp["x_ema"])
    ref = ostep.hot_path_step(it, teacher, student, host, wl.C, loss_cfg=cfg, rng=np.random.RandomState(11))
    torch.cuda.synchronize()
    for a, b in zip(d_teacher, teacher):
        assert torch.equal(a.cpu(), b)                                        # EMA bit-exact
    assert torch.equal(out["pseudo_label"].cpu(), ref["pseudo_label"])        # labels bit-exact
    safe = (ref["pseudo_conf"] - np.float32(0.98)).abs() > 1e-6
    assert torch.equal(out["pseudo_conf"].cpu().ge(0.98)[safe], ref["large"][safe])
    assert torch.equal(out["mix_masks"].cpu(), ref["mix_masks"])              # masks bit-exact
    assert torch.equal(out["mixed_lbl"].cpu(), ref["mixed_lbl"])
    assert torch.equal(out["mixed_img"].cpu(), ref["mixed_img"])
    lo, lr = out["losses"].cpu(), ref["losses"]
    assert torch.all((lo - lr).abs() <= 1e-5 * lr.abs() + 1e-9), (lo, lr)
    po, pr = float(out["proto_loss"].cpu()), float(ref["proto_loss"])
    assert abs(po - pr) <= 1e-5 * abs(pr)
    assert (out["mu"].cpu() - ref["mu"]).abs().max() <= 1e-5 * ref["mu"].abs().max()
    for key in ("grad_x_src", "grad_logits_trg"):
        go, gr = out[key].cpu(), ref[key]
        if gr is None:
            gr = torch.zeros_like(go)
        assert (go - gr).abs().max() <= 1e-5 * gr.abs().max() + 1e-12, key


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_full_size_properties(cuda, name):
    wl = WORKLOADS[name]
    inp = {k: v.to(cuda) for k, v in step_inputs(wl, 99).items()}
    # prototypes: the class sums add up to the sum over all valid pixels; counts to their number
    label, conf, count, _ = ops.pseudo_label(inp["ema_logits"], 0.98)
    bank = PrototypeBank(wl.C, wl.D, cuda)
    bank.accumulate(inp["x_ema"], label)
    packed = bank.packed.double().cpu()
    h, w = inp["x_ema"].shape[2:]
    stride = wl.H // h
    lab_f = label[:, ::stride, ::stride][:, :h, :w]
    valid = (lab_f >= 0) & (lab_f < wl.C)
    want = (inp["x_ema"].double() * valid.unsqueeze(1)).sum(dim=(0, 2, 3)).cpu()
    got = packed[:wl.C * wl.D].view(wl.C, wl.D).sum(0)
    assert (got - want).abs().max() <= 1e-5 * want.abs().max()
    assert int(packed[wl.C * wl.D:].sum()) == int(valid.sum())
    # count of confident pixels == number of conf >= thr
    assert int(count) == int((conf >= 0.98).sum())
    # ClassMix is a per-pixel selection: out = mask ? source : target, for images and labels
    chosen = torch.zeros((wl.B, 8), dtype=torch.int32, device=cuda)
    chosen[:, 0] = 0b101
    mimg, mlbl, mw, mask = ops.class_mix(inp["gt"], chosen, inp["img"], inp["target_img_strong_aug"], label,
                                         count=count, ps_size=label.numel())
    m = mask.bool()
    assert torch.equal(m, (inp["gt"] == 0) | (inp["gt"] == 2))
    assert torch.equal(mimg, torch.where(m, inp["img"], inp["target_img_strong_aug"]))
    assert torch.equal(mlbl, torch.where(m, inp["gt"], label.unsqueeze(1)))
    # neighbourhood gradient is linear in its coefficient maps
    x = inp["x_src"]
    g = torch.Generator().manual_seed(5)
    c1 = torch.randn((x.shape[0], 9, h, w), generator=g).to(cuda)
    c2 = torch.randn((x.shape[0], 9, h, w), generator=g).to(cuda)
    fd = 1 if name == "cfg4" else wl.dilation
    g12 = ops.neigh_grad(x, 2.0 * c1 + c2, fd)
    g1, g2 = ops.neigh_grad(x, c1, fd), ops.neigh_grad(x, c2, fd)
    ref = 2.0 * g1 + g2
    assert (g12 - ref).abs().max() <= 2e-5 * ref.abs().max()
