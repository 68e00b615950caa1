"""Whole-step integration: the PFGST drop-in (pfst_b200.uda.PFGST) against three
train_step iterations of the REFERENCE PFGST on the same tiny segmentor and batches
(golden: tests/golden/pfgst_step.npz). The two runs differ by cuDNN-vs-CPU convolution
rounding in the segmentor passes, so scalars are compared at 2e-3 relative here; the
exact parity of every hot-path op is covered by the per-op tests."""
import random
from pathlib import Path

import numpy as np
import pytest
import torch

from pfst_b200.uda import PFGST
from pfst_b200 import ops
from tests.fake_segmentor import TinySegmentor
from tests.golden.make_golden import STEP_CFG, step_batches

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


def _build(cuda, **extra):
    cfg = dict(STEP_CFG)
    cfg.update(extra)
    cfg['model'] = lambda: TinySegmentor(6, 16, seed=0)
    return PFGST(**cfg).to(cuda)


def _to(batch, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def test_three_train_steps_match_reference(cuda):
    z = np.load(G / "pfgst_step.npz")
    m = _build(cuda)
    opt = torch.optim.SGD(m.model.parameters(), lr=0.01)
    random.seed(1); np.random.seed(1); torch.manual_seed(1)
    for it, batch in enumerate(step_batches()):
        out = m.train_step(_to(batch, cuda), opt)
        keys = list(out['log_vars'].keys())
        assert keys == list(z[f"log_keys_{it}"]), (keys, z[f"log_keys_{it}"])
        got = np.array(list(out['log_vars'].values()))
        want = z[f"log_vals_{it}"]
        assert np.all(np.abs(got - want) <= 2e-3 * np.abs(want) + 2e-4), (it, dict(zip(keys, zip(got, want))))
        assert out['num_samples'] == 2
        assert {'vis|density_sim_feat', 'vis|seg_mask_src', 'vis|seg_mask_mix'} <= set(out['states'])
    ema = torch.cat([p.detach().reshape(-1) for p in m.ema_model.parameters()]).cpu().numpy()
    stu = torch.cat([p.detach().reshape(-1) for p in m.model.parameters()]).cpu().numpy()
    assert np.abs(stu - z["student_params"]).max() <= 2e-3
    assert np.abs(ema - z["ema_params"]).max() <= 2e-3
    assert m.local_iter == 3
    assert all(not p.requires_grad for p in m.get_ema_model().parameters())


def test_ema_of_the_module_is_bit_exact_given_identical_students(cuda):
    from oracle import ema as oema
    m = _build(cuda)
    m._init_ema_weights()
    assert all(torch.equal(a, b) for a, b in zip(m.get_ema_model().parameters(), m.get_model().parameters()))
    teacher = [p.detach().cpu().clone() for p in m.get_ema_model().parameters()]
    with torch.no_grad():
        for p in m.get_model().parameters():
            p.add_(0.01 * torch.randn_like(p))
    student = [p.detach().cpu().clone() for p in m.get_model().parameters()]
    for it in (1, 2, 3, 50):
        m._update_ema(it)
        oema.ema_update(teacher, student, it, 0.999)
        for a, b in zip(m.get_ema_model().parameters(), teacher):
            assert torch.equal(a.detach().cpu(), b)


def test_kornia_branches_fail_loudly_or_skip(cuda):
    batch = _to(next(iter(step_batches(1))), cuda)
    m = _build(cuda, blur=True, color_jitter_probability=0.0, kornia_aug='error')
    with pytest.raises(ops.PfstError):
        for _ in range(8):                      # blur is active with probability 1/2 per draw
            m.forward_train(**batch)
    m = _build(cuda, blur=True, color_jitter_probability=0.0, kornia_aug='skip')
    with pytest.warns(UserWarning):
        m.forward_train(**batch)


def test_blur_branch_consumes_the_numpy_stream_like_the_reference(cuda, monkeypatch):
    """blur=True, jitter disabled: per iteration the reference draws B class choices (get_class_masks)
    and then, inside the mixing loop, one sigma per image (dacs_transforms.py:93). The drop-in must
    leave the global numpy stream in exactly that state, and the mixed image must be blurred."""
    batch = _to(next(iter(step_batches(1))), cuda)
    m = _build(cuda, blur=True, color_jitter_probability=1.0)
    monkeypatch.setattr(random, "uniform", lambda a, b: 0.9)      # color_jitter=0.9 <= p, blur=0.9 > 0.5
    seen = {}
    orig = ops.gaussian_blur
    monkeypatch.setattr(ops, "gaussian_blur", lambda data, sigmas, *a, **k: seen.setdefault(
        "out", (list(sigmas), orig(data, sigmas, *a, **k)))[1])
    np.random.seed(17)
    m.forward_train(**batch)
    after = np.random.random()
    rs = np.random.RandomState(17)
    B = batch['img'].shape[0]
    n = int(torch.unique(batch['gt_semantic_seg']).numel())
    for _ in range(B):
        rs.choice(n, int((n + n % 2) / 2), replace=False)
    sig = [rs.uniform(0.15, 1.15) for _ in range(B)]
    assert after == rs.random_sample()
    assert seen["out"][0] == sig and tuple(seen["out"][1].shape) == tuple(batch['img'].shape)


def test_part_threshold_and_prototype_extension(cuda):
    batch = _to(next(iter(step_batches(1))), cuda)
    m = _build(cuda, thre_type='part', prototypes=dict(weight=0.1, conf_threshold=0.5),
               pseudo_threshold_per_class=[0.5, 0.6, 0.7, 0.5, 0.6, 0.7])
    random.seed(0); np.random.seed(0)
    log_vars, _ = m.forward_train(**batch)
    assert 'loss_proto_dist' in log_vars and np.isfinite(log_vars['loss_proto_dist'])
    log_vars, _ = m.forward_train(**batch)
    assert np.isfinite(log_vars['loss'])
    assert int(m.proto_bank.seen.sum()) >= 1 and m.proto_bank.iter == 2


def _run_steps(m, cuda, n_iters, seed=3):
    opt = torch.optim.SGD(m.model.parameters(), lr=0.01)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    logs = []
    batches = list(step_batches(2))
    for it in range(n_iters):
        out = m.train_step(_to(batches[it % 2], cuda), opt)
        logs.append(out['log_vars'])
    torch.cuda.synchronize()
    params = torch.cat([p.detach().reshape(-1) for p in m.model.parameters()]).cpu()
    ema = torch.cat([p.detach().reshape(-1) for p in m.ema_model.parameters()]).cpu()
    return logs, params, ema


@pytest.mark.parametrize("protos", [False, True])
def test_fused_launch_groups_equal_the_generic_module_path(cuda, protos, monkeypatch):
    """The engine's fused groups (CUDA-graph replay from the third pass on) against the generic
    path — PFGSTLoss autograd module + proto_dist_loss, every kernel launched eagerly: same log
    variables, same student and teacher after 6 optimizer steps."""
    extra = dict(prototypes=dict(weight=0.1)) if protos else {}
    res = {}
    for name, fused, graphs in (("fused", True, "1"), ("fused_eager", True, "0"), ("generic", False, "0")):
        monkeypatch.setenv("PFST_PLUGIN_GRAPHS", graphs)
        m = _build(cuda, fused_hot_path=fused, **extra)
        res[name] = _run_steps(m, cuda, 6)
        if name == "fused":
            assert m._engine._gb.graphs or m._engine._gb.misses > 2     # replay (or moving inputs: eager)
        m.close()
    keys = list(res["generic"][0][0].keys())
    assert 'loss_sim_pos' in keys and ('loss_proto_dist' in keys) == protos
    for other in ("fused", "fused_eager"):
        for it, (a, b) in enumerate(zip(res[other][0], res["generic"][0])):
            assert list(a.keys()) == list(b.keys())
            for k in a:
                assert abs(float(a[k]) - float(b[k])) <= 2e-5 * abs(float(b[k])) + 1e-7, (other, it, k, float(a[k]), float(b[k]))
        assert (res[other][1] - res["generic"][1]).abs().max() <= 1e-5
        assert (res[other][2] - res["generic"][2]).abs().max() <= 1e-6
    # graph replay vs eager launches of the same groups: identical kernels on identical inputs
    # (the tiny segmentor's cuDNN passes are not bit-reproducible run to run, so not an exact comparison)
    for a, b in zip(res["fused"][0], res["fused_eager"][0]):
        for k in a:
            assert abs(float(a[k]) - float(b[k])) <= 2e-5 * abs(float(b[k])) + 1e-7, k


def test_log_vars_are_lazy_and_equal_the_reference_arithmetic(cuda):
    """_parse_losses on device scalars: no host sync until a value is read; `loss` is the
    left-to-right fp32 sum of the 'loss' keys (base.py:200-202); values equal the tensors."""
    from pfst_b200.uda.log_ledger import LazyScalar, ledger_for
    from pfst_b200.uda.uda_decorator import UDADecorator
    g = torch.Generator().manual_seed(5)
    vals = torch.randn(5, generator=g)
    losses = {"decode.loss_ce": vals[0].to(cuda).requires_grad_(True), "decode.acc_seg": vals[1].to(cuda),
              "aux.loss_ce": (vals[2:4].to(cuda)).requires_grad_(True), "loss_list": [vals[4].to(cuda), vals[0].to(cuda)]}
    led = ledger_for(cuda)
    before = led.d2h_bytes
    loss, lv = UDADecorator._parse_losses(losses)
    assert led.d2h_bytes == before and all(isinstance(v, LazyScalar) for v in lv.values())
    want = {"decode.loss_ce": vals[0], "decode.acc_seg": vals[1], "aux.loss_ce": vals[2:4].mean(),
            "loss_list": vals[4] + vals[0]}
    ref_loss = sum(v for k, v in want.items() if 'loss' in k)
    assert list(lv) == list(want) + ["loss"]
    for k, v in want.items():
        assert float(lv[k]) == float(v), k
    assert float(lv["loss"]) == float(ref_loss) == float(loss)
    assert led.d2h_bytes > before
    loss.backward()
    assert float(losses["decode.loss_ce"].grad) == 1.0
    assert torch.equal(losses["aux.loss_ce"].grad.cpu(), torch.tensor([0.5, 0.5]))
    tot = led.weighted_total([loss.detach(), losses["decode.acc_seg"]], [1.0, 0.25])
    assert float(tot) == float(torch.tensor(0.) + ref_loss + vals[1] * 0.25)
