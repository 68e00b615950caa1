"""CUDA path vs the committed golden fixtures (outputs of the REFERENCE code, see
tests/golden/make_golden.py). Integer results bit-exact; floats within north_star's
tolerances. Nothing here reads /root/reference."""
from pathlib import Path

import numpy as np
import pytest
import torch

from pfst_b200 import ops
from pfst_b200.evaluation import metrics as M
from pfst_b200.losses.pfgst_loss import PFGSTLoss, LOSS_KEYS
from pfst_b200.utils import dacs_transforms as T

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
W6 = {'src_pos': 0.1, 'src_neg': 0.1, 'sim_pos': 0.1, 'sim_neg': 0.1, 'src_pos_std': 0.1, 'src_neg_std': 0.1}


def test_metrics_golden(cuda):
    z = np.load(G / "metrics.npz")
    C = int(z["C"])
    ret = M.eval_metrics(z["pred"], z["label"], C, 255, metrics=["mIoU", "mDice", "mFscore"])
    for k, v in ret.items():
        assert np.array_equal(np.asarray(v), z["all_" + k], equal_nan=True), k
    per = [M.intersect_and_union(z["pred"][i], z["label"][i], C, 255) for i in range(10)]
    assert np.array_equal(np.stack([np.stack([a.numpy() for a in p]) for p in per]), z["per_image"])
    assert np.array_equal(M.pre_eval_to_metrics(per, ["mIoU"])["IoU"], z["pre_eval_IoU"])
    a = M.intersect_and_union(z["lm_pred"], z["lm_label"], 6, 255, {7: 0, 6: 255}, True)
    assert np.array_equal(np.stack([t.numpy() for t in a]), z["lm_areas"])


def test_ema_golden(cuda):
    z = np.load(G / "ema.npz")
    n = 6
    student = [torch.from_numpy(z[f"student_{i}"].copy()).to(cuda) for i in range(n)]
    teacher = [torch.from_numpy(z[f"teacher0_{i}"].copy()).to(cuda) for i in range(n)]
    table = ops.EmaTable(teacher, student)
    for it in z["iters"]:
        table.update(*ops.ema_coeffs(int(it), 0.999))
        for i in range(n):
            assert np.array_equal(teacher[i].cpu().numpy(), z[f"teacher_it{it}_{i}"]), (it, i)
    table.update(0.0, 1.0, mode=1)
    for i in range(n):
        assert np.array_equal(teacher[i].cpu().numpy(), z[f"teacher_init_{i}"])


def test_pseudo_and_mix_golden(cuda):
    z = np.load(G / "pseudo_mix.npz")
    logits = torch.from_numpy(z["logits"]).to(cuda)
    gt = torch.from_numpy(z["gt"]).long().to(cuda)
    thr, top, bottom = float(z["thr"]), int(z["top"]), int(z["bottom"])
    label, conf, count, _ = ops.pseudo_label(logits, thr)
    assert np.array_equal(label.cpu().numpy(), z["pseudo_label"])
    assert np.allclose(conf.cpu().numpy(), z["pseudo_prob"], rtol=0, atol=1e-6)
    safe = np.abs(z["pseudo_prob"] - np.float32(thr)) > 1e-6
    large = conf.cpu().numpy() >= np.float32(thr)
    assert np.array_equal(large[safe], z["large"][safe])
    np.random.seed(int(z["seed"]))
    plan = T.ClassMixPlan(cuda, max_batch=gt.shape[0])
    plan.start(gt)
    chosen = plan.choose()
    mi, ml, mw, mm = T.class_mix_batch(torch.from_numpy(z["img"]).to(cuda), torch.from_numpy(z["trg"]).to(cuda),
                                       gt, label, chosen, count=count, ps_size=label.numel(),
                                       ignore_top=top, ignore_bottom=bottom)
    assert np.array_equal(mm.cpu().numpy(), z["mix_masks"])
    assert np.array_equal(ml.cpu().numpy(), z["mixed_lbl"])
    assert np.array_equal(mi.cpu().numpy(), z["mixed_img"])
    if np.array_equal(large, z["large"]):          # same count -> same ratio -> bit-exact weights
        assert np.array_equal(mw.cpu().numpy(), z["mixed_weight"])
    else:
        assert np.allclose(mw.cpu().numpy(), z["mixed_weight"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("name", ["a", "b"])
def test_pfgst_loss_golden(cuda, name):
    z = np.load(G / "pfgst_loss.npz")
    dil, down = z[f"{name}_cfg"]
    lt = torch.from_numpy(z[f"{name}_logits"]).to(cuda).requires_grad_(True)
    xs = torch.from_numpy(z[f"{name}_x_src"]).to(cuda).requires_grad_(True)
    mod = PFGSTLoss(top_k=3, dilation=int(dil), kernel_size=3, weights=W6, sim_type='cosine', feat_level=None,
                    detach_unfold=True, downscale=None if down < 0 else float(down))
    res = mod(dict(logits_trg=lt, gt_src=torch.from_numpy(z[f"{name}_gt"]).long().to(cuda),
                   x_ema=torch.from_numpy(z[f"{name}_x_ema"]).to(cuda), x_src=xs, img_trg=None,
                   mix_masks=torch.from_numpy(z[f"{name}_mix"]).long().to(cuda)))
    sum(res[k] for k in LOSS_KEYS).backward()
    got = np.array([float(res[k].detach()) for k in LOSS_KEYS])
    want = z[f"{name}_losses"].astype(np.float64)
    assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 1e-9), (got, want)
    for key, g in (("grad_x_src", xs.grad), ("grad_logits", lt.grad)):
        ref = z[f"{name}_{key}"]
        assert np.abs(g.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max(), key
    assert np.array_equal(res['vis|density_sim_feat'][2].cpu().numpy(), z[f"{name}_eroded"])
    assert np.allclose(res['vis|density_sim_feat'][1].cpu().numpy(), z[f"{name}_density"], rtol=0, atol=2e-6)


def test_eval_logits_golden(cuda):
    """fused arg-max + confusion kernel vs the fixture written by the reference's own
    inference / simple_test / intersect_and_union on the same synthetic logits."""
    from tests.golden.make_golden import eval_logits_cases, eval_logits_inputs
    z = np.load(G / "eval_logits.npz")
    for name, N, C, H, W, lm, rz in eval_logits_cases():
        logits, gt = eval_logits_inputs(name, N, C, H, W)
        top2 = torch.softmax(logits, 1).topk(2, dim=1).values
        assert not bool(((top2[:, 0] - top2[:, 1]) < 1e-6).any()), "fixture holds a near tie"
        pred = M.seg_argmax(logits.to(cuda), torch.uint8).cpu().numpy()
        assert np.array_equal(pred, z[name + "_pred"]), name
        per = M.pre_eval_logits(logits.to(cuda), gt.to(cuda), C, 255, lm, rz)
        assert np.array_equal(np.stack([np.stack([a.numpy() for a in t]) for t in per]), z[name + "_areas"]), name
