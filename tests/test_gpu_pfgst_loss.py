"""L1-L6 parity: PFGSTLoss drop-in (four CUDA kernels) vs the oracle, which is
bit-identical to the reference module (tests/test_oracle_pins.py). Losses within 1e-5
relative (north_star); gradients within 1e-5 of their scale; masks bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mixing as omix, pfgst_loss as OL
from pfst_b200 import ops
from pfst_b200.losses.pfgst_loss import PFGSTLoss, LOSS_KEYS
from pfst_b200.synthetic import WORKLOADS, step_inputs

pytestmark = pytest.mark.gpu
W6 = {'src_pos': 0.1, 'src_neg': 0.1, 'sim_pos': 0.1, 'sim_neg': 0.1, 'src_pos_std': 0.1, 'src_neg_std': 0.1}


def _dots_ref(x, d):
    """five maps: |x|^2 and the four forward dot products, zero padding."""
    B, D, h, w = x.shape
    xp = F.pad(x, (d, d, d, d))
    c = xp[:, :, d:d + h, d:d + w]
    sh = lambda dy, dx: xp[:, :, d + dy:d + dy + h, d + dx:d + dx + w]
    return torch.stack([(c * c).sum(1), (c * sh(0, d)).sum(1), (c * sh(d, -d)).sum(1),
                        (c * sh(d, 0)).sum(1), (c * sh(d, d)).sum(1)], dim=1)


@pytest.mark.parametrize("B,D,h,w,d", [(2, 32, 16, 16, 2), (1, 512, 64, 64, 2), (2, 24, 20, 36, 1),
                                        (1, 16, 40, 64, 4), (3, 16, 15, 15, 1), (2, 13, 9, 22, 2),
                                        (1, 8, 4, 4, 2)])
def test_neigh_dots(cuda, B, D, h, w, d):
    g = torch.Generator().manual_seed(1)
    xa = torch.relu(torch.randn((B, D, h, w), generator=g))
    xb = torch.randn((B, D, h, w), generator=g)
    dots, ks = ops.neigh_dots(xa.to(cuda), xb.to(cuda), d)
    got = dots.sum(0).cpu()
    for t, x in enumerate((xa, xb)):
        want = _dots_ref(x.double(), d)
        err = (got[t].double() - want).abs().max() / want.abs().max()
        assert err < 2e-6, (t, float(err))
    one, _ = ops.neigh_dots(xa.to(cuda), None, d)
    assert torch.allclose(one.sum(0)[0].cpu(), got[0], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("B,D,h,w,d", [(2, 32, 16, 16, 2), (1, 512, 64, 64, 2), (2, 24, 20, 36, 1),
                                        (1, 16, 40, 64, 4), (3, 16, 15, 15, 1), (2, 13, 9, 22, 2)])
def test_neigh_grad(cuda, B, D, h, w, d):
    g = torch.Generator().manual_seed(2)
    x = torch.randn((B, D, h, w), generator=g)
    coef = torch.randn((B, 9, h, w), generator=g)
    got = ops.neigh_grad(x.to(cuda), coef.to(cuda), d).cpu()
    xp = F.pad(x.double(), (d, d, d, d))
    want = torch.zeros_like(x, dtype=torch.float64)
    for k in range(9):
        dy, dx = (k // 3 - 1) * d, (k % 3 - 1) * d
        want += coef[:, k:k + 1].double() * xp[:, :, d + dy:d + dy + h, d + dx:d + dx + w]
    assert (got.double() - want).abs().max() < 1e-5 * want.abs().max()


def _run_case(cuda, wl, seed=1234, mask_seed=3):
    inp = step_inputs(wl, seed)
    np.random.seed(mask_seed)
    mix = torch.cat(omix.class_masks(inp['gt']), 0)
    downscale = wl.downscale if wl.downscale != 1.0 else None

    def tensors(dev):
        return dict(logits_trg=inp['logits_trg'].clone().to(dev).requires_grad_(True), logits_ema=None,
                    gt_src=inp['gt'].to(dev), x_ema=inp['x_ema'].to(dev),
                    x_src=inp['x_src'].clone().to(dev).requires_grad_(True), img_trg=inp['img'].to(dev),
                    mix_masks=mix.to(dev))

    t_o = tensors('cpu')
    out_o = OL.pfgst_loss(t_o, OL.LossCfg(dilation=wl.dilation, downscale=downscale))
    sum(out_o[k] for k in LOSS_KEYS).backward()

    mod = PFGSTLoss(top_k=3, dilation=wl.dilation, kernel_size=3, weights=W6, sim_type='cosine',
                    feat_level=None, detach_unfold=True, downscale=downscale)
    t_c = tensors(cuda)
    out_c = mod(t_c)
    sum(out_c[k] for k in LOSS_KEYS).backward()
    return out_o, t_o, out_c, t_c


@pytest.mark.parametrize("name", ["tiny", "tiny33", "cfg1"])
def test_pfgst_loss_matches_oracle(cuda, name):
    out_o, t_o, out_c, t_c = _run_case(cuda, WORKLOADS[name])
    for k in LOSS_KEYS:
        a, b = float(out_c[k]), float(out_o[k])
        assert out_c[k].dim() == 0 and out_c[k].requires_grad
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-9, (k, a, b)
    # vis tuple: density map and the eroded target mask (bit-exact)
    _, dens_o, er_o = out_o['vis|density_sim_feat']
    _, dens_c, er_c = out_c['vis|density_sim_feat']
    assert torch.equal(er_c.cpu(), er_o) and er_c.dtype == torch.bool
    assert torch.allclose(dens_c.cpu(), dens_o, rtol=0, atol=2e-6)
    for key in ('x_src', 'logits_trg'):
        go, gc = t_o[key].grad, t_c[key].grad.cpu()
        assert gc.shape == go.shape
        scale = go.abs().max()
        assert (gc - go).abs().max() <= 1e-5 * scale + 1e-12, (key, float((gc - go).abs().max()), float(scale))
    # the reference sends no gradient to x_ema
    assert t_c['x_ema'].grad is None


def test_pfgst_loss_upstream_grad_weights(cuda):
    """backward must honour arbitrary upstream gradients of the six scalars."""
    wl = WORKLOADS["tiny"]
    inp = step_inputs(wl)
    np.random.seed(3)
    mix = torch.cat(omix.class_masks(inp['gt']), 0)
    coeffs = [0.3, -1.2, 2.0, 0.7, 1.5, -0.4]

    def run(dev, fn):
        t = dict(logits_trg=inp['logits_trg'].clone().to(dev).requires_grad_(True), gt_src=inp['gt'].to(dev),
                 x_ema=inp['x_ema'].to(dev), x_src=inp['x_src'].clone().to(dev).requires_grad_(True),
                 img_trg=None, mix_masks=mix.to(dev))
        out = fn(t)
        sum(c * out[k] for c, k in zip(coeffs, LOSS_KEYS)).backward()
        return t

    t_o = run('cpu', lambda t: OL.pfgst_loss(t, OL.LossCfg()))
    mod = PFGSTLoss(top_k=3, dilation=2, kernel_size=3, weights=W6, sim_type='cosine', feat_level=None,
                    detach_unfold=True, downscale=0.5)
    t_c = run(cuda, mod)
    for key in ('x_src', 'logits_trg'):
        go, gc = t_o[key].grad, t_c[key].grad.cpu()
        assert (gc - go).abs().max() <= 1e-5 * go.abs().max()


def test_pfgst_loss_empty_target_region(cuda):
    """`if ignore_mask.sum() > 1` (pfgst_loss.py:227): with no target pixel both sim losses are 0."""
    wl = WORKLOADS["tiny"]
    inp = step_inputs(wl)
    mix = torch.ones((wl.B, 1, wl.H, wl.W), dtype=torch.int64)
    mod = PFGSTLoss(top_k=3, dilation=2, kernel_size=3, weights=W6, sim_type='cosine', feat_level=None,
                    detach_unfold=True, downscale=0.5)
    t = dict(logits_trg=inp['logits_trg'].to(cuda).requires_grad_(True), gt_src=inp['gt'].to(cuda),
             x_ema=inp['x_ema'].to(cuda), x_src=inp['x_src'].to(cuda).requires_grad_(True), img_trg=None,
             mix_masks=mix.to(cuda))
    out = mod(t)
    assert float(out['loss_sim_pos']) == 0.0 and float(out['loss_sim_neg']) == 0.0
    sum(out[k] for k in LOSS_KEYS).backward()
    assert float(t['logits_trg'].grad.abs().max()) == 0.0
    t_o = dict(logits_trg=inp['logits_trg'], gt_src=inp['gt'], x_ema=inp['x_ema'], x_src=inp['x_src'],
               img_trg=None, mix_masks=mix)
    out_o = OL.pfgst_loss(t_o, OL.LossCfg())
    for k in LOSS_KEYS[:4]:
        assert abs(float(out[k]) - float(out_o[k])) <= 1e-5 * abs(float(out_o[k]))


def test_unsupported_options_raise():
    for kw in (dict(src_loss_type='hinge'), dict(kernel_size=5), dict(src_perc=0.5), dict(top_k=7),
               dict(src_loss_type='margin', margin=[1.5, 0.5]),
               dict(proj_net_cfg=dict(in_channels=4, out_channels=4)), dict(cross_prob_type='src')):
        args = dict(top_k=3, dilation=2, kernel_size=3, weights=W6, sim_type='cosine', feat_level=None,
                    detach_unfold=True, downscale=0.5)
        args.update(kw)
        with pytest.raises(ops.PfstError):
            PFGSTLoss(**args)


@pytest.mark.parametrize("name", ["gauss", "gauss33", "ema", "unfold", "unfold33", "margin", "margin2", "topk_none"])
def test_pfgst_loss_options_match_reference_golden_and_oracle(cuda, name):
    """The PFGSTLoss options outside the shipped configuration (pfgst_loss.py:16-18): sim_type='gaussian',
    cross_prob_type='ema', detach_unfold=False — CUDA path vs the fixture the reference module wrote
    (tests/golden/pfgst_loss_options.npz) and vs the oracle, 1e-5 of the gradient scale."""
    from pathlib import Path
    from tests.golden.make_golden import LOSS_OPTION_CASES, loss_option_inputs, loss_option_keys
    z = np.load(Path(__file__).resolve().parent / "golden" / "pfgst_loss_options.npz")
    c, opts = LOSS_OPTION_CASES[name]
    KEYS = loss_option_keys(opts)
    gt, logits, x_src, x_ema, logits_ema = loss_option_inputs(c)
    mix = torch.from_numpy(z[f"{name}_mix"]).long()
    kw = dict(top_k=3)
    kw.update(opts)
    mod = PFGSTLoss(dilation=c["dil"], kernel_size=3, weights=W6, feat_level=None, downscale=c["down"], **kw)
    assert not mod.shipped_branch
    t = dict(logits_trg=logits.clone().to(cuda).requires_grad_(True), logits_ema=logits_ema.to(cuda),
             gt_src=gt.to(cuda), x_ema=x_ema.to(cuda), x_src=x_src.clone().to(cuda).requires_grad_(True),
             img_trg=None, mix_masks=mix.to(cuda))
    out = mod(t)
    assert set(k for k in out if k.startswith("loss_")) == set(KEYS)
    sum(out[k] for k in KEYS).backward()
    cfg = OL.LossCfg(dilation=c["dil"], downscale=c["down"], sim_type=opts.get("sim_type", "cosine"),
                     sigma=opts.get("sigma", 30.0), cross_prob_type=opts.get("cross_prob_type", "trg"),
                     detach_unfold=opts.get("detach_unfold", True), top_k=opts.get("top_k", 3),
                     src_loss_type=opts.get("src_loss_type", "mean_std"),
                     margin=tuple(opts.get("margin", (0.5, 0.5))))
    to = dict(logits_trg=logits.clone().requires_grad_(True), logits_ema=logits_ema, gt_src=gt, x_ema=x_ema,
              x_src=x_src.clone().requires_grad_(True), img_trg=None, mix_masks=mix)
    oo = OL.pfgst_loss(to, cfg)
    sum(oo[k] for k in KEYS).backward()
    for i, k in enumerate(KEYS):
        a = float(out[k])
        for b in (float(z[f"{name}_losses"][i]), float(oo[k])):
            assert abs(a - b) <= 1e-5 * abs(b) + 1e-9, (k, a, b)
    assert np.array_equal(out['vis|density_sim_feat'][2].cpu().numpy(), z[f"{name}_eroded"])
    assert np.allclose(out['vis|density_sim_feat'][1].cpu().numpy(), z[f"{name}_density"], rtol=0, atol=2e-6)
    for key, ref, ora in (("x_src", z[f"{name}_grad_x_src"], to["x_src"].grad),
                          ("logits_trg", z[f"{name}_grad_logits"], to["logits_trg"].grad)):
        g = t[key].grad.cpu()
        for want in (torch.from_numpy(ref), ora):
            scale = want.abs().max()
            assert (g - want).abs().max() <= 1e-5 * scale + 1e-12, (key, float((g - want).abs().max()), float(scale))


def test_pfgst_loss_option_errors(cuda):
    with pytest.raises(ValueError):
        PFGSTLoss(top_k=3, dilation=2, kernel_size=3, weights=W6, sim_type='l2', detach_unfold=True)
    mod = PFGSTLoss(top_k=3, dilation=2, kernel_size=3, weights=W6, sim_type='cosine', feat_level=None,
                    cross_prob_type='ema', detach_unfold=True, downscale=0.5)
    inp = step_inputs(WORKLOADS["tiny"])
    t = dict(logits_trg=inp['logits_trg'].to(cuda), logits_ema=inp['logits_trg'].to(cuda), gt_src=inp['gt'].to(cuda),
             x_ema=inp['x_ema'].to(cuda), x_src=inp['x_src'].to(cuda), img_trg=None,
             mix_masks=torch.zeros_like(inp['gt']).to(cuda))
    with pytest.raises(ops.PfstError):        # logits_ema is not on the (down-scaled) loss grid: the reference fails too
        mod(t)
