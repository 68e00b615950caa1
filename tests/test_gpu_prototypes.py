"""P1-P3 (north_star extension, parity unpinned by the reference): CUDA prototype kernels
vs the float64 oracle built on PFGST.masked_feat_dist. 1e-5 relative; counts exact."""
import pytest
import torch

from oracle import prototypes as OP, ema as oema
from pfst_b200 import ops, prototypes as P
from pfst_b200.synthetic import blocky_labels

pytestmark = pytest.mark.gpu


def _case(B, D, h, w, C, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    feats = torch.relu(torch.randn((B, D, h, w), generator=g))
    labels = blocky_labels(B, H, W, C, g, min_rect=2, max_rect=max(4, H // 2))
    return feats, labels


@pytest.mark.parametrize("B,D,h,w,C,H,W", [(2, 32, 16, 16, 6, 128, 128), (8, 512, 64, 64, 6, 512, 512),
                                            (3, 16, 15, 15, 33, 120, 120), (2, 64, 128, 128, 2, 1024, 1024),
                                            (1, 8, 5, 7, 4, 40, 56)])
def test_accumulate_and_finalize(cuda, B, D, h, w, C, H, W):
    feats, labels = _case(B, D, h, w, C, H, W)
    sums, counts = OP.proto_accumulate(feats, labels[:, 0], C)
    bank = P.PrototypeBank(C, D, cuda)
    bank.accumulate(feats.to(cuda), labels.to(cuda))
    packed = bank.packed.cpu()
    got_sums, got_cnt = packed[:C * D].view(C, D), packed[C * D:]
    assert torch.equal(got_cnt.long(), counts)
    scale = sums.abs().max()
    assert (got_sums.double() - sums).abs().max() <= 1e-5 * scale
    mu = bank.finalize().cpu()
    mu_o, seen_o = OP.proto_finalize(sums, counts)
    assert torch.equal(bank.seen.cpu().bool(), seen_o)
    assert torch.equal(bank.counts.cpu(), counts)
    assert (mu.double() - mu_o.double()).abs().max() <= 1e-5 * mu_o.abs().max()
    assert float(bank.packed.abs().sum()) == 0.0


def test_confidence_mask_and_chained_batches(cuda):
    B, D, h, w, C, H, W = 2, 32, 16, 16, 6, 128, 128
    feats, labels = _case(B, D, h, w, C, H, W, seed=3)
    g = torch.Generator().manual_seed(9)
    conf = torch.rand((B, H, W), generator=g)
    s1, c1 = OP.proto_accumulate(feats, labels[:, 0], C, conf, 0.5)
    feats2, labels2 = _case(B, D, h, w, C, H, W, seed=4)
    s2, c2 = OP.proto_accumulate(feats2, labels2[:, 0], C)
    bank = P.PrototypeBank(C, D, cuda)
    bank.accumulate(feats.to(cuda), labels.to(cuda), conf.to(cuda), 0.5)
    bank.accumulate(feats2.to(cuda), labels2.to(cuda))
    packed = bank.packed.cpu()
    assert torch.equal(packed[C * D:].long(), c1 + c2)
    assert (packed[:C * D].view(C, D).double() - (s1 + s2)).abs().max() <= 1e-5 * (s1 + s2).abs().max()


@pytest.mark.parametrize("B,D,h,w,C,H,W,use_conf", [
    (2, 37, 36, 36, 6, 288, 288, True),      # 1296 px: one partial tile (no whole 128-float4 chunk), odd channel count
    (1, 24, 72, 72, 3, 144, 144, False),     # 5184 px: a full 4096-px tile + a 1088-px tail, labels 2x up-sampled
    (3, 9, 32, 32, 8, 256, 256, True),       # fewer channels than warps, the largest class count of the kernel
    (2, 16, 40, 28, 1, 320, 224, False),     # a single class, non-square
    (4, 130, 64, 64, 2, 512, 512, True)])
def test_masked_accumulation_path(cuda, B, D, h, w, C, H, W, use_conf):
    """C <= 8, h*w % 4 == 0, h*w > 576: `accumulate` is the single masked-accumulation launch
    (proto_accum_masked_kernel). Counts exact, sums within 1e-5 of the fp64 oracle, with and without
    the confidence mask, chained over two calls."""
    feats, labels = _case(B, D, h, w, C, H, W, seed=21)
    labels[:, :, : H // 5] = 255                       # an ignored band
    bank = P.PrototypeBank(C, D, cuda)
    assert bank.masked(h, w, feats.to(cuda))
    conf = torch.rand((B, H, W), generator=torch.Generator().manual_seed(5)) if use_conf else None
    s1, c1 = OP.proto_accumulate(feats, labels[:, 0], C, conf, 0.4 if use_conf else 0.0)
    feats2, labels2 = _case(B, D, h, w, C, H, W, seed=22)
    s2, c2 = OP.proto_accumulate(feats2, labels2[:, 0], C)
    bank.accumulate(feats.to(cuda), labels.to(cuda), None if conf is None else conf.to(cuda), 0.4 if use_conf else 0.0)
    bank.accumulate(feats2.to(cuda), labels2.to(cuda))
    packed = bank.packed.cpu()
    assert torch.equal(packed[C * D:].long(), c1 + c2)
    want = s1 + s2
    assert (packed[:C * D].view(C, D).double() - want).abs().max() <= 1e-5 * want.abs().max()


def test_prototype_ema_over_iterations(cuda):
    B, D, h, w, C, H, W = 2, 16, 8, 8, 5, 64, 64
    bank = P.PrototypeBank(C, D, cuda, alpha=0.999)
    mu_o, seen_o = None, None
    for it in range(4):
        feats, labels = _case(B, D, h, w, C, H, W, seed=10 + it)
        if it == 1:
            labels[labels == 2] = 0          # class 2 absent in this iteration: keeps its value
        sums, counts = OP.proto_accumulate(feats, labels[:, 0], C)
        a = oema.alpha_teacher(max(it, 1), 0.999)
        mu_o, seen_o = OP.proto_finalize(sums, counts, mu_o, seen_o, float(torch.tensor(a, dtype=torch.float32)),
                                         float(torch.tensor(1 - a, dtype=torch.float32)))
        mu = bank.update(feats.to(cuda), labels.to(cuda)).cpu()
        assert (mu.double() - mu_o.double()).abs().max() <= 1e-5 * mu_o.abs().max(), it
        assert torch.equal(bank.seen.cpu().bool(), seen_o)


@pytest.mark.parametrize("B,D,h,w,C,H,W", [(2, 32, 16, 16, 6, 128, 128), (4, 512, 64, 64, 6, 512, 512),
                                            (3, 16, 15, 15, 33, 120, 120), (1, 8, 5, 7, 4, 40, 56)])
def test_proto_dist_loss_and_grad(cuda, B, D, h, w, C, H, W):
    feats, labels = _case(B, D, h, w, C, H, W, seed=5)
    g = torch.Generator().manual_seed(6)
    mu = torch.relu(torch.randn((C, D), generator=g))
    seen = torch.ones(C, dtype=torch.bool)
    seen[C - 1] = False
    f_o = feats.clone().double().requires_grad_(True)
    loss_o, valid = OP.proto_dist_loss(f_o, labels[:, 0], mu.double(), seen)
    (2.5 * loss_o).backward()
    f_c = feats.clone().to(cuda).requires_grad_(True)
    loss_c = P.proto_dist_loss(f_c, labels.to(cuda), mu.to(cuda), seen.to(cuda).to(torch.uint8))
    (2.5 * loss_c).backward()
    assert abs(float(loss_c) - float(loss_o)) <= 1e-5 * abs(float(loss_o))
    go, gc = f_o.grad, f_c.grad.cpu().double()
    assert (gc - go).abs().max() <= 1e-5 * go.abs().max()
    assert float(gc[:, :, ~valid[0]].abs().max() if (~valid[0]).any() else 0.0) >= 0.0


def test_proto_dist_all(cuda):
    for (B, D, h, w, C) in ((2, 32, 16, 16, 6), (2, 16, 15, 15, 33), (1, 512, 32, 32, 2)):
        g = torch.Generator().manual_seed(7)
        feats = torch.relu(torch.randn((B, D, h, w), generator=g))
        mu = torch.relu(torch.randn((C, D), generator=g))
        want = OP.proto_dist_all(feats.double(), mu.double())
        got = P.proto_dist_all(feats.to(cuda), mu.to(cuda)).cpu().double()
        assert (got - want).abs().max() <= 1e-5 * want.abs().max()


@pytest.mark.parametrize("B,D,h,w,C,H,W", [(8, 96, 64, 64, 6, 512, 512), (2, 40, 128, 128, 2, 1024, 1024),
                                            (5, 24, 15, 15, 33, 120, 120), (1, 8, 5, 7, 4, 40, 56)])
def test_single_launch_accumulate_equals_the_split_form(cuda, B, D, h, w, C, H, W):
    """pfst_proto_accum (self-contained) vs pfst_proto_order + pfst_proto_accum_ordered: identical
    counts, sums equal up to the summation order."""
    feats, labels = _case(B, D, h, w, C, H, W, seed=5)
    g = torch.Generator().manual_seed(2)
    conf = torch.rand((B, H, W), generator=g).to(cuda)
    a, b = P.PrototypeBank(C, D, cuda), P.PrototypeBank(C, D, cuda)
    a.accumulate(feats.to(cuda), labels.to(cuda), conf, 0.3)
    b.accumulate_single_launch(feats.to(cuda), labels.to(cuda), conf, 0.3)
    pa, pb = a.packed.cpu(), b.packed.cpu()
    assert torch.equal(pa[C * D:], pb[C * D:])
    assert (pa[:C * D] - pb[:C * D]).abs().max() <= 1e-5 * pb[:C * D].abs().max()


def test_peer_board_single_rank_equals_local_finalize(cuda):
    """csrc/peer.cu with one rank (push into the own board, token wait, sum of one chunk) is
    bit-identical to proto_finalize_kernel over several iterations (EMA of the prototypes, device-
    resident iteration, packed re-zeroed). The N > 1 exchange is tests/test_gpu_multi_rank.py."""
    import socket
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
        created = True
    try:
        B, D, h, w, C, H, W = 2, 48, 16, 16, 7, 128, 128
        local, peer = P.PrototypeBank(C, D, cuda), P.PrototypeBank(C, D, cuda)
        board = peer.attach_peer_board(timeout_s=2.0)
        assert board.world == 1
        for it in range(4):
            feats, labels = _case(B, D, h, w, C, H, W, seed=20 + it)
            if it == 2:
                labels[labels == 3] = 255            # a class without pixels keeps its prototype
            for bank in (local, peer):
                bank.accumulate(feats.to(cuda), labels.to(cuda))
                bank.finalize()
            assert torch.equal(local.mu, peer.mu), it
            assert torch.equal(local.seen, peer.seen) and torch.equal(local.counts, peer.counts)
            assert torch.equal(local.iter_state, peer.iter_state)
            assert float(peer.packed.abs().sum()) == 0.0
        board.check()
        assert int(board.status[1]) == 4
        board.close()
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("B,D,h,w", [(2, 512, 64, 64), (1, 19, 5, 7), (3, 64, 15, 15)])
def test_masked_feat_dist_matches_the_reference_method(cuda, B, D, h, w):
    """G1: PFGST.masked_feat_dist (pfgst.py:168-177; oracle pinned to the reference method in
    tests/test_oracle_pins.py) — value and both gradients, with / without a mask, 1e-5."""
    g = torch.Generator().manual_seed(B * 31 + D)
    f1 = torch.randn((B, D, h, w), generator=g)
    f2 = torch.randn((B, D, h, w), generator=g)
    f2[0, :, 0, 0] = f1[0, :, 0, 0]                          # a zero distance: gradient 0 there, not NaN
    mask = torch.rand((B, 1, h, w), generator=g) > 0.3
    mask[0, 0, 0, 0] = True
    for m in (None, mask):
        a1, a2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
        ref = OP.masked_feat_dist(a1, a2, m)
        (ref * 0.37).backward()
        b1, b2 = f1.to(cuda).requires_grad_(True), f2.to(cuda).requires_grad_(True)
        got = P.masked_feat_dist(b1, b2, None if m is None else m.to(cuda))
        (got * 0.37).backward()
        assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref))
        for x, y in ((b1.grad, a1.grad), (b2.grad, a2.grad)):
            assert (x.cpu() - y).abs().max() <= 1e-5 * y.abs().max()
        assert float(b1.grad[0, :, 0, 0].abs().max()) == 0.0
    empty = torch.zeros((B, 1, h, w), dtype=torch.bool)
    assert torch.isnan(P.masked_feat_dist(f1.to(cuda), f2.to(cuda), empty.to(cuda)))
    assert torch.isnan(OP.masked_feat_dist(f1, f2, empty))


def test_drop_in_accepts_the_imnet_feature_distance_keys(cuda):
    """pfgst.py:84-87: imnet_feature_dist_lambda > 0 builds a third copy of the segmentor and nothing
    else (the reference's forward_train has no feature-distance term); masked_feat_dist is callable."""
    from pfst_b200.uda import PFGST
    from tests.fake_segmentor import TinySegmentor
    from tests.golden.make_golden import STEP_CFG
    cfg = dict(STEP_CFG)
    cfg.update(imnet_feature_dist_lambda=0.005, imnet_feature_dist_classes=[1, 2], imnet_feature_dist_scale_min_ratio=0.75)
    cfg['model'] = lambda: TinySegmentor(6, 16, seed=0)
    m = PFGST(**cfg).to(cuda)
    assert m.enable_fdist and m.get_imnet_model() is not None and m.get_imnet_model() is not m.get_model()
    f = torch.randn((2, 16, 8, 8), device=cuda)
    assert float(m.masked_feat_dist(f, f + 1.0)) == pytest.approx(4.0, rel=1e-6)
