"""The rest of the offline class-wise pseudo-labelling path on the GPU:
_cal_loc_dis, _cal_sigmas (PseudoLabelingHookV4) and the loader's label rule against the CPU
restatement (oracle/offline_labels.py, pinned to the reference) and the reference-written fixture
tests/golden/offline_labels.npz. Distances / sigmas 1e-5 relative; labels bit-exact except pixels
whose entropy lies within 1e-6 of the threshold."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import offline_labels as OL
from pfst_b200 import pseudo_labeling as PL
from tests.golden.make_golden import (loader_rule_cases, loader_rule_inputs, offline_label_cases,
                                      offline_label_feats)

pytestmark = pytest.mark.gpu
Z = np.load(Path(__file__).resolve().parent / "golden" / "offline_labels.npz")


def _close(a, b, tol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.abs(b) + tol * np.abs(b).max() * 1e-2 + 1e-12)


def test_loc_dis_and_sigmas_match_the_reference_fixture(cuda):
    for name, n, C, H, W, dils, means, ratio, seed in offline_label_cases():
        feats = offline_label_feats(n, C, H, W, seed)
        lds = [PL.cal_loc_dis([f.to(cuda) for f in fl], 3, dils) for fl in feats]
        for i, ld in enumerate(lds):
            assert list(ld) == [f'level{l}_dila@{d}' for l in (0, 1) for d in dils]
            for k, v in ld.items():
                want = Z[f"{name}_locdis_{i}_{k}"]
                assert tuple(v.shape) == want.shape and _close(v.cpu().numpy(), want), (name, i, k)
        sig = PL.cal_sigmas(lds, [0, 1], dils, means, ratio, np.random.RandomState(40 + seed))
        for k, v in sig.items():
            want = float(Z[f"{name}_sigma_{k}"])
            assert abs(v - want) <= 1e-5 * want + 2e-6, (name, k, v, want)


@pytest.mark.parametrize("B,C,H,W,dil", [(2, 512, 64, 64, 2), (3, 40, 15, 15, 1), (1, 7, 5, 9, 4), (2, 33, 30, 30, 3)])
def test_loc_dis_batch_vs_oracle(cuda, B, C, H, W, dil):
    g = torch.Generator().manual_seed(C + dil)
    feats = torch.relu(torch.randn((B, C, H, W), generator=g))
    got = PL.loc_dis_batch(feats.to(cuda), dil).cpu()
    for b in range(B):
        want = OL.cal_loc_dis([feats[b]], 3, [dil])[f'level0_dila@{dil}'][0]
        assert _close(got[b].numpy(), want.numpy()), b
    assert float(got[..., 4].abs().max()) == 0.0                       # centre tap: distance to itself
    x0 = feats[:, :, 0, 0].double().pow(2).sum(1)                       # corner: out-of-image taps read zeros
    assert _close(got[:, 0, 0, 0].numpy(), x0.numpy())


def test_sigma_search_large_sample_equals_oracle(cuda):
    g = torch.Generator().manual_seed(3)
    feats = [[torch.relu(torch.randn((64, 48, 48), generator=g))] for _ in range(4)]
    lds_o = [OL.cal_loc_dis(f, 3, [2]) for f in feats]
    lds_g = [PL.cal_loc_dis([f[0].to(cuda)], 3, [2]) for f in feats]
    want = OL.cal_sigmas(lds_o, [0], [2], [0.3, 0.6, 0.9], 0.5, np.random.RandomState(9))
    got = PL.cal_sigmas(lds_g, [0], [2], [0.3, 0.6, 0.9], 0.5, np.random.RandomState(9))
    assert list(got) == list(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-5 * want[k] + 2e-6, (k, got[k], want[k])


def test_loader_rule_matches_the_reference_fixture(cuda):
    for name, C, H, W, rz, seed in loader_rule_cases():
        logits, thres = loader_rule_inputs(C, H, W, seed)
        got = PL.loader_pseudo_labels(torch.from_numpy(logits).to(cuda), thres, rz).cpu().numpy()
        want = Z[f"loader_{name}"]
        assert got.dtype == np.uint8 and got.shape == want.shape
        # pixels whose entropy is within 1e-6 of their threshold may flip (expf vs numpy's exp)
        p = np.exp(logits) / np.exp(logits).sum(axis=0)
        ent = -(p * np.log(p + 1e-8)).sum(axis=0)
        safe = np.abs(ent - thres[logits.argmax(axis=0)]) > 1e-6
        assert np.array_equal(got[safe], want[safe]), name
        assert (~safe).mean() < 0.01


def test_loader_rule_batch_and_nan(cuda):
    g = torch.Generator().manual_seed(8)
    logits = 2 * torch.randn((5, 6, 33, 17), generator=g)
    logits[0, 2, 3, 4] = float("nan")
    logits[1, :, 0, 0] = 0.75
    thres = np.array([0.9, 0.4, 1.3, 0.0, 0.7, 1.0], dtype=np.float32)
    for rz in (False, True):
        got = PL.loader_pseudo_labels(logits.to(cuda), thres, rz).cpu().numpy()
        for b in range(5):
            want = OL.loader_pseudo_labels(logits[b].numpy(), thres, rz)
            z = logits[b].numpy()
            p = np.exp(z) / np.exp(z).sum(axis=0)
            ent = -(p * np.log(p + 1e-8)).sum(axis=0)
            safe = ~(np.abs(ent - thres[z.argmax(axis=0)]) <= 1e-6)
            assert np.array_equal(got[b][safe], want[safe]), (rz, b)
