"""Small-shape invocations of the kernels added last (fused arg-max + confusion, Gaussian blur)
for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_new.py` (development tool)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (N, C, H, W, ldt) in [(2, 6, 40, 36, torch.uint8), (1, 6, 15, 15, torch.uint8), (2, 33, 16, 16, torch.int64),
                          (1, 2, 64, 64, torch.int32), (1, 255, 8, 8, torch.uint8)]:
    x = torch.randn((N, C, H, W), generator=g).to(dev)
    lab = torch.randint(0, min(C + 1, 255), (N, H, W), generator=g).to(ldt).to(dev)
    conf, pred = ops.argmax_confusion(x, lab, C, per_image=True, return_pred=torch.uint8)
    conf, pred = ops.argmax_confusion(x, lab, C, return_pred=torch.int64)
    assert int(conf.sum()) == int((lab != 255).sum())
for (N, C, H, W, ks, sig) in [(2, 3, 70, 50, None, [0.3, 1.1]), (1, 3, 128, 128, None, [1.15]),
                              (1, 1, 33, 65, (9, 21), [4.0])]:
    x = torch.randn((N, C, H, W), generator=g).to(dev)
    y = ops.gaussian_blur(x, sig, ks)
    assert bool(torch.isfinite(y).all())
torch.cuda.synchronize()
print("sanitize_new ok")
