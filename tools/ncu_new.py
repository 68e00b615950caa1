"""One launch each (after a warm-up launch) of the kernels added last, at their benchmark sizes, for
`ncu --set full -k regex:'gaussian_blur|argmax_confusion' --launch-skip 2 -c 2` (development tool)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200 import ops  # noqa: E402
from pfst_b200.synthetic import blocky_labels, teacher_logits  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
logits = teacher_logits(8, 6, 1024, 1024, g).to(dev)
gt = blocky_labels(8, 1024, 1024, 6, g)[:, 0].to(torch.uint8).to(dev)
img = torch.randn((8, 3, 512, 512), generator=g).to(dev)
out = torch.zeros((8, 7, 7), dtype=torch.int64, device=dev)
for _ in range(2):
    ops.argmax_confusion(logits, gt, 6, per_image=True, out=out)
    ops.gaussian_blur(img, [1.15] * 8)
torch.cuda.synchronize()
print("ncu_new ok")
