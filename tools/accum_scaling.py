"""proto_accum time vs channel count (development tool): T(D) = prologue + D * per-plane cost."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200.prototypes import PrototypeBank  # noqa: E402
from pfst_b200.synthetic import blocky_labels  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, C, H, W, h, w = 8, 6, 512, 512, 64, 64
lab = blocky_labels(B, H, W, C, g)[:, 0].contiguous().to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for D in (64, 128, 256, 512, 1024, 2048):
    xs = [torch.relu(torch.randn((B, D, h, w), generator=g)).to(dev) for _ in range(4)]
    bank = PrototypeBank(C, D, dev)
    for x in xs:
        bank.accumulate(x, lab)
    torch.cuda.synchronize()
    ts = []
    for r in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        bank.accumulate(xs[r % 4], lab)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    mb = B * D * h * w * 4 / 1e6
    print(f"D={D:5d}  {mb:7.1f} MB  median {ts[len(ts) // 2]:7.2f} us  best {ts[0]:7.2f} us  -> {mb / ts[len(ts) // 2] / 1e3:.2f} TB/s")
