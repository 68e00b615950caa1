"""Per-kernel time breakdown of one hot-path step (development tool): runs the step under
torch.profiler and prints the CUDA kernel table."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200.step import SelfTrainingStep  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, model_params, step_inputs  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
inp = {k: v.to(dev) for k, v in step_inputs(wl, 1234).items()}
student = [p.to(dev) for p in model_params(wl.C, g)]
teacher = [p.to(dev) for p in model_params(wl.C, g)]
step = SelfTrainingStep(teacher, student, wl.C, wl.D, dev, dilation=wl.dilation,
                        downscale=wl.downscale if wl.downscale != 1.0 else None, max_batch=max(wl.B, 64), graphs="--graphs" in sys.argv)
run = lambda it: step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                          inp["logits_trg"], inp["x_src"], inp["x_ema"]) and None or step.prefetch(inp["gt"])
for i in range(5):
    run(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(10):
        run(10 + i)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / max(e.count, 1), e.count) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(t * c for _, t, c in rows) / 10
for name, t, c in sorted(rows, key=lambda r: -r[1] * r[2]):
    print(f"{t:9.2f} us x{c // 10 if c >= 10 else c:<3d} {name[:90]}")
print(f"{tot:9.2f} us  device time per step (sum over kernels, memsets and copies)")
