"""Stress check of DESIGN.md 3.2 for the current schedules: N graph replays of the one-graph step per workload,
dot maps compared bit for bit with the TMA dots kernel launched alone; also the prototype sums (the other
x_ema reader) against the first replay's. Prints one line per workload; exits 1 on any mismatch."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200 import ops  # noqa: E402
from pfst_b200.step import SelfTrainingStep  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, step_inputs  # noqa: E402

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 60
names = sys.argv[2:] or ["cfg1", "cfg2", "cfg3", "cfg4"]
cuda = torch.device("cuda:0")
bad_total = 0
for name in names:
    wl = WORKLOADS[name]
    inp = {k: v.to(cuda) for k, v in step_inputs(wl, 1234).items()}
    down = wl.downscale if wl.downscale != 1.0 else None
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 3, 3, 3), (64,), (wl.C, 512, 1, 1), (wl.C,), (100003,), (4_000_000,)]
    student = [(0.02 * torch.randn(s, generator=g)).to(cuda) for s in shapes]
    teacher = [(0.02 * torch.randn(s, generator=g)).to(cuda) for s in shapes]
    step = SelfTrainingStep(teacher, student, wl.C, wl.D, cuda, dilation=wl.dilation, downscale=down,
                            max_batch=max(64, wl.B), graphs=True)
    bad = 0
    ref = None
    for it in range(n_rep):
        np.random.seed(7)                      # same class draw every replay: outputs must repeat exactly
        step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"], inp["logits_trg"],
                 inp["x_src"], inp["x_ema"])
        torch.cuda.synchronize()
        b, geo = next(iter(step._bufs.values()))
        if ref is None:
            d = geo.dilation // geo.up
            ref = torch.empty_like(b.dots)
            ops.neigh_dots_slot(inp["x_ema"], d, 0, ref)
            ops.neigh_dots_slot(inp["x_src"], d, 1, ref)
            torch.cuda.synchronize()
        bad += int((b.dots != ref).any())
    masked = step.bank.masked(inp["x_ema"].shape[2], inp["x_ema"].shape[3], inp["x_ema"])
    print(f"{name}: {n_rep} replays, masked accumulation={masked}, replays with a wrong dot map: {bad}", flush=True)
    bad_total += bad
    step.close() if hasattr(step, "close") else None
    del step, inp
    torch.cuda.empty_cache()
sys.exit(1 if bad_total else 0)
