"""Development aid: replays the one-graph step and compares its dot maps with the TMA dots kernel launched alone."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200 import ops
from pfst_b200.step import SelfTrainingStep
from pfst_b200.synthetic import WORKLOADS, step_inputs
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"]
cuda = torch.device("cuda:0")
inp = {k: v.to(cuda) for k, v in step_inputs(wl, 1234).items()}
down = wl.downscale if wl.downscale != 1.0 else None
g = torch.Generator().manual_seed(3)
shapes = [(64, 3, 3, 3), (64,), (wl.C, 512, 1, 1), (wl.C,), (100003,)]
student = [(0.02 * torch.randn(s, generator=g)).to(cuda) for s in shapes]
teacher = [(0.02 * torch.randn(s, generator=g)).to(cuda) for s in shapes]
step = SelfTrainingStep(teacher, student, wl.C, wl.D, cuda, dilation=wl.dilation, downscale=down, max_batch=64, graphs=True)
ref0, ks = ops.neigh_dots_slot(inp["x_ema"], wl.dilation // 1 if down is None else wl.dilation // 1, 0)
for it in range(6):
    out = step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"], inp["logits_trg"], inp["x_src"], inp["x_ema"])
    torch.cuda.synchronize()
    b, geo = next(iter(step._bufs.values()))
    d = geo.dilation // geo.up
    ref = torch.empty_like(b.dots)
    ops.neigh_dots_slot(inp["x_ema"], d, 0, ref); ops.neigh_dots_slot(inp["x_src"], d, 1, ref)
    torch.cuda.synchronize()
    diff = (b.dots - ref).abs()
    per = diff.amax(dim=(2, 4, 5))    # (ks, slot, map)
    bad = (diff > 0).nonzero()
    if hasattr(step, "_scratch_dots"):
        sd = (step._scratch_dots[:, 1] != ref[:, 1]).sum().item()
        print(f"   scratch dots (next to accum/finalize/dist): n_bad={sd}", "SCRATCH_CLEAN" if sd == 0 else "SCRATCH_BAD")
    print(f"it {it}: ks={b.ks} max diff per (split,slot,map):\n{per.cpu().numpy()}\n   n_bad={bad.shape[0]}", bad[:5].tolist() if bad.numel() else "")
b, geo = next(iter(step._bufs.values()))
ws = step.bank._order_ws
print("dots ptr", hex(b.dots.data_ptr()), "bytes", b.dots.numel() * 4, "end", hex(b.dots.data_ptr() + b.dots.numel() * 4))
print("order ws ptr", hex(ws.data_ptr()), "bytes", ws.numel(), "end", hex(ws.data_ptr() + ws.numel()))
print("packed", hex(step.bank.packed.data_ptr()), "label", hex(b.label.data_ptr()), b.label.numel() * 8)
flat_bad = (b.dots.flatten() != ref.flatten()).nonzero().flatten()
print("first bad flat idx", flat_bad[:10].tolist(), "last", flat_bad[-5:].tolist())
print("got", b.dots.flatten()[flat_bad[:12]].tolist())
print("ref", ref.flatten()[flat_bad[:12]].tolist())
# runs of consecutive bad indices
d = flat_bad[1:] - flat_bad[:-1]
starts = torch.cat([flat_bad[:1], flat_bad[1:][d > 1]])
print("n runs", starts.numel(), "run starts (first 10)", starts[:10].tolist())
