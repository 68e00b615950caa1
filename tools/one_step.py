"""Runs W warm-up steps and then N hot-path steps inside a cudaProfilerStart/Stop range
(development tool for `ncu --profile-from-start off`).

    python tools/one_step.py [workload] [steps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200.step import SelfTrainingStep  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, model_params, step_inputs  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
inp = {k: v.to(dev) for k, v in step_inputs(wl, 1234).items()}
student = [p.to(dev) for p in model_params(wl.C, g)]
teacher = [p.to(dev) for p in model_params(wl.C, g)]
step = SelfTrainingStep(teacher, student, wl.C, wl.D, dev, dilation=wl.dilation,
                        downscale=wl.downscale if wl.downscale != 1.0 else None, max_batch=max(wl.B, 64), graphs="--graphs" in sys.argv)


def run(it):
    return step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                    inp["logits_trg"], inp["x_src"], inp["x_ema"])


for i in range(3):
    run(i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(steps):
    run(10 + i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("one_step ok")
