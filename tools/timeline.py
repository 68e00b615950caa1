"""Prints the device timeline (start/end per kernel, per stream) of two consecutive hot-path
steps with CUDA graphs on — shows gaps and overlap (development tool)."""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200.step import SelfTrainingStep  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, model_params, step_inputs  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
np.random.seed(1234)
inp = {k: v.to(dev) for k, v in step_inputs(wl, 1234).items()}
student = [p.to(dev) for p in model_params(wl.C, g)]
teacher = [p.to(dev) for p in model_params(wl.C, g)]
step = SelfTrainingStep(teacher, student, wl.C, wl.D, dev, dilation=wl.dilation,
                        downscale=wl.downscale if wl.downscale != 1.0 else None, max_batch=max(wl.B, 64),
                        graphs="--eager" not in sys.argv)
run = lambda it: step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                          inp["logits_trg"], inp["x_src"], inp["x_ema"]) and None or step.prefetch(inp["gt"])
for i in range(20):
    run(i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
        run(30 + i)
    torch.cuda.synchronize()
path = tempfile.mktemp(suffix=".json")
prof.export_chrome_trace(path)
allev = json.load(open(path))["traceEvents"]
ev = [e for e in allev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") or
      (e.get("cat") == "cuda_runtime" and e.get("name") in ("cudaGraphLaunch", "cudaLaunchKernel", "cudaEventSynchronize",
                                                            "cudaMemcpyAsync", "cudaStreamWaitEvent"))]
ev.sort(key=lambda e: e["ts"])
# take steps 3 and 4 (steady state): split at class_presence kernels
starts = [i for i, e in enumerate(ev) if "class_presence" in e["name"] and e.get("cat") == "kernel"]
lo, hi = starts[3], starts[5]
t0 = ev[lo]["ts"]
for e in ev[lo:hi]:
    name = e["name"].replace("pfst::", "").split("(")[0][:44]
    where = "  host" if e.get("cat") == "cuda_runtime" else f"stream {e['args'].get('stream', '?'):>3}"
    print(f"{e['ts'] - t0:8.1f} -> {e['ts'] - t0 + e['dur']:8.1f}  ({e['dur']:6.1f} us)  {where}  {name}")
print(f"two steps: {ev[hi]['ts'] - t0:.1f} us")
