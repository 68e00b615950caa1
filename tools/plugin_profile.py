"""Host-side cost of PFGST.forward_train on the bench's replay segmentor (development tool):
wall time per call vs device time, and a cProfile of the call."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, step_inputs  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
g = torch.Generator().manual_seed(1234)
np.random.seed(1234)
host = step_inputs(wl, 1234)
pinned = {k: v.pin_memory() for k, v in host.items()}
pinned["logits_src"] = (2.0 * torch.randn(host["logits_trg"].shape, generator=g)).pin_memory()
pinned["target_img"] = torch.randn(host["img"].shape, generator=g).pin_memory()
pinned["loss_ce"] = torch.rand((), generator=g).pin_memory()
feed = bench.HostFeed(pinned, dev)
bench.ReplaySegmentor.FEED = feed
model = bench.build_plugin(wl, 1234, dev)
metas = [{'img_norm_cfg': {'mean': [0., 0., 0.], 'std': [1., 1., 1.]}}] * wl.B
feed.issue(0)
d = feed.acquire(0)
torch.cuda.synchronize()
step = lambda: model.forward_train(d["img"], metas, d["gt"], d["target_img"], metas, d["target_img_strong_aug"])[0]
for _ in range(10):
    step()
torch.cuda.synchronize()
if "--timeline" in sys.argv:
    # device + runtime-call timeline of two steady-state iterations (gaps, overlap, blocking calls)
    import json
    import tempfile
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(6):
            step()
        torch.cuda.synchronize()
    path = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(path)
    allev = json.load(open(path))["traceEvents"]
    keep = ("cudaGraphLaunch", "cudaLaunchKernel", "cudaEventSynchronize", "cudaMemcpyAsync", "cudaStreamWaitEvent",
            "cudaStreamSynchronize")
    ev = [e for e in allev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") or
          (e.get("cat") == "cuda_runtime" and e.get("name") in keep)]
    ev.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "ema_multi" in e["name"] and e.get("cat") == "kernel"]
    lo, hi = starts[3], starts[5]
    t0 = ev[lo]["ts"]
    for e in ev[max(lo - 12, 0):hi]:
        name = e["name"].replace("pfst::", "").split("(")[0][:44]
        where = "  host" if e.get("cat") == "cuda_runtime" else f"stream {e['args'].get('stream', '?'):>3}"
        print(f"{e['ts'] - t0:8.1f} -> {e['ts'] - t0 + e['dur']:8.1f}  ({e['dur']:6.1f} us)  {where}  {name}")
    print(f"two steps: {ev[hi]['ts'] - t0:.1f} us")
    sys.exit(0)
N = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(N):
    lv = step()
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"{wl.name}: host {1e6 * t_host / N:.1f} us per forward_train (no sync), device {1e3 * e0.elapsed_time(e1) / N:.1f} us per step")
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(35)
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
