"""Host-side cost of SelfTrainingStep.run (development tool): wall time per call without a
device sync vs the device time per step, plus a cProfile of the call. Works under torchrun."""
import cProfile
import os
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200.step import SelfTrainingStep  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, model_params, step_inputs  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
g = torch.Generator().manual_seed(1234 + rank)
np.random.seed(1234 + rank)
inp = {k: v.to(dev) for k, v in step_inputs(wl, 1234 + rank).items()}
student = [p.to(dev) for p in model_params(wl.C, g)]
teacher = [p.to(dev) for p in model_params(wl.C, g)]
step = SelfTrainingStep(teacher, student, wl.C, wl.D, dev, graphs=True)
run = lambda it: step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                          inp["logits_trg"], inp["x_src"], inp["x_ema"]) and None or step.prefetch(inp["gt"])
for i in range(10):
    run(i)
torch.cuda.synchronize()
N = 200
host, blocked = [], [0.0]
_choose = step.plan.choose


def timed_choose(*a, **k):                    # draw + (possible) wait for the presence bits
    t = time.perf_counter()
    r = _choose(*a, **k)
    blocked[0] += time.perf_counter() - t
    return r


step.plan.choose = timed_choose
t0 = time.perf_counter()
for i in range(N):
    a = time.perf_counter()
    run(20 + i)
    host.append(time.perf_counter() - a)
step.plan.choose = _choose
torch.cuda.synchronize()
t1 = time.perf_counter()
if rank == 0:
    host.sort()
    print(f"world {world}: wall per step {1e6 * (t1 - t0) / N:.1f} us; host time in run(): median "
          f"{1e6 * host[N // 2]:.1f} us, p10 {1e6 * host[N // 10]:.1f}, p90 {1e6 * host[9 * N // 10]:.1f}; "
          f"of which in ClassMixPlan.choose (draw + wait for the presence bits) {1e6 * blocked[0] / N:.1f} us per step")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(100):
        run(300 + i)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
else:
    for i in range(100):
        run(300 + i)
    torch.cuda.synchronize()
if world > 1:
    dist.destroy_process_group()
