"""N ranks, N GPUs: the multi-rank schedule of SelfTrainingStep against the CPU oracle fed with
every rank's inputs, in its three forms — "peer": the one-shot all-reduce over NVLink peer
memory fused into the finalise kernel (csrc/peer.cu; the default), "nccl_graph": ncclAllReduce
captured inside the step's CUDA graph, "nccl_eager": three graphs around an eager all-reduce.
Needs >= 2 GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2` / `--gpus 8`)."""
import os
import socket
import sys
import time
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu
ITERS = 3
MODES = {"peer": ("1", "1"), "nccl_graph": ("0", "1"), "nccl_eager": ("0", "0")}   # PEER_REDUCE, NCCL_IN_GRAPH


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _params(rank):
    g = torch.Generator().manual_seed(50)            # identical student / teacher on every rank (DDP)
    shapes = [(5,), (4097,), (16, 3, 3, 3)]
    return [0.02 * torch.randn(s, generator=g) for s in shapes], [0.02 * torch.randn(s, generator=g) for s in shapes]


def _worker(rank, port, out_dir, mode, WORLD, graphs):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PFST_PEER_REDUCE=MODES[mode][0],
                      PFST_NCCL_IN_GRAPH=MODES[mode][1])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    try:
        from pfst_b200.step import SelfTrainingStep
        from pfst_b200.synthetic import WORKLOADS, step_inputs
        wl = WORKLOADS["tiny"]
        inp = {k: v.to(dev) for k, v in step_inputs(wl, 1234 + rank).items()}
        student, teacher = _params(rank)
        step = SelfTrainingStep([p.to(dev) for p in teacher], [p.to(dev) for p in student], wl.C, wl.D, dev, graphs=graphs)
        assert step.peer_reduce == (mode == "peer")
        res = []
        for it in range(ITERS):
            np.random.seed(100 * rank + it)
            out = step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                           inp["logits_trg"], inp["x_src"], inp["x_ema"])
            torch.cuda.synchronize()
            res.append({k: out[k].cpu().clone() for k in ("losses", "proto_loss", "mu", "grad_x_src", "mix_masks")})
        assert (step.bank.peer is not None) == (mode == "peer")
        if step.bank.peer is not None:
            step.bank.peer.check()
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
        step.close()                       # collective: graphs (NCCL work) and peer boards go before the communicator
        del step, out
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("WORLD,mode,graphs", [(2, "peer", True), (2, "peer", False), (2, "nccl_graph", True),
                                               (2, "nccl_eager", True), (8, "peer", True)])
def test_multi_rank_step_matches_oracle(tmp_path, WORLD, mode, graphs):
    if torch.cuda.device_count() < WORLD:
        pytest.skip(f"needs {WORLD} GPUs")
    ctx = mp.spawn(_worker, args=(_free_port(), str(tmp_path), mode, WORLD, graphs), nprocs=WORLD, join=False)
    deadline = time.time() + 150                    # a hung collective must not hold the GPUs
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for proc in ctx.processes:
                proc.kill()
            pytest.fail("multi-rank step did not finish within 150 s")
    from oracle import prototypes as OP, pseudo as opl, step as ostep
    from pfst_b200.synthetic import WORKLOADS, step_inputs
    wl = WORKLOADS["tiny"]
    hosts = [step_inputs(wl, 1234 + r) for r in range(WORLD)]
    outs = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(WORLD)]
    # what each rank contributes to the all-reduce (inputs are the same every iteration)
    contrib = []
    for h in hosts:
        label, _, _ = opl.pseudo_label(h["ema_logits"], 0.98)
        contrib.append(OP.proto_accumulate(h["x_ema"], label, wl.C))
    for r in range(WORLD):
        student, teacher = _params(r)
        state = None
        for it in range(ITERS):
            ref = ostep.hot_path_step(it, teacher, student, hosts[r], wl.C, proto_state=state,
                                      rng=np.random.RandomState(100 * r + it),
                                      peer_protos=[c for q, c in enumerate(contrib) if q != r])
            state = ref["proto_state"]
            got = outs[r][it]
            assert torch.equal(got["mix_masks"], ref["mix_masks"])
            assert (got["mu"] - ref["mu"]).abs().max() <= 1e-5 * ref["mu"].abs().max(), (r, it)
            assert torch.all((got["losses"] - ref["losses"]).abs() <= 1e-5 * ref["losses"].abs() + 1e-9)
            assert abs(float(got["proto_loss"]) - float(ref["proto_loss"])) <= 1e-5 * abs(float(ref["proto_loss"]))
            assert (got["grad_x_src"] - ref["grad_x_src"]).abs().max() <= 1e-5 * ref["grad_x_src"].abs().max()
    # the prototypes are global: identical on every rank (bit-identical by construction on the
    # peer board — fixed rank order; NCCL gives every rank the same reduced buffer)
    for it in range(ITERS):
        for r in range(1, WORLD):
            assert torch.equal(outs[0][it]["mu"], outs[r][it]["mu"])
