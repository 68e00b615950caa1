"""BASELINE configs[4]: mIoU evaluation sweep — 6-class confusion matrix over 10 000 synthetic
1024x1024 label maps sharded across the ranks (contiguous shards, `evaluation.shard_range`), the
int64 confusion matrix all-reduced once at the end (the only collective of evaluation).

    python tools/eval_sweep.py [--maps 10000] [--resident 64] [--mode labels|logits]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/eval_sweep.py ...

A shard of 1250 maps is 11.8 GB as (int64 pred, uint8 gt); `--resident` distinct maps are kept in
HBM (larger than L2) and cycled. `--mode logits` times the fused arg-max + confusion kernel on
(N,6,1024,1024) fp32 logits instead (25 B/px). Device time by CUDA events, max over ranks.
Prints one JSON line on rank 0. (No CPU leg here: only tests/, smoke() and bench.py may run the
oracle; the reference's intersect_and_union takes ~26 ms per 1024^2 map on the host)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pfst_b200 import ops  # noqa: E402
from pfst_b200.evaluation import metrics as M  # noqa: E402
from pfst_b200.synthetic import eval_maps  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--maps", type=int, default=10000)
    ap.add_argument("--resident", type=int, default=64)
    ap.add_argument("--batch", type=int, default=16, help="maps per launch")
    ap.add_argument("--mode", default="labels", choices=["labels", "logits"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ops.device_check()
    C, H, W = 6, 1024, 1024
    lo, hi = M.shard_range(args.maps, rank, world)
    n_local = hi - lo
    res = max(args.batch, min(args.resident, n_local) // args.batch * args.batch)
    pred, gt = eval_maps(res, H, W, C, seed=1234 + rank)
    d_gt = torch.from_numpy(gt).to(dev)
    if args.mode == "labels":
        d_in = torch.from_numpy(pred).to(dev)
        bytes_per_map = 9 * H * W
    else:
        g = torch.Generator().manual_seed(1234 + rank)
        d_in = torch.empty((res, C, H, W), dtype=torch.float32, device=dev)
        for i in range(0, res, 8):
            d_in[i:i + 8] = (4 * torch.randn((min(8, res - i), C, H, W), generator=g)).to(dev)
        bytes_per_map = (4 * C + 1) * H * W
    meter = M.ConfusionMeter(C, device=dev)
    update = meter.update if args.mode == "labels" else meter.update_logits

    def sweep():
        done = 0
        while done < n_local:
            i = (done % res)
            n = min(args.batch, n_local - done, res - i)
            update(d_in[i:i + n], d_gt[i:i + n])
            done += n

    for _ in range(2):      # warm-up over a short prefix
        update(d_in[:args.batch], d_gt[:args.batch])
    meter.conf.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    sweep()
    meter.all_reduce()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total = int(meter.conf.sum())
    if rank == 0:
        out = {"workload": f"eval_sweep_{args.maps}x{H}x{W}_C{C}", "mode": args.mode, "n_gpus": world,
               "maps": args.maps, "ms": float(ms), "maps_per_s": args.maps / (float(ms) * 1e-3),
               "gbs_per_gpu": n_local * bytes_per_map / (float(ms) * 1e-3) / 1e9,
               "bytes_per_map": bytes_per_map, "resident_maps_per_gpu": res, "counted_pixels": total,
               "mIoU": float(np.nanmean(meter.metrics()["IoU"]))}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
