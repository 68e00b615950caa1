#!/usr/bin/env python
"""bench.py — PFST self-training hot-path throughput on B200 (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (EMA teacher update, pseudo labels, ClassMix,
PFGST loss fwd+bwd, prototype accumulate/all-reduce/distance fwd+bwd) over one batch of
synthetic network outputs, excluding the three network passes.
Rank 0 prints ONE JSON line:
  value          pixels/s, whole job, inputs resident in HBM, CUDA-event timed
  e2e            same metric through the PFGST plugin class with HOST (pinned) buffers:
                 H2D of every step input and D2H of the log vars inside the timed region
  roofline       dominant kernel (multi-tensor EMA, 12 B/param): bytes / CUDA-event time
                 of that launch inside the timed steps, vs MEASURED_PEAKS.json
  cpu_baseline   the CPU oracle (a port of the reference's path) on this box's host cores
`--impl reference` times that CPU port alone, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from pfst_b200.synthetic import WORKLOADS, model_params, step_inputs  # noqa: E402

METRIC = "self_training_hot_path_throughput"
UNIT = "pixels/s"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arms
_CPU_STATE = {}


def cpu_step_time(wl, sample_b, seed):
    """Oracle step on `sample_b` images of the workload + the FULL EMA (it does not scale with
    the batch). Returns (t_ema, t_rest_for_the_sample_batch) in seconds."""
    from oracle import pfgst_loss as OL, step as ostep
    key = (wl.name, sample_b, seed)
    if key not in _CPU_STATE:
        g = torch.Generator().manual_seed(seed)
        inp = {k: v[:sample_b].contiguous() for k, v in step_inputs(wl, seed).items()}
        _CPU_STATE[key] = (inp, model_params(wl.C, g), model_params(wl.C, g), np.random.RandomState(seed))
    inp, student, teacher, rs = _CPU_STATE[key]
    cfg = OL.LossCfg(dilation=wl.dilation, downscale=wl.downscale if wl.downscale != 1.0 else None)
    tm = {}
    ostep.hot_path_step(5000, teacher, student, inp, wl.C, loss_cfg=cfg, rng=rs, timings=tm)
    return tm["ema"], tm["pseudo_mix"] + tm["loss_proto"]


def run_reference(args, wl, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sample_b = max(1, min(wl.B, args.ref_sample_images))
    scale = wl.B / sample_b
    for _ in range(args.warmup):
        cpu_step_time(wl, sample_b, 1234)
    ts = []
    for _ in range(args.steps):
        t_ema, t_rest = cpu_step_time(wl, sample_b, 1234)
        ts.append(t_ema + t_rest * scale)      # extrapolated time of the full step
    t = sum(ts) / len(ts)
    value = wl.B * wl.H * wl.W / t    # one host: its cores do not multiply with the GPU count
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "batch_per_gpu": wl.B, "classes": wl.C, "image": [wl.H, wl.W],
                       "feature_dim": wl.D},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"oracle (ATen-op-for-op port of the reference path, torch {torch.__version__} "
                                       f"CPU) on {sample_b}/{wl.B} images per step + the full 43.6 M-parameter EMA; "
                                       f"batch-dependent time scaled x{scale:g}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
class HostFeed:
    """The e2e leg's inputs live in pinned HOST memory. Every step's tensors (the data batch and
    the network outputs the replay segmentor 'produces') cross PCIe inside the timed region; the
    copies of step k+1 run on a copy stream while step k computes (two device slots)."""

    def __init__(self, pinned: dict, device):
        self.pinned = pinned
        self.slots = [{k: torch.empty_like(v, device=device) for k, v in pinned.items()} for _ in range(2)]
        self.stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.cur = 0
        self.bytes_per_step = sum(v.numel() * v.element_size() for v in pinned.values())
        self.steps = 0

    def issue(self, slot: int) -> None:
        if self.used[slot]:
            self.stream.wait_event(self.free[slot])        # the step that read this slot has finished
        with torch.cuda.stream(self.stream):
            for k, v in self.pinned.items():
                self.slots[slot][k].copy_(v, non_blocking=True)
        self.ready[slot].record(self.stream)
        self.steps += 1

    def acquire(self, slot: int) -> dict:
        torch.cuda.current_stream().wait_event(self.ready[slot])
        self.cur = slot
        return self.slots[slot]

    def release(self, slot: int) -> None:
        self.free[slot].record(torch.cuda.current_stream())
        self.used[slot] = True


class ReplaySegmentor(torch.nn.Module):
    """Stands in for the DeepLabV3+ R50-D8 segmentor in the e2e leg: it owns a parameter list
    of the real shapes (so the EMA runs over the true 214 tensors) and 'produces' the network
    outputs from the current HostFeed slot — tensors that crossed PCIe for this step."""
    FEED = None      # class attribute: shared by the student and its deep-copied teacher

    def __init__(self, wl, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.params = torch.nn.ParameterList([torch.nn.Parameter(p) for p in model_params(wl.C, g)])
        self.calls = 0
        self.num_classes = wl.C
        self.train_cfg, self.test_cfg = {}, {}

    def _fetch(self, key):
        feed = ReplaySegmentor.FEED
        return feed.slots[feed.cur][key]

    def encode_decode(self, img, img_metas):
        logits = self._fetch("ema_logits")
        feats = self._fetch("x_ema")
        return logits, {"feats": feats, "decoded_features": feats, "seg_logits": logits}

    def forward_train(self, img, img_metas, gt, seg_weight=None, return_feats=False, return_decoded_feats=False,
                      return_logits=False, return_states=False):
        first = self.calls % 2 == 0
        self.calls += 1
        out = {"decode.loss_ce": self.params[-1].sum() * 0.0}
        if first:     # source pass: decoded features + logits of the source image
            x = self._fetch("x_src").detach().requires_grad_(True)
            out.update(features=x, decoded_features=x, logits=self._fetch("logits_src"))
        else:         # mixed pass: logits of the mixed image (gradient target of the sim losses)
            out.update(features=None, logits=self._fetch("logits_trg").detach().requires_grad_(True))
        return out


def run_ours(args, wl, rank, world, local_rank):
    import torch.distributed as dist
    from pfst_b200 import ops
    from pfst_b200.step import SelfTrainingStep, algorithmic_bytes
    from pfst_b200.uda import PFGST

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ops.device_check()
    seed = 1234 + rank
    g = torch.Generator().manual_seed(seed)
    np.random.seed(seed)
    host = step_inputs(wl, seed)
    inp = {k: v.to(dev) for k, v in host.items()}
    student = [p.to(dev) for p in model_params(wl.C, g)]
    teacher = [p.to(dev) for p in model_params(wl.C, g)]
    n_params = sum(p.numel() for p in student)
    step = SelfTrainingStep(teacher, student, wl.C, wl.D, dev, dilation=wl.dilation,
                            downscale=wl.downscale if wl.downscale != 1.0 else None, max_batch=max(wl.B, 64),
                            graphs=not args.no_graphs)

    def one(it):
        out = step.run(it, inp["img"], inp["target_img_strong_aug"], inp["gt"], inp["ema_logits"],
                       inp["logits_trg"], inp["x_src"], inp["x_ema"])
        # the next batch's labels are known before its network passes: their class-presence bits
        # (M1 part 1, one kernel + a 36-byte D2H) are prefetched so the host never waits for them
        step.prefetch(inp["gt"])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        one(i)
    sampler = ClockSampler(local_rank)       # sampled on rank 0 only (one nvidia-smi poller per job)
    if rank != 0:
        sampler.start = lambda: None
    ema_pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                 for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    t_wall0 = time.time()
    ev0.record()
    for k in range(args.steps):
        step.ema_events = ema_pairs[k]
        one(args.warmup + k + 1)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    step.ema_events = None
    ms = ev0.elapsed_time(ev1)
    try:
        ema_overlapped_ms = statistics.mean(a.elapsed_time(b) for a, b in ema_pairs)
    except (ValueError, RuntimeError):     # the EMA is a CUDA-graph node: no per-launch events inside the step
        ema_overlapped_ms = None
    # The dominant kernel (multi-tensor EMA) runs on its own stream inside the step and shares
    # HBM with the other kernels there, so its roofline point is taken from launches of the
    # SAME kernel/grid timed alone, on the launching stream (523 MB per launch: nothing fits in L2).
    coeffs = ops.ema_coeffs(5000, 0.999)
    alone = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for _ in range(3):
        step.table.update(*coeffs, blocks_per_sm=step.ema_blocks_per_sm)
    torch.cuda.synchronize()
    for a, b in alone:
        a.record()
        step.table.update(*coeffs, blocks_per_sm=step.ema_blocks_per_sm)
        b.record()
    torch.cuda.synchronize()
    ema_ms = statistics.mean(a.elapsed_time(b) for a, b in alone)

    # ---- e2e: PFGST plugin class, host (pinned) buffers in, log vars out ------------------
    pinned = {k: v.pin_memory() for k, v in host.items()}
    pinned["logits_src"] = (2.0 * torch.randn(host["logits_trg"].shape, generator=g)).pin_memory()
    pinned["target_img"] = torch.randn(host["img"].shape, generator=g).pin_memory()
    net_keys = ("ema_logits", "x_ema", "x_src", "logits_src", "logits_trg")
    batch_keys = ("img", "gt", "target_img", "target_img_strong_aug")
    feed = HostFeed({k: pinned[k] for k in net_keys + batch_keys}, dev)
    ReplaySegmentor.FEED = feed
    factory = lambda: ReplaySegmentor(wl, seed)  # noqa: E731
    model = PFGST(model=factory, max_iters=40000, alpha=0.999, pseudo_threshold=0.98, pseudo_weight_ignore_top=0,
                  pseudo_weight_ignore_bottom=0, imnet_feature_dist_lambda=0, imnet_feature_dist_classes=None,
                  imnet_feature_dist_scale_min_ratio=None, mix='class', blur=False, color_jitter_strength=0.2,
                  color_jitter_probability=1.0, print_grad_magnitude=False, trg_loss_weight=1.,
                  use_decoded_feats=True, thre_type='all', compute_vis=False,
                  prototypes=dict(weight=0.1),
                  aux_losses=[dict(type='PFGSTLoss', kernel_size=3, dilation=wl.dilation, top_k=3,
                                   weights={'src_pos': 0.1, 'src_neg': 0.1, 'sim_pos': 0.1, 'sim_neg': 0.1,
                                            'src_pos_std': 0.1, 'src_neg_std': 0.1},
                                   sim_type='cosine', feat_level=None, detach_unfold=True,
                                   downscale=wl.downscale if wl.downscale != 1.0 else None)]).to(dev)
    metas = [{'img_norm_cfg': {'mean': [0., 0., 0.], 'std': [1., 1., 1.]}}] * wl.B
    step_no = [0]

    def e2e_step():
        slot = step_no[0] % 2
        step_no[0] += 1
        d = feed.acquire(slot)                    # this step's inputs have crossed PCIe
        feed.issue(slot ^ 1)                      # the next step's copies overlap this step's compute
        log_vars, _ = model.forward_train(d["img"], metas, d["gt"], d["target_img"], metas,
                                          d["target_img_strong_aug"])      # log vars arrive via D2H
        feed.release(slot)
        return log_vars

    e2e_steps = max(3, min(args.steps, 20))
    feed.issue(0)
    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        lv = e2e_step()
    torch.cuda.current_stream().wait_stream(feed.stream)   # the look-ahead copy belongs to the timed region too
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    h2d = feed.bytes_per_step                     # one full input set is copied per step
    d2h = 4 * (len(lv) + 3) + 36                  # three stacked log-var vectors + presence bits

    times = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    reduce_mode = "peer-board (one-shot NVLink all-reduce inside proto_finalize)" if step.bank.peer is not None else \
        ("nccl all-reduce" if world > 1 else "none (single rank)")
    if step.bank.peer is not None:
        step.bank.peer.check()
    step.close()                   # collective: graphs and peer boards go before the process group
    if rank != 0:
        return
    px = wl.B * wl.H * wl.W
    value = px * world * args.steps / (ms * 1e-3)
    e2e_value = px * world * e2e_steps / (e2e_ms * 1e-3)
    peak, peak_src = measured_peak()
    ab = algorithmic_bytes(wl.B, wl.C, wl.H, wl.W, wl.D, inp["x_src"].shape[2], inp["x_src"].shape[3], n_params)
    achieved = ab["ema"] / (ema_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("ema_multi_kernel")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "batch_per_gpu": wl.B, "classes": wl.C, "image": [wl.H, wl.W],
                       "feature_dim": wl.D, "params": n_params,
                       "l2": "no flush: per-step working set (params 349 MB + maps 0.5 GB) exceeds the 126 MB L2",
                       "step": "EMA + pseudo-label + ClassMix + PFGST loss fwd/bwd + prototypes fwd/bwd (+ all-reduce)",
                       "cuda_graphs": not args.no_graphs, "proto_reduce": reduce_mode},
            "iters_per_s": world * args.steps / (ms * 1e-3) / world,
            "step_bytes": sum(ab.values()), "step_gbs_per_gpu": sum(ab.values()) / (ms / args.steps * 1e-3) / 1e9,
            "step_frac_of_peak": sum(ab.values()) / (ms / args.steps * 1e-3) / 1e9 / peak,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "api": "pfst_b200.uda.PFGST.forward_train (plugin class) with pinned host buffers; copies of step k+1 overlap step k (copy stream, two device slots)"},
            "gpu_launches": SelfTrainingStep.KERNEL_LAUNCHES * args.steps,
            "roofline": {"bound": "hbm", "kernel": "ema_multi_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": ab["ema"], "us_per_launch": ema_ms * 1e3,
                         "timed": "20 launches alone after the timed steps (same grid: %d blocks/SM)" % step.ema_blocks_per_sm,
                         "us_per_launch_overlapped_in_step": None if ema_overlapped_ms is None else ema_overlapped_ms * 1e3}}
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sample_b = max(1, min(wl.B, args.ref_sample_images))
        cpu_step_time(wl, sample_b, 1234)                       # warm-up (allocator, thread pools)
        t_ema, t_rest = cpu_step_time(wl, sample_b, 1234)
        t_full = t_ema + t_rest * wl.B / sample_b
        line["cpu_baseline"] = {"value": px / t_full, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"oracle step on {sample_b}/{wl.B} images + full EMA "
                                          f"({t_ema * 1e3:.0f} ms EMA, {t_rest:.2f} s rest), rest scaled x{wl.B / sample_b:g}"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-sample-images", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly (no CUDA-graph segments)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, wl, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
