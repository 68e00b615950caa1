"""Per-kernel micro-benchmark: achieved HBM GB/s of every kernel on its BASELINE shape.

Development tool (bench.py is the contract benchmark). Each kernel is timed with
CUDA events on the launching stream, L2 flushed (256 MB memset) before every
timed launch, median of `--iters` launches after warm-up.

    python tools/kbench.py [--workload cfg2] [--iters 20] [--only ema,pl,...]
"""
from __future__ import annotations

import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from pfst_b200 import ops  # noqa: E402
from pfst_b200.synthetic import WORKLOADS, model_params, step_inputs, eval_maps  # noqa: E402


def peak_gbs() -> float:
    p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"])
    return 6650.0


class Timer:
    def __init__(self, dev):
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time(self, fn, iters=20, warmup=3):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(iters):
            self.flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e) * 1e-3)
        return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    dev = torch.device("cuda:0")
    ops.device_check()
    wl = WORKLOADS[args.workload]
    T = Timer(dev)
    peak = peak_gbs()
    rows = []

    def report(name, nbytes, fn):
        if only and name not in only:
            return
        med, best = T.time(fn, args.iters)
        rows.append(dict(kernel=name, bytes=nbytes, us_median=med * 1e6, us_best=best * 1e6,
                         gbs=nbytes / med / 1e9, frac_of_measured_peak=nbytes / med / 1e9 / peak))
        print(json.dumps(rows[-1]), flush=True)

    g = torch.Generator().manual_seed(1234)
    inp = {k: v.to(dev) for k, v in step_inputs(wl).items()}
    B, C, H, W = wl.B, wl.C, wl.H, wl.W
    P = B * H * W

    if not only or "ema" in only:
        student = [p.to(dev) for p in model_params(C, g)]
        teacher = [p.to(dev) for p in model_params(C, g)]
        table = ops.EmaTable(teacher, student)
        a, b = ops.ema_coeffs(5000, 0.999)
        report("ema", 12 * table.total, lambda: table.update(a, b))
        flat_t = torch.randn(table.total, device=dev)
        flat_s = torch.randn(table.total, device=dev)
        report("ema_flat", 12 * table.total, lambda: ops.ema_update_flat(flat_t, flat_s, a, b))
        del student, teacher, flat_t, flat_s

    report("pl", (4 * C + 12) * P, lambda: ops.pseudo_label(inp["ema_logits"], 0.98))
    lab, conf, count, _ = ops.pseudo_label(inp["ema_logits"], 0.98)
    report("presence", 8 * P, lambda: ops.class_presence(inp["gt"]))
    chosen = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    chosen[:, 0] = 0b010101
    report("mix", (8 + 24 + 8 + 12 + 8 + 4 + 8) * P,
           lambda: ops.class_mix(inp["gt"], chosen, inp["img"], inp["target_img_strong_aug"], lab,
                                 count=count, ps_size=P))

    if not only or "conf" in only:
        n = 64
        pred, gt = eval_maps(n, 1024, 1024, 6, seed=1)
        dp, dg = torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev)
        out = torch.zeros((1, 7, 7), dtype=torch.int64, device=dev)
        report("conf", 9 * n * 1024 * 1024, lambda: ops.confusion_accum(dp, dg, 6, out=out))
        dp8 = dp.to(torch.uint8)
        report("conf_u8", 2 * n * 1024 * 1024, lambda: ops.confusion_accum(dp8, dg, 6, out=out))
        outp = torch.zeros((n, 7, 7), dtype=torch.int64, device=dev)
        report("conf_per_image", 9 * n * 1024 * 1024,
               lambda: ops.confusion_accum(dp, dg, 6, per_image=True, out=outp))
    print(json.dumps({"peak_gbs": peak, "workload": wl.name}))


if __name__ == "__main__":
    main()
