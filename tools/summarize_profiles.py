"""Turns gpurun_out/<tag>_launches.csv (ncu gpu__time_duration pass) and, when present,
gpurun_out/<tag>_step_full.ncu-rep (ncu --set full) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> <out-name>
    python tools/summarize_profiles.py --rep <file.ncu-rep> <out-name>     # any --set full capture; leaves traffic.json alone
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
only_rep = None
if sys.argv[1] == "--rep":
    only_rep, tag, out = Path(sys.argv[2]), "", sys.argv[3]
else:
    tag, out = sys.argv[1], sys.argv[2]
src = ROOT / "gpurun_out"
dst = ROOT / "profiles"
dst.mkdir(exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.pct",
        "smsp__average_warp_latency_issue_stalled_barrier.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]

for lc, suffix, title in ((src / f"{tag}_bench_launches.csv", "bench_launches",
                           "python bench.py --steps 2 --warmup 3 --no-cpu-baseline (device leg, then the e2e leg)"),
                          (src / f"{tag}_launches.csv", "launches", "tools/one_step.py cfg2 2 (eager launches of two steps)")):
    if only_rep is not None or not lc.exists():
        continue
    lines = [l for l in lc.read_text().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    per = defaultdict(list)
    for r in rows:
        per[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]) / 1e3)
    tot = sum(sum(v) for v in per.values())
    md = [f"# ncu launch list `{tag}`: {title} (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)",
          "", "| kernel | launches | avg us | share of step |", "|---|---|---|---|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        md.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {sum(v) / tot:.3f} |")
    md.append(f"\nTotal kernel time: {tot:.1f} us over {len(rows)} launches.")
    (dst / f"{out}_{suffix}.md").write_text("\n".join(md) + "\n")
    (dst / f"{out}_{suffix}.csv").write_text("\n".join(lines) + "\n")
    print("\n".join(md))

rep = only_rep or src / f"{tag}_step_full.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    body = rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    table, traffic = [], {}
    for r in body:
        name = r[idx["Kernel Name"]].split("(")[0]
        rec = {"kernel": name}
        for k in KEYS:
            if k in idx:
                rec[k] = r[idx[k]]
        table.append(rec)
        try:
            unit_r, unit_w = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            t = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * mult.get(unit_r, 1) + \
                float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * mult.get(unit_w, 1)
            traffic[name.replace("void ", "").split("::")[-1].split("<")[0].strip()] = t
        except (KeyError, ValueError):
            pass
    (dst / f"{out}_full_summary.json").write_text(json.dumps({"units": dict(zip(hdr, rows[1])), "kernels": table}, indent=1))
    if only_rep is None:
        (dst / "traffic.json").write_text(json.dumps(traffic, indent=1))
    else:
        (dst / f"{out}_traffic.json").write_text(json.dumps(traffic, indent=1))
    for rec in table:
        print(rec)
