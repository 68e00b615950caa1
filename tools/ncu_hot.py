"""Top stall locations of one kernel from an ncu report (source page, SASS view).
    python tools/ncu_hot.py <report.ncu-rep> <kernel-regex> [N]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several launches may match: split on "Kernel Name" rows, keep the first
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
end = start[1] if len(start) > 1 else len(rows)
hdr = rows[start[0] + 1]
body = [r for r in rows[start[0] + 2:end] if len(r) == len(hdr)]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[i_samp]) for r in body)
print(rows[start[0]][1][:100])
print("total samples", tot, "warp instr", sum(int(r[i_exec]) for r in body), "sass lines", len(body))
top = sorted(enumerate(body), key=lambda t: -int(t[1][i_samp]))[:n]
for idx, r in sorted(top):
    print(f"{idx:5d} {int(r[i_samp]):6d} {100 * int(r[i_samp]) / max(tot, 1):5.1f}% exec={r[i_exec]:>8s}  {r[i_src][:100]}")
