#!/bin/bash
# ncu --set full of the kernels matching $1 in `tools/kbench.py --only $2` (plain launches).
# usage: bash tools/prof_kernel.sh <kernel-regex> <kbench-group> <tag> [workload]
pat=$1; grp=$2; tag=$3; wl=${4:-cfg2}
mkdir -p gpurun_out
python tools/kbench.py --only $grp --no-graph --iters 2 --workload $wl > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$pat -c 4 \
    -o gpurun_out/${tag} -f python tools/kbench.py --only $grp --no-graph --iters 2 --workload $wl > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu.log
