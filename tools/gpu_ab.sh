#!/bin/bash
# One gpurun call: parity tests and the per-kernel table of a candidate library
# (tools/build_candidates.sh) next to the main build. Usage:
#   tools/gpu_ab.sh <tag> "<pytest files>" "<kbench --only list>"
tag=${1:-ab}
files=${2:-"tests/test_gpu_blur.py tests/test_gpu_weighted_ce.py tests/test_gpu_pseudo_label.py"}
only=${3:-"blur,ce,pl"}
cand=pfst_b200/csrc/libpfst_sm100_cand.so
mkdir -p gpurun_out
[ -f $cand ] || { echo "no $cand (run tools/build_candidates.sh first)"; exit 1; }
PFST_LIB=$PWD/$cand timeout 300 python -m pytest $files -m gpu -q > gpurun_out/${tag}_pytest_cand.log 2>&1; echo "pytest(candidate) rc=$?"
tail -15 gpurun_out/${tag}_pytest_cand.log
for w in cfg2 cfg4; do
  timeout 200 python tools/kbench.py --workload $w --only $only --iters 10 > gpurun_out/${tag}_kbench_main_$w.jsonl 2>&1; echo "kbench main $w rc=$?"
  PFST_LIB=$PWD/$cand timeout 200 python tools/kbench.py --workload $w --only $only --iters 10 > gpurun_out/${tag}_kbench_cand_$w.jsonl 2>&1; echo "kbench cand $w rc=$?"
  echo "== $w main"; grep '^{' gpurun_out/${tag}_kbench_main_$w.jsonl
  echo "== $w candidate"; grep '^{' gpurun_out/${tag}_kbench_cand_$w.jsonl
done
