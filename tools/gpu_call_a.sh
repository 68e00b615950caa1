#!/bin/bash
# Short gpurun call: only what changed since the last full round, so that it fits a small
# GPU-minute budget. Usage: tools/gpu_call_a.sh <tag> "<pytest files>" "<kbench --only list>"
tag=${1:-r1n}
files=${2:-"tests/test_gpu_eval_logits.py tests/test_gpu_blur.py tests/test_gpu_golden.py tests/test_gpu_classmix.py"}
only=${3:-"evallogits,blur"}
mkdir -p gpurun_out
timeout 200 python -m pytest $files -m gpu -q > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"
tail -25 gpurun_out/${tag}_pytest_new.log
timeout 120 python tools/kbench.py --only $only --iters 10 > gpurun_out/${tag}_kbench_new.jsonl 2>&1; echo "kbench rc=$?"
cat gpurun_out/${tag}_kbench_new.jsonl | tail -12
