#!/bin/bash
# Short gpurun call: only what changed since the last full round (new kernels), so that it fits
# a small GPU-minute budget. Outputs under gpurun_out/<tag>_*.
tag=${1:-r1n}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_eval_logits.py tests/test_gpu_blur.py tests/test_gpu_golden.py \
    tests/test_gpu_classmix.py -m gpu -q > gpurun_out/${tag}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"
tail -25 gpurun_out/${tag}_pytest_new.log
timeout 120 python tools/kbench.py --only evallogits,blur --iters 10 > gpurun_out/${tag}_kbench_new.jsonl 2>&1; echo "kbench rc=$?"
cat gpurun_out/${tag}_kbench_new.jsonl | tail -12
timeout 100 python tools/eval_sweep.py --mode logits --maps 1250 > gpurun_out/${tag}_eval_sweep.jsonl 2>&1; echo "sweep logits rc=$?"
tail -3 gpurun_out/${tag}_eval_sweep.jsonl
