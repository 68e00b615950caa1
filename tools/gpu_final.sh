#!/bin/bash
# Round-end validation in one short gpurun call: full GPU test suite, smoke, the contract bench,
# the per-kernel table of the kernels added last, the evaluation sweep. Outputs: gpurun_out/<tag>_*.
tag=${1:-r1r}
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${tag}_pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 120 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
timeout 90 python tools/kbench.py --only evallogits,blur --iters 10 > gpurun_out/${tag}_kbench_new.jsonl 2>&1; echo "kbench rc=$?"
grep "^{" gpurun_out/${tag}_kbench_new.jsonl
timeout 60 python tools/eval_sweep.py --mode labels > gpurun_out/${tag}_eval_sweep.jsonl 2>&1; echo "sweep rc=$?"
timeout 60 python tools/eval_sweep.py --mode logits --maps 1250 >> gpurun_out/${tag}_eval_sweep.jsonl 2>&1; echo "sweep logits rc=$?"
tail -2 gpurun_out/${tag}_eval_sweep.jsonl
