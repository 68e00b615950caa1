#!/bin/bash
# r2h (1 GPU): whole GPU suite, plugin host profile, bench cfg2, kbench cfg2.
tag=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${tag}_pytest.log
timeout 300 python tools/plugin_profile.py cfg2 > gpurun_out/${tag}_plugin_profile.txt 2>&1; echo "profile rc=$?"; head -12 gpurun_out/${tag}_plugin_profile.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 300 python tools/kbench.py --workload cfg2 --iters 10 > gpurun_out/${tag}_kbench.jsonl 2>&1; echo "kbench rc=$?"; grep '^{' gpurun_out/${tag}_kbench.jsonl | cut -c1-220
