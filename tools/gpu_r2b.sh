#!/bin/bash
# r2b (1 GPU): full GPU test suite, smoke, bench cfg2 (value / plugin / e2e) + the reference arm,
# per-kernel table, and the ncu launch list of a short bench run (at:: kernels of the plugin leg).
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${tag}_bench_ref.json
timeout 600 python tools/kbench.py > gpurun_out/${tag}_kbench.jsonl 2>&1; echo "kbench rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu bench launches rc=$?"
