#!/bin/bash
# r2f (1 GPU): offline-label tests; bench lines of every BASELINE config (ours + reference arm).
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_offline_labels.py tests/test_gpu_class_thresholds.py -m gpu -x -q > gpurun_out/${tag}_pytest_offline.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${tag}_pytest_offline.log
for w in cfg1 cfg3 cfg4 cfg5; do
  steps=200; [ $w = cfg5 ] && steps=40
  timeout 600 python bench.py --workload $w --steps $steps > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "bench $w rc=$?"
  cut -c1-250 gpurun_out/${tag}_bench_$w.json; tail -3 gpurun_out/${tag}_bench_$w.err
  timeout 600 python bench.py --workload $w --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_$w.json 2> gpurun_out/${tag}_bench_ref_$w.err; echo "ref $w rc=$?"
  cut -c1-200 gpurun_out/${tag}_bench_ref_$w.json
done
