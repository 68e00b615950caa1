#!/bin/bash
# r2m (1 GPU), state at the end of round 2: full GPU suite, smoke, bench lines of every BASELINE config
# (ours + reference arm), per-kernel tables (cfg2, cfg4), ncu launch list of the bench command, one
# --set full capture of an eager step (top kernels) and the step timeline.
tag=${1:-r2m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench cfg2 rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_cfg2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref_cfg2.json 2> gpurun_out/${tag}_bench_ref_cfg2.err; echo "ref cfg2 rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_ref_cfg2.json
for w in cfg1 cfg3 cfg4 cfg5; do
  steps=200; [ $w = cfg5 ] && steps=40
  timeout 600 python bench.py --workload $w --steps $steps > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err; echo "bench $w rc=$?"
  cut -c1-250 gpurun_out/${tag}_bench_$w.json; tail -3 gpurun_out/${tag}_bench_$w.err
  timeout 600 python bench.py --workload $w --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_$w.json 2> gpurun_out/${tag}_bench_ref_$w.err; echo "ref $w rc=$?"
done
timeout 600 python tools/kbench.py > gpurun_out/${tag}_kbench.jsonl 2>&1; echo "kbench rc=$?"
timeout 300 python tools/kbench.py --workload cfg4 --only pl,mix,feat > gpurun_out/${tag}_kbench_cfg4.jsonl 2>&1; echo "kbench cfg4 rc=$?"
for w in cfg2 cfg4; do timeout 300 python tools/step_profile.py $w > gpurun_out/${tag}_step_profile_$w.txt 2>&1; done
timeout 300 python tools/timeline.py cfg2 > gpurun_out/${tag}_timeline_cfg2.txt 2>&1
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu bench launches rc=$?"
timeout 200 python tools/one_step.py cfg2 1 > gpurun_out/${tag}_one_step.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kernel \
    -o gpurun_out/${tag}_step_full -f python tools/one_step.py cfg2 1 > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
