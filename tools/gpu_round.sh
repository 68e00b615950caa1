#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (ours + reference arm), per-kernel table, ncu launch
# list and one --set full capture of a whole step. Outputs under gpurun_out/<tag>_*.
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${tag}_bench_ref.json
python tools/step_profile.py cfg2 > gpurun_out/${tag}_step_profile.txt 2>&1; echo "step_profile rc=$?"
python tools/kbench.py > gpurun_out/${tag}_kbench.jsonl 2>&1; echo "kbench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2> gpurun_out/${tag}_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu bench launches rc=$?"
python tools/one_step.py cfg2 2 > gpurun_out/${tag}_one_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/one_step.py cfg2 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/one_step.py cfg2 1 > gpurun_out/${tag}_one_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kernel \
    -o gpurun_out/${tag}_step_full python tools/one_step.py cfg2 1 > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/eval_sweep.py --mode labels > gpurun_out/${tag}_eval_sweep.jsonl 2>&1; python tools/eval_sweep.py --mode logits --maps 1250 >> gpurun_out/${tag}_eval_sweep.jsonl 2>&1
echo "eval sweep rc=$?"
python tools/ncu_new.py > gpurun_out/${tag}_ncu_new.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gaussian_blur|argmax_confusion" --launch-skip 2 -c 2 \
    -o gpurun_out/${tag}_new_kernels python tools/ncu_new.py > gpurun_out/${tag}_ncu_new_full.log 2>&1
echo "ncu new kernels rc=$?"
