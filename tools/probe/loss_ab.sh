timeout 600 python -m pytest tests/test_gpu_pfgst_loss.py tests/test_gpu_step_fused.py tests/test_gpu_full_size.py tests/test_gpu_golden.py tests/test_gpu_pfgst_step.py -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
  unset PFST_LOSS_NO_UP2; [ $v = 1 ] && export PFST_LOSS_NO_UP2=1
  echo "== NO_UP2=$v"
  timeout 200 python tools/kbench.py --workload cfg4 --only feat --iters 10 2>&1 | grep '^{' | grep -i "loss" | cut -c1-170
  timeout 300 python bench.py --workload cfg4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"config\"][\"workload\"], d[\"ms_per_step\"], d[\"step_frac_of_peak\"])"
done
