timeout 600 python -m pytest tests/test_gpu_pfgst_loss.py tests/test_gpu_step_fused.py tests/test_gpu_full_size.py tests/test_gpu_golden.py tests/test_gpu_pfgst_step.py -m gpu -x -q 2>&1 | tail -3
for w in cfg2 cfg4; do
  timeout 200 python tools/kbench.py --workload $w --only feat --iters 10 2>&1 | grep '^{' | grep -i "loss" | cut -c1-170
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"config\"][\"workload\"], d[\"ms_per_step\"], d[\"step_frac_of_peak\"], d.get(\"plugin\",{}).get(\"ms_per_step\"))"
done
