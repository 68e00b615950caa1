timeout 600 python -m pytest tests/test_gpu_pfgst_loss.py tests/test_gpu_step_fused.py tests/test_gpu_full_size.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -4
for w in cfg2 cfg4; do
  echo "== $w"
  timeout 200 python tools/kbench.py --workload $w --only feat --iters 10 2>&1 | grep '^{' | grep -i "loss\|neigh\|proto" | cut -c1-170
  timeout 200 python tools/step_profile.py $w 2>&1 | grep -i "pfst::\|device time" | cut -c1-110
done
