timeout 600 python -m pytest tests/test_gpu_prototypes.py tests/test_gpu_step_fused.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do
  unset PFST_DIST_NO_SMALL; [ $v = 1 ] && export PFST_DIST_NO_SMALL=1
  echo "== NO_SMALL=$v"
  timeout 200 python tools/kbench.py --workload cfg4 --only feat --iters 10 2>&1 | grep '^{' | grep -i "dist_fwd" | cut -c1-170
  for r in 1 2; do timeout 300 python bench.py --workload cfg4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"config\"][\"workload\"], round(d[\"ms_per_step\"]*1000,1), round(d[\"step_frac_of_peak\"],3))"; done
done
