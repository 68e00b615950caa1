timeout 300 python -m pytest tests/test_gpu_step_fused.py -m gpu -x -q 2>&1 | tail -1
for cfgv in "0:0" "1:0" "1:4" "1:8"; do
  export PFST_EMA_BACKGROUND=${cfgv%%:*} PFST_EMA_BG_BLOCKS=${cfgv##*:}
  for w in cfg2 cfg1 cfg3; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bg=$cfgv', d[\"config\"][\"workload\"], round(d[\"ms_per_step\"]*1000,1), round(d[\"step_frac_of_peak\"],3))"
  done
done
