for v in 1 2 3 4; do
  export PFST_EMA_BLOCKS_PER_SM=$v
  for w in cfg2 cfg3; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ema_blocks=$v', d[\"config\"][\"workload\"], round(d[\"ms_per_step\"]*1000,1), round(d[\"step_frac_of_peak\"],3), round(d[\"roofline\"][\"frac\"],3))"
  done
done
