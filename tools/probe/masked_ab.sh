for rep in 1 2; do for v in 0 1; do
  unset PFST_ACCUM_NO_MASKED; [ $v = 1 ] && export PFST_ACCUM_NO_MASKED=1
  for w in cfg2 cfg1; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no_masked=$v', d[\"config\"][\"workload\"], round(d[\"ms_per_step\"]*1000,1), round(d[\"step_frac_of_peak\"],3), round(d[\"plugin\"][\"ms_per_step\"]*1000,1))"
  done
done; done
