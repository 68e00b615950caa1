timeout 300 python -m pytest tests/test_gpu_weighted_ce.py -m gpu -q 2>&1 | tail -3
PFST_CE_TY=16 timeout 300 python -m pytest tests/test_gpu_weighted_ce.py -m gpu -q 2>&1 | tail -3
for ty in 8 16; do
  export PFST_CE_TY=$ty
  echo "== TY=$ty"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:weighted_ce --csv python tools/kbench.py --only ce --no-graph --iters 3 --workload cfg2 2>/dev/null | grep weighted_ce_s4 | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
