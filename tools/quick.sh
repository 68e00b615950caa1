python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/kbench.py --only feat 2>&1 | tail -10
python tools/kbench.py --only feat --workload cfg3 2>&1 | grep -E "proto_accum|dist_fwd|dots|grad"
python tools/kbench.py --only feat --workload cfg4 2>&1 | grep -E "proto_accum|dist_fwd|dots|grad"
