#!/bin/bash
# r2a (2 GPUs): N>1 parity of the step in its three reduce forms, the single-rank peer-board test,
# and bench.py at N=2 with the peer board vs NCCL-in-graph. Logs under gpurun_out/r2a_*.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_prototypes.py -m gpu -x -q -rs > gpurun_out/${tag}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -8 gpurun_out/${tag}_pytest_multi.log
for mode in 1 0; do
  PFST_PEER_REDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/${tag}_bench_n2_peer${mode}.json 2> gpurun_out/${tag}_bench_n2_peer${mode}.err
  echo "bench N=2 peer=$mode rc=$?"; cut -c1-400 gpurun_out/${tag}_bench_n2_peer${mode}.json; tail -3 gpurun_out/${tag}_bench_n2_peer${mode}.err
done
timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench N=1 rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_n1.json
