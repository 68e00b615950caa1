#!/bin/bash
# r2e (8 GPUs): N>1 parity on 8 ranks, bench cfg2 at N=8 and N=4 (weak scaling), cfg5 sweep sharded over 8.
tag=${1:-r2e}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 500 python -m pytest tests/test_gpu_multi_rank.py -m gpu -x -q -rs > gpurun_out/${tag}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -4 gpurun_out/${tag}_pytest_multi.log
run() { # n workload steps extra
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 \
    bench.py --gpus $1 --workload $2 --steps $3 --warmup 5 > gpurun_out/${tag}_bench_$2_n$1.json 2> gpurun_out/${tag}_bench_$2_n$1.err
  echo "bench $2 N=$1 rc=$?"; cut -c1-330 gpurun_out/${tag}_bench_$2_n$1.json; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${tag}_bench_$2_n$1.err | tail -3
}
run 8 cfg2 200
run 4 cfg2 200
run 8 cfg5 20
