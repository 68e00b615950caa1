#!/bin/bash
# r2p (8 GPUs): N>1 parity on 8 ranks and the cfg2 weak-scaling line with the background EMA.
tag=${1:-r2p}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi_rank.py -m gpu -x -q -rs -k "8-peer" > gpurun_out/${tag}_pytest_multi_8gpu.log 2>&1; echo "pytest multi rc=$?"
tail -3 gpurun_out/${tag}_pytest_multi_8gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/${tag}_bench_cfg2_n8.json 2> gpurun_out/${tag}_bench_cfg2_n8.err
echo "bench N=8 rc=$?"; grep "^{" gpurun_out/${tag}_bench_cfg2_n8.json | cut -c1-330
