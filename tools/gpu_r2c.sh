#!/bin/bash
# r2c (1 GPU): plugin-path tests, host profile of forward_train, bench cfg2 with the re-ordered step DAG.
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pfgst_step.py tests/test_gpu_step_fused.py tests/test_gpu_prototypes.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${tag}_pytest.log
timeout 300 python tools/plugin_profile.py cfg2 > gpurun_out/${tag}_plugin_profile.txt 2>&1; echo "profile rc=$?"; head -3 gpurun_out/${tag}_plugin_profile.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 300 python tools/step_profile.py cfg2 > gpurun_out/${tag}_step_profile.txt 2>&1; echo "step_profile rc=$?"; tail -30 gpurun_out/${tag}_step_profile.txt
