#!/bin/bash
# round 2, run A: validate the persistent fused kernel (tests + short bench)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "sage or trainer or engine or fast or linear" > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -15 gpurun_out/a_tests.log
timeout 600 python bench.py --steps 500 --warmup 10 > gpurun_out/a_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/a_bench.log
tail -3 gpurun_out/a_bench.log
