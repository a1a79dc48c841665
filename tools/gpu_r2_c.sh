#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "bwd_dw or sage or trainer or engine or fast" > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c_tests.log
tail -8 gpurun_out/c_tests.log
timeout 300 python tools/persist_timeline.py > gpurun_out/c_timeline.log 2>&1
echo "timeline rc=$?" >> gpurun_out/c_timeline.log
timeout 300 python bench.py --steps 500 --warmup 10 > gpurun_out/c_bench.log 2>&1
tail -1 gpurun_out/c_bench.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c_launches.csv python tools/profile_step.py > gpurun_out/c_launches.log 2>&1
