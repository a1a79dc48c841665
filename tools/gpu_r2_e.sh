#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/e_tests.log
tail -6 gpurun_out/e_tests.log
timeout 600 python bench.py --steps 1000 --warmup 10 > gpurun_out/e_bench.log 2>&1
echo "bench rc=$?"; tail -3 gpurun_out/e_bench.log | cut -c1-1800
