#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/f_tests.log
tail -12 gpurun_out/f_tests.log
timeout 300 python bench.py --steps 1000 --warmup 10 --no-secondary > gpurun_out/f_bench.log 2>&1
tail -1 gpurun_out/f_bench.log | cut -c1-250
timeout 600 python bench.py --config sage3 --steps 200 --warmup 5 --no-secondary > gpurun_out/f_sage3.log 2>&1
tail -1 gpurun_out/f_sage3.log | cut -c1-400
timeout 600 python bench.py --config deepwalk --steps 100 --warmup 5 > gpurun_out/f_deepwalk.log 2>&1
tail -1 gpurun_out/f_deepwalk.log | cut -c1-600
timeout 600 python tools/bench_dgs.py > gpurun_out/f_dgs.log 2>&1
tail -1 gpurun_out/f_dgs.log | cut -c1-900
