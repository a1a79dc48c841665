#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r_tests.log
timeout 600 python bench.py --config taobao_gat --steps 200 --warmup 5 > gpurun_out/r_gat.log 2>gpurun_out/r_gat.err; tail -1 gpurun_out/r_gat.log | cut -c1-1500; tail -3 gpurun_out/r_gat.err
