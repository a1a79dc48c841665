#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "tensor_core_aggregation" > gpurun_out/j_tc_test.log 2>&1; tail -15 gpurun_out/j_tc_test.log
for tc in 0 1; do
  GLB_SAGE_TCAGG=$tc timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 --no-secondary > gpurun_out/j_bench1_tc$tc.log 2>&1
  echo "tc$tc 1gpu: $(tail -1 gpurun_out/j_bench1_tc$tc.log | cut -c1-260)"
  GLB_SAGE_TCAGG=$tc timeout 300 python tools/graph_timeline.py > gpurun_out/j_timeline_tc$tc.log 2>&1
  grep -A12 "timeline of replay" gpurun_out/j_timeline_tc$tc.log | tail -11
done
GLB_SAGE_TCAGG=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gsl_engine_gpu.py -q -x > gpurun_out/j_tests_tc1.log 2>&1; tail -5 gpurun_out/j_tests_tc1.log
