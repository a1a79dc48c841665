#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gsl_engine_gpu.py tests/test_cpp_api_gpu.py -q > gpurun_out/l_tests.log 2>&1; tail -12 gpurun_out/l_tests.log
for pf in 1 0; do
  GLB_SAGE_L2PF=$pf timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 --no-secondary > gpurun_out/l_bench1_pf$pf.log 2>&1
  echo "pf$pf 1gpu: $(tail -1 gpurun_out/l_bench1_pf$pf.log | cut -c1-260)"
  GLB_SAGE_L2PF=$pf timeout 300 python tools/graph_timeline.py > gpurun_out/l_timeline_pf$pf.log 2>&1
  grep -A12 "timeline of replay" gpurun_out/l_timeline_pf$pf.log | tail -11
done
