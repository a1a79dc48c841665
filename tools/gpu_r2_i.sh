#!/bin/bash
# role-layout variants of the persistent fused kernel: 1-GPU step time + per-kernel timeline, then 2 GPUs (remote rows)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/i_tests.log 2>&1; tail -3 gpurun_out/i_tests.log
for v in 0 1 2; do
  GLB_SAGE_VARIANT=$v timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 --no-secondary > gpurun_out/i_bench1_v$v.log 2>&1
  echo "v$v 1gpu: $(tail -1 gpurun_out/i_bench1_v$v.log | cut -c1-260)"
  GLB_SAGE_VARIANT=$v timeout 300 python tools/graph_timeline.py > gpurun_out/i_timeline_v$v.log 2>&1
  grep -A12 "timeline of replay" gpurun_out/i_timeline_v$v.log | tail -11
done
for v in 0 1 2; do
  GLB_SAGE_VARIANT=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2972$v bench.py --gpus 2 --steps 300 --warmup 5 --no-secondary > gpurun_out/i_bench2_v$v.log 2>&1
  echo "v$v 2gpu: $(tail -1 gpurun_out/i_bench2_v$v.log | cut -c1-260)"
done
timeout 900 python -m pytest tests -q -m "gpu and multigpu" > gpurun_out/i_mtests.log 2>&1; tail -3 gpurun_out/i_mtests.log
