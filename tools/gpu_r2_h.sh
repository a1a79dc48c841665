#!/bin/bash
# 2-GPU validation: multi-GPU tests (peer sampling / remote rows in the fused kernel / fused all-reduce+Adam / replica cache),
# headline bench at N=2 (cache off) and the per-kernel graph timeline of rank 0
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m "gpu and multigpu" > gpurun_out/h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/h_tests.log
tail -8 gpurun_out/h_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 300 --warmup 5 > gpurun_out/h_bench2.log 2>&1
echo "bench rc=$?" >> gpurun_out/h_bench2.log
tail -3 gpurun_out/h_bench2.log | cut -c1-1500
timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 > gpurun_out/h_bench1.log 2>&1
tail -1 gpurun_out/h_bench1.log | cut -c1-800
