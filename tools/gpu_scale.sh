#!/bin/bash
# lean N-GPU scaling point: bench with default flags (+ replica-cache secondary line) and the per-rank graph timeline
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/s_bench$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/s_bench$N.log | cut -c1-2600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29832 tools/graph_timeline.py > gpurun_out/s_timeline$N.log 2>&1
grep -A40 "timeline of replay" gpurun_out/s_timeline$N.log | head -36
