#!/bin/bash
# 8-GPU validation: scaling bench (cache off = headline, replica cache = secondary line), 8-rank peer test, per-rank timeline
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/s_bench$N.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/s_bench$N.log | cut -c1-2500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29812 tools/graph_timeline.py > gpurun_out/s_timeline$N.log 2>&1
grep -A40 "timeline of replay" gpurun_out/s_timeline$N.log | head -40
if [ "$N" = "8" ]; then timeout 600 python -m pytest tests/test_dist.py -q -k "eight_ranks" > gpurun_out/s_test8.log 2>&1; tail -2 gpurun_out/s_test8.log; fi
