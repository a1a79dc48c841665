#!/bin/bash
# 2-GPU: dist tests (with and without the replica cache) + bench A/B of the cache
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q -k "two_ranks" 2>&1 | tail -5
for c in -1 0 600000; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 1000 --warmup 20 --feature-cache-rows $c > gpurun_out/bench2_cache_$c.log 2>&1
grep '"metric"' gpurun_out/bench2_cache_$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cache', d['config']['feature_cache_rows_per_gpu'], 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'])" || tail -5 gpurun_out/bench2_cache_$c.log
done
