#!/bin/bash
# N-GPU pass: bench with default flags (partitioned table, 128-byte aligned rows; + replica-cache secondary line), dense-row A/B,
# per-rank graph timeline, multi-GPU tests.   gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_multi.sh N'
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
summ() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f steps/s  ms/step %.4f  e2e %.0f  replica %s  remoteGBps %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('replica_cache_run') or {}).get('value'), d['config'].get('remote_feature_GBps_per_gpu')))"; }
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 300 --warmup 5 "${@:3}" > gpurun_out/$2 2>&1; echo "$2 rc=$?"; summ gpurun_out/$2; }
run 29821 m_bench$N.log
GLB_FEATURE_ROW_ALIGN=16 run 29823 m_bench${N}_dense.log --no-secondary
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29822 tools/graph_timeline.py > gpurun_out/m_timeline$N.log 2>&1
grep -A30 "timeline of replay" gpurun_out/m_timeline$N.log | head -34
timeout 900 python -m pytest tests -q -m "gpu and multigpu" > gpurun_out/m_mtests$N.log 2>&1; tail -3 gpurun_out/m_mtests$N.log
