#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "engine or trainer" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_dist.py -m gpu -x -q -k "two_ranks_gpu_peer" 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 1500 --warmup 20 > gpurun_out/bench2_v7.log 2>&1
grep '"metric"' gpurun_out/bench2_v7.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['final_loss'])" || tail -5 gpurun_out/bench2_v7.log
timeout 200 python bench.py --steps 2000 > gpurun_out/bench1_v7.log 2>&1
grep '"metric"' gpurun_out/bench1_v7.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['final_loss'])" || tail -5 gpurun_out/bench1_v7.log
