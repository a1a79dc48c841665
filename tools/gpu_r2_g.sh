#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g_tests.log
tail -8 gpurun_out/g_tests.log
timeout 600 python bench.py --config sage3 --steps 200 --warmup 5 --no-secondary > gpurun_out/g_sage3.log 2>&1
tail -1 gpurun_out/g_sage3.log | cut -c1-400
timeout 600 python bench.py --config taobao_gat --steps 100 --warmup 5 > gpurun_out/g_gat.log 2>&1
tail -2 gpurun_out/g_gat.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sage_persist|sage_bwd_dw" -o gpurun_out/g_step -f python tools/profile_step.py > gpurun_out/g_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/g_launches.csv python tools/profile_step.py > gpurun_out/g_launches.log 2>&1
