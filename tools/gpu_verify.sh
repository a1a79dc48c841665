#!/bin/bash
# quick one-GPU verification: full GPU test tier + headline bench (incl. the fp32 / fp8 row-storage secondary lines)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/o_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/o_tests.log
timeout 600 python bench.py --steps 3000 --warmup 20 > gpurun_out/o_bench1.log 2>&1; tail -1 gpurun_out/o_bench1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value'], {k: v['value'] for k, v in d.items() if k.endswith('_feature_rows_run')})"
