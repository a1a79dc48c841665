#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t6.log 2>&1; echo "t6 exit $?" >> gpurun_out/t6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --feature-dtype bf16 > gpurun_out/bench_bf16.log 2>&1; echo "exit $?" >> gpurun_out/bench_bf16.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sage_fused -s 2 -c 1 -o gpurun_out/sage_final python tools/one_gather.py > gpurun_out/prof_full.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref.log
for f in t6 smoke bench bench_bf16 bench_ref; do echo "=== $f"; grep -v "^\[20\|Warn\|warn" gpurun_out/$f.log | tail -n 5 | cut -c1-600; done
