#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sage or trainer" > gpurun_out/t2.log 2>&1; echo "t2 exit $?" >> gpurun_out/t2.log
timeout 1200 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sage_fused --profile-from-start off -c 3 \
  -o gpurun_out/sage_fused_v2 python tools/profile_step.py > gpurun_out/prof_full.log 2>&1
for f in t2 bench; do echo "=== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-600; done
