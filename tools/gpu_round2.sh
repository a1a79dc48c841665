#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python tools/diag_sage.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sage or trainer" > gpurun_out/t2.log 2>&1; echo "t2 exit $?" >> gpurun_out/t2.log
timeout 1200 python bench.py --steps 200 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
for f in diag t2 bench; do echo "=== $f"; tail -n 14 gpurun_out/$f.log; done
