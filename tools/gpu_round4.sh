#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "fast_engine or random_walk or trainer" > gpurun_out/t4.log 2>&1; echo "t4 exit $?" >> gpurun_out/t4.log
timeout 1200 python bench.py --steps 300 --warmup 5 > gpurun_out/bench_fast.log 2>&1; echo "exit $?" >> gpurun_out/bench_fast.log
timeout 1200 python bench.py --steps 300 --warmup 5 --feature-dtype bf16 > gpurun_out/bench_fast_bf16.log 2>&1; echo "exit $?" >> gpurun_out/bench_fast_bf16.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fast.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
for f in t4 bench_fast bench_fast_bf16; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-700; done
