#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python tools/diag_sage.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/t5.log 2>&1; echo "t5 exit $?" >> gpurun_out/t5.log
timeout 1200 python bench.py --steps 300 --warmup 5 > gpurun_out/bench_fast.log 2>&1; echo "exit $?" >> gpurun_out/bench_fast.log
timeout 1200 python bench.py --steps 300 --warmup 5 --feature-dtype bf16 > gpurun_out/bench_fast_bf16.log 2>&1; echo "exit $?" >> gpurun_out/bench_fast_bf16.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fast.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
GLB_FDT=bf16 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fast_bf16.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step_bf16.log 2>&1
for f in diag t5 bench_fast bench_fast_bf16; do echo "=== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-500; done
