#!/bin/bash
# per-kernel device times of one eager step + full ncu capture of the fused kernel
mkdir -p gpurun_out
export PYTHONPATH=$PWD
# 3 warmup + capture-less eager run; skip the graph build + warm-up launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
echo "launches exit $?" >> gpurun_out/prof_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sage_fused --profile-from-start off -c 3 \
  -o gpurun_out/sage_fused python tools/profile_step.py > gpurun_out/prof_full.log 2>&1
echo "full exit $?" >> gpurun_out/prof_full.log
tail -5 gpurun_out/prof_step.log gpurun_out/prof_full.log
ls -la gpurun_out
