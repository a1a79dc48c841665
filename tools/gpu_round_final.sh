#!/bin/bash
# Round-end validation on ONE GPU: full gpu test-suite, smoke(), bench (default flags), reference arm (short),
# ncu launch list of one eager step + --set full of the fused kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
GLB_FDT=bf16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v7.csv \
  --profile-from-start off python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
GLB_FDT=bf16 timeout 400 ncu --set full --clock-control none --import-source on -k regex:sage_fused_fwd --profile-from-start off -c 3 \
  -o gpurun_out/sage_fused_v7 python tools/profile_step.py > gpurun_out/prof_full.log 2>&1
timeout 300 python bench.py > gpurun_out/bench1_final.log 2>&1
grep '"metric"' gpurun_out/bench1_final.log | cut -c1-400
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_v7.csv 2>&1 | tail -3
