#!/bin/bash
# round 2, run B: launch list + full ncu capture of the persistent kernel + bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/b_launches.csv python tools/profile_step.py > gpurun_out/b_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sage_persist -o gpurun_out/b_persist -f python tools/profile_step.py > gpurun_out/b_ncu.log 2>&1
timeout 600 python bench.py --steps 500 --warmup 10 > gpurun_out/b_bench.log 2>&1
tail -2 gpurun_out/b_bench.log | cut -c1-400
grep -v "^==" gpurun_out/b_launches.csv | cut -d, -f5,12- | head -40
