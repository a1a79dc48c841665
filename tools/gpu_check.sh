#!/bin/bash
# One-GPU validation pass: GPU tests, smoke, headline bench, secondary configs, reference arms, ncu (full set on the hot
# kernels + launch list).  Usage (from the repo root):  gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/n_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/n_smoke.log
timeout 600 python bench.py > gpurun_out/n_bench1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/n_bench1.log | cut -c1-400
for c in sage3 deepwalk; do
  timeout 600 python bench.py --config $c --steps 200 --warmup 5 --no-secondary > gpurun_out/n_$c.log 2>&1; tail -1 gpurun_out/n_$c.log | cut -c1-300
  timeout 900 python bench.py --impl reference --config $c --steps 20 --warmup 3 > gpurun_out/n_ref_$c.log 2>&1; tail -1 gpurun_out/n_ref_$c.log | cut -c1-300
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/n_ref_products.log 2>&1; tail -1 gpurun_out/n_ref_products.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sage_persist|sage_bwd_dw|adam_pack" -o gpurun_out/n_step -f python tools/profile_step.py > gpurun_out/n_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/n_launches.csv python tools/profile_step.py > gpurun_out/n_launches.log 2>&1
echo done
