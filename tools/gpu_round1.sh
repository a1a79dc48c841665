#!/bin/bash
# first GPU validation pass: kernel tests, tcgen05 diagnostics, trainer, benches
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not sage and not trainer" > gpurun_out/t1.log 2>&1; echo "t1 exit $?" >> gpurun_out/t1.log
timeout 300 python tools/diag_sage.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sage" > gpurun_out/t2.log 2>&1; echo "t2 exit $?" >> gpurun_out/t2.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "trainer" > gpurun_out/t3.log 2>&1; echo "t3 exit $?" >> gpurun_out/t3.log
timeout 600 python bench.py --small --steps 50 --warmup 5 > gpurun_out/bench_small.log 2>&1; echo "exit $?" >> gpurun_out/bench_small.log
timeout 1200 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
for f in t1 diag t2 t3 bench_small bench; do echo "=== $f"; tail -n 25 gpurun_out/$f.log; done
