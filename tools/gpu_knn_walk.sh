#!/bin/bash
# KNN tests + throughput (flat / IVF-flat, recall), optional: KNN_N=... for the database size
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gsl_engine_gpu.py -q -k knn > gpurun_out/t_knn_tests.log 2>&1; tail -4 gpurun_out/t_knn_tests.log
timeout 600 python tools/bench_knn.py > gpurun_out/t_knn1.log 2>&1; tail -2 gpurun_out/t_knn1.log | cut -c1-1000
