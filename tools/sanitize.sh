#!/bin/bash
# Race / memory checking of the hand-written kernels (the reference has no sanitizer jobs at all).
# Run on a GPU box:   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'   (or racecheck / synccheck / initcheck)
# The kernel tests are small enough for compute-sanitizer's ~50x slowdown; the CUDA-graph engine tests are
# excluded (graph capture is not supported under the sanitizer) - their kernels are covered eagerly by
# test_fast_engine_matches_autograd.
tool=${1:-memcheck}
limit=${2:-700}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout $limit compute-sanitizer --tool "$tool" --error-exitcode 3 --log-file gpurun_out/sanitizer_$tool.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gsl_engine_gpu.py -m gpu -q \
  -k "(exact_rows and 300) or (exact_rows and 77) or multi_segment or bwd_dw or matches_autograd or fp8_block or timestamp_filter or in_degree_strategy or node2vec or sample_topk or sample_full or gather_rows or relabel or negative_sampler or random_walk or sparse_adam or gat_fused or knn or dgs_kernels or edge_spmm" \
  > gpurun_out/sanitizer_$tool.pytest.log 2>&1
echo "exit $?"
tail -5 gpurun_out/sanitizer_$tool.log
tail -3 gpurun_out/sanitizer_$tool.pytest.log
