// Dynamic Graph Service kernels: the streaming TopK-by-timestamp sampler state lives in HBM tables
//   nbr / ts / w : [num_vertices, K]      (K most recent out-edges per vertex and edge type)
// Reference semantics: dynamic_graph_service/src/core/storage/topk_sampler.cc:23-41 (fixed-capacity sample, an
// incoming edge replaces the OLDEST kept one if it is newer) and src/core/execution/query_executor.cc:42-125
// (a query is a chain of prefix lookups).  The reference applies updates one record at a time on CPU actors and
// keeps the samples in RocksDB; here one kernel launch applies a whole record batch and one launch serves a hop
// of a whole batch of queries.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "host_utils.h"

namespace glb {

constexpr int kDgsMaxK = 64;

// The batch is sorted by (src, ts).  Only the last min(len, K) records of a source's segment can survive, so the
// thread sitting on the LAST record of every segment walks back over at most K records and inserts them into the
// vertex row (replace the oldest slot when newer) - O(K^2) per touched vertex, one launch per batch, no host
// round trip and no atomics (one thread owns a row).
__global__ void dgs_apply_edges_kernel(int64_t* __restrict__ nbr, int64_t* __restrict__ tst, float* __restrict__ wt,
                                       int64_t* __restrict__ count, int64_t n_vertices, int K, const int64_t* __restrict__ src,
                                       const int64_t* __restrict__ dst, const int64_t* __restrict__ ts,
                                       const float* __restrict__ w, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = src[i];
  if (i + 1 < n && src[i + 1] == v) return;            // not the end of a segment
  if (v < 0 || v >= n_vertices) return;                // hostile / unknown id: dropped
  int64_t* rn = nbr + v * K;
  int64_t* rt = tst + v * K;
  float* rw = wt + v * K;
  int64_t lo = i;
  for (int c = 1; c < K && lo > 0 && src[lo - 1] == v; ++c) --lo;
  for (int64_t e = lo; e <= i; ++e) {                  // ascending timestamps
    int slot = 0;
    int64_t oldest = rt[0];
    for (int s = 1; s < K; ++s) {
      const int64_t t = rt[s];
      if (t < oldest) { oldest = t; slot = s; }
    }
    if (ts[e] > oldest) { rn[slot] = dst[e]; rt[slot] = ts[e]; rw[slot] = w ? w[e] : 1.f; }
  }
  int c = 0;
  for (int s = 0; s < K; ++s) c += rn[s] >= 0;
  count[v] = c;
}

void dgs_apply_edges(const at::Tensor& nbr, const at::Tensor& ts_tab, const at::Tensor& w_tab, const at::Tensor& count,
                     const at::Tensor& src, const at::Tensor& dst, const at::Tensor& ts, const c10::optional<at::Tensor>& w) {
  TORCH_CHECK(nbr.is_cuda() && nbr.scalar_type() == at::kLong && nbr.dim() == 2 && nbr.is_contiguous());
  TORCH_CHECK(ts_tab.scalar_type() == at::kLong && ts_tab.is_contiguous() && w_tab.scalar_type() == at::kFloat && w_tab.is_contiguous());
  TORCH_CHECK(count.scalar_type() == at::kLong && count.numel() == nbr.size(0));
  check_cuda_i64(src, "src"); check_cuda_i64(dst, "dst"); check_cuda_i64(ts, "ts");
  TORCH_CHECK(src.is_contiguous() && dst.is_contiguous() && ts.is_contiguous() && dst.numel() == src.numel() && ts.numel() == src.numel());
  const int64_t n = src.numel();
  if (n == 0) return;
  c10::cuda::CUDAGuard guard(nbr.device());
  const float* wp = nullptr;
  at::Tensor wc;
  if (w.has_value() && w->defined()) { wc = w->contiguous(); TORCH_CHECK(wc.scalar_type() == at::kFloat && wc.numel() == n); wp = wc.data_ptr<float>(); }
  dgs_apply_edges_kernel<<<(unsigned)((n + 255) / 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      nbr.data_ptr<int64_t>(), ts_tab.data_ptr<int64_t>(), w_tab.data_ptr<float>(), count.data_ptr<int64_t>(), nbr.size(0),
      (int)nbr.size(1), src.data_ptr<int64_t>(), dst.data_ptr<int64_t>(), ts.data_ptr<int64_t>(), wp, n);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// One hop of a batch of queries: the k most recent kept samples of every query vertex, newest first, -1 padded.
// One thread per (vertex): k selection passes over the K slots (K <= 64, rows are 0.5 KB: L1/L2 resident).
__global__ void dgs_lookup_kernel(const int64_t* __restrict__ nbr, const int64_t* __restrict__ tst, const float* __restrict__ wt,
                                  int64_t n_vertices, int K, const int64_t* __restrict__ vids, int64_t B, int k,
                                  int64_t* __restrict__ out_n, int64_t* __restrict__ out_t, float* __restrict__ out_w) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t v = vids[b];
  int64_t* on = out_n + b * k;
  int64_t* ot = out_t + b * k;
  float* ow = out_w + b * k;
  if (v < 0 || v >= n_vertices) {
    for (int j = 0; j < k; ++j) { on[j] = -1; ot[j] = -(1LL << 62); ow[j] = 0.f; }
    return;
  }
  const int64_t* rn = nbr + v * K;
  const int64_t* rt = tst + v * K;
  const float* rw = wt + v * K;
  unsigned long long used = 0ull;
  for (int j = 0; j < k; ++j) {
    int best = -1;
    int64_t bt = -(1LL << 62);
    for (int s = 0; s < K; ++s) {
      if ((used >> s) & 1ull) continue;
      const int64_t t = rt[s];
      if (rn[s] >= 0 && (best < 0 || t > bt)) { best = s; bt = t; }
    }
    if (best < 0) { on[j] = -1; ot[j] = -(1LL << 62); ow[j] = 0.f; }
    else { used |= 1ull << best; on[j] = rn[best]; ot[j] = bt; ow[j] = rw[best]; }
  }
}

std::vector<at::Tensor> dgs_lookup(const at::Tensor& nbr, const at::Tensor& ts_tab, const at::Tensor& w_tab, const at::Tensor& vids,
                                   int64_t k) {
  TORCH_CHECK(nbr.is_cuda() && nbr.scalar_type() == at::kLong && nbr.dim() == 2 && nbr.is_contiguous() && nbr.size(1) <= kDgsMaxK,
              "sample tables hold at most ", kDgsMaxK, " slots per vertex");
  check_cuda_i64(vids, "vids");
  c10::cuda::CUDAGuard guard(nbr.device());
  auto v = vids.contiguous().view(-1);
  const int64_t B = v.numel();
  auto on = at::empty({B, k}, nbr.options());
  auto ot = at::empty({B, k}, nbr.options());
  auto ow = at::empty({B, k}, w_tab.options());
  if (B * k == 0) return {on, ot, ow};
  dgs_lookup_kernel<<<(unsigned)((B + 127) / 128), 128, 0, at::cuda::getCurrentCUDAStream()>>>(
      nbr.data_ptr<int64_t>(), ts_tab.data_ptr<int64_t>(), w_tab.data_ptr<float>(), nbr.size(0), (int)nbr.size(1), v.data_ptr<int64_t>(),
      B, (int)k, on.data_ptr<int64_t>(), ot.data_ptr<int64_t>(), ow.data_ptr<float>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {on, ot, ow};
}

}  // namespace glb
