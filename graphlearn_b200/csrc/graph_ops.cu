// K4: de-duplication / relabelling through a device hash table, and
// K6 (sparse flavour): fused gather x edge-weight -> scatter-add ("SpMM over an edge list") with
// vectorised float4 atomics, plus the per-edge dot product needed by its backward.
//
// Reference behaviour being replaced:
//   * subgraph_sampler.cc:35-95 / pyg_dataloader.py:56-64 / temporal_batch_loader.py:99-121 build
//     unique node sets and id -> local-index maps with std::unordered_map / numpy on the host;
//   * gcn_conv.py:48-77, sage_conv.py:61-92, gat_conv.py:69-119 run gather -> multiply -> segment_sum
//     as three TF ops.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "host_utils.h"

namespace glb {

namespace {
constexpr long long kEmpty = (long long)0x8000000000000000ULL;   // INT64_MIN never is a valid id

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// insert every valid id; remember the smallest input position that carried it
__global__ void __launch_bounds__(256)
relabel_insert_kernel(const int64_t* __restrict__ ids, int64_t n, long long* keys, long long* first,
                      uint64_t mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long id = ids[i];
  if (id < 0) return;
  uint64_t h = mix64((uint64_t)id) & mask;
  while (true) {
    long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(keys + h),
                                          (unsigned long long)kEmpty, (unsigned long long)id);
    if (prev == kEmpty || prev == id) { atomicMin(first + h, (long long)i); return; }
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ int64_t find_slot(const long long* __restrict__ keys, uint64_t mask, long long id) {
  uint64_t h = mix64((uint64_t)id) & mask;
  while (true) {
    long long k = keys[h];
    if (k == id) return (int64_t)h;
    if (k == kEmpty) return -1;
    h = (h + 1) & mask;
  }
}

// flag[i] = 1 when position i is the first occurrence of its id
__global__ void __launch_bounds__(256)
relabel_flag_kernel(const int64_t* __restrict__ ids, int64_t n, const long long* __restrict__ keys,
                    const long long* __restrict__ first, uint64_t mask, int64_t* __restrict__ slot_of,
                    int64_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long id = ids[i];
  int64_t s = id < 0 ? -1 : find_slot(keys, mask, id);
  slot_of[i] = s;
  flag[i] = (s >= 0 && first[s] == (long long)i) ? 1 : 0;
}

// rank_incl = inclusive prefix sum of flag: compact index of an id = rank_incl[first occurrence] - 1
__global__ void __launch_bounds__(256)
relabel_finish_kernel(const int64_t* __restrict__ ids, int64_t n, const int64_t* __restrict__ slot_of,
                      const long long* __restrict__ first, const int64_t* __restrict__ rank_incl,
                      int64_t* __restrict__ uniq, int64_t* __restrict__ inverse,
                      int64_t* __restrict__ slot_rank) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t s = slot_of[i];
  if (s < 0) { inverse[i] = -1; return; }
  long long f = first[s];
  int64_t r = rank_incl[f] - 1;
  inverse[i] = r;
  if (f == (long long)i) { uniq[r] = ids[i]; slot_rank[s] = r; }
}

__global__ void __launch_bounds__(256)
relabel_lookup_kernel(const int64_t* __restrict__ q, int64_t n, const long long* __restrict__ keys,
                      const int64_t* __restrict__ slot_rank, uint64_t mask, int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long id = q[i];
  int64_t s = id < 0 ? -1 : find_slot(keys, mask, id);
  out[i] = s < 0 ? -1 : slot_rank[s];
}

// ------------------------------------------------------------------ edge-list SpMM
// out[row[e], h, :] += w[e, h] * x[col[e], h, :]     one warp per edge, lanes over the H*D features
template <bool VEC4>
__global__ void __launch_bounds__(256)
edge_scatter_kernel(const float* __restrict__ x, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                    const float* __restrict__ w, int64_t E, int H, int D, int64_t n_src, int64_t n_out,
                    float* __restrict__ out) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  const int64_t r = row[e], c = col[e];
  if (r < 0 || r >= n_out || c < 0 || c >= n_src) return;
  const int F = H * D;
  const float* xs = x + c * (int64_t)F;
  float* od = out + r * (int64_t)F;
  if constexpr (VEC4) {
    for (int f = 4 * lane; f < F; f += 128) {
      float4 v = *reinterpret_cast<const float4*>(xs + f);
      const float ww = w ? w[e * H + f / D] : 1.f;           // D % 4 == 0: the 4 features share a head
      v.x *= ww; v.y *= ww; v.z *= ww; v.w *= ww;
      atomicAdd(reinterpret_cast<float4*>(od + f), v);        // red.global.add.v4.f32 (sm_90+)
    }
  } else {
    for (int f = lane; f < F; f += 32) {
      const float ww = w ? w[e * H + f / D] : 1.f;
      atomicAdd(od + f, ww * xs[f]);
    }
  }
}

// out[e, h] = < a[row[e], h, :], b[col[e], h, :] >          one warp per edge
__global__ void __launch_bounds__(256)
edge_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, const int64_t* __restrict__ row,
                const int64_t* __restrict__ col, int64_t E, int H, int D, int64_t na, int64_t nb,
                float* __restrict__ out) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  const int64_t r = row[e], c = col[e];
  const bool ok = r >= 0 && r < na && c >= 0 && c < nb;
  const int F = H * D;
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    if (ok)
      for (int d = lane; d < D; d += 32) s += a[r * (int64_t)F + h * D + d] * b[c * (int64_t)F + h * D + d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[e * H + h] = s;
  }
}

inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }
}  // namespace

// returns (uniq [U] in first-occurrence order, inverse [n] (-1 for negative ids), table_keys, table_rank)
// sync_free: `uniq` is allocated with the upper bound n (entries past the real count are -1) and the count comes back as
// a device tensor - no value is read on the host, so the op can sit inside a captured graph.
std::vector<at::Tensor> relabel(const at::Tensor& ids_in, bool sync_free) {
  check_cuda_i64(ids_in, "ids");
  c10::cuda::CUDAGuard guard(ids_in.device());
  auto ids = ids_in.contiguous().view({-1});
  const int64_t n = ids.numel();
  int64_t cap = 64;
  while (cap < 2 * n) cap <<= 1;
  auto opts = ids.options();
  auto keys = at::full({cap}, (int64_t)kEmpty, opts);
  auto first = at::full({cap}, std::numeric_limits<int64_t>::max(), opts);
  auto slot_rank = at::full({cap}, (int64_t)-1, opts);
  auto inverse = at::empty({n}, opts);
  if (n == 0) return {at::empty({0}, opts), inverse, keys, slot_rank, at::zeros({1}, opts)};
  auto slot_of = at::empty({n}, opts);
  auto flag = at::empty({n}, opts);
  auto stream = at::cuda::getCurrentCUDAStream();
  const uint64_t mask = (uint64_t)cap - 1;
  auto kp = reinterpret_cast<long long*>(keys.data_ptr<int64_t>());
  auto fp = reinterpret_cast<long long*>(first.data_ptr<int64_t>());
  relabel_insert_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(ids.data_ptr<int64_t>(), n, kp, fp, mask);
  relabel_flag_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(ids.data_ptr<int64_t>(), n, kp, fp, mask,
                                                              slot_of.data_ptr<int64_t>(), flag.data_ptr<int64_t>());
  auto rank = at::cumsum(flag, 0);
  const int64_t U = sync_free ? n : rank[n - 1].item<int64_t>();    // eager API: exact, data dependent size
  auto uniq = sync_free ? at::full({U}, (int64_t)-1, opts) : at::empty({U}, opts);
  relabel_finish_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(
      ids.data_ptr<int64_t>(), n, slot_of.data_ptr<int64_t>(), fp, rank.data_ptr<int64_t>(),
      uniq.data_ptr<int64_t>(), inverse.data_ptr<int64_t>(), slot_rank.data_ptr<int64_t>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {uniq, inverse, keys, slot_rank, rank.slice(0, n - 1, n)};
}

// compact index of every query id in a table built by relabel(); -1 when absent
at::Tensor relabel_lookup(const at::Tensor& keys, const at::Tensor& slot_rank, const at::Tensor& queries) {
  check_cuda_i64(queries, "queries");
  check_cuda_i64(keys, "keys");
  c10::cuda::CUDAGuard guard(queries.device());
  auto q = queries.contiguous().view({-1});
  auto out = at::empty({q.numel()}, q.options());
  if (q.numel() == 0) return out;
  relabel_lookup_kernel<<<blocks_for(q.numel(), 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      q.data_ptr<int64_t>(), q.numel(), reinterpret_cast<const long long*>(keys.data_ptr<int64_t>()),
      slot_rank.data_ptr<int64_t>(), (uint64_t)keys.numel() - 1, out.data_ptr<int64_t>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

// out [n_out, H*D] (fp32, zero-initialised here) ;  x [n_src, H*D] fp32 ; w [E, H] fp32 or none
at::Tensor edge_scatter(const at::Tensor& x, const at::Tensor& row, const at::Tensor& col,
                        const c10::optional<at::Tensor>& w, int64_t H, int64_t n_out) {
  check_cuda_i64(row, "row");
  check_cuda_i64(col, "col");
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kFloat && x.dim() == 2, "x must be a CUDA fp32 matrix");
  c10::cuda::CUDAGuard guard(x.device());
  auto xc = x.contiguous();
  auto rc = row.contiguous(), cc = col.contiguous();
  const int64_t E = rc.numel(), F = xc.size(1);
  TORCH_CHECK(cc.numel() == E && H >= 1 && F % H == 0, "bad edge list / head count");
  const int D = (int)(F / H);
  at::Tensor wc;
  const float* wp = nullptr;
  if (w.has_value()) {
    wc = w->contiguous();
    TORCH_CHECK(wc.is_cuda() && wc.scalar_type() == at::kFloat && wc.numel() == E * H, "w must be fp32 [E, H]");
    wp = wc.data_ptr<float>();
  }
  auto out = at::zeros({n_out, F}, xc.options());
  if (E == 0 || F == 0) return out;
  auto stream = at::cuda::getCurrentCUDAStream();
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(xc.data_ptr()) & 15) == 0);
  if (vec)
    edge_scatter_kernel<true><<<blocks_for(E * 32, 256), 256, 0, stream>>>(
        xc.data_ptr<float>(), rc.data_ptr<int64_t>(), cc.data_ptr<int64_t>(), wp, E, (int)H, D, xc.size(0), n_out,
        out.data_ptr<float>());
  else
    edge_scatter_kernel<false><<<blocks_for(E * 32, 256), 256, 0, stream>>>(
        xc.data_ptr<float>(), rc.data_ptr<int64_t>(), cc.data_ptr<int64_t>(), wp, E, (int)H, D, xc.size(0), n_out,
        out.data_ptr<float>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

at::Tensor edge_dot(const at::Tensor& a, const at::Tensor& b, const at::Tensor& row, const at::Tensor& col,
                    int64_t H) {
  check_cuda_i64(row, "row");
  check_cuda_i64(col, "col");
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.scalar_type() == at::kFloat && b.scalar_type() == at::kFloat);
  c10::cuda::CUDAGuard guard(a.device());
  auto ac = a.contiguous(), bc = b.contiguous();
  auto rc = row.contiguous(), cc = col.contiguous();
  const int64_t E = rc.numel(), F = ac.size(1);
  TORCH_CHECK(bc.size(1) == F && F % H == 0);
  auto out = at::zeros({E, H}, ac.options());
  if (E == 0) return out;
  edge_dot_kernel<<<blocks_for(E * 32, 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      ac.data_ptr<float>(), bc.data_ptr<float>(), rc.data_ptr<int64_t>(), cc.data_ptr<int64_t>(), E, (int)H,
      (int)(F / H), ac.size(0), bc.size(0), out.data_ptr<float>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

}  // namespace glb
