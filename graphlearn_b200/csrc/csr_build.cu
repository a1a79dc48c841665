// K11: CSR build on the device without global sorts.
//
// The portable build (store/shards.py CsrShard.from_coo) orders the E edges of a shard with two stable global sorts
// (by timestamp / weight, then by source row).  This file builds the same permutation with a counting pass instead:
//
//   1. csr_count_kernel      counts[row]++                                   (one atomic per edge)
//   2. indptr = exclusive prefix sum of counts                               (at::cumsum)
//   3. csr_fill_kernel       slot = indptr[row] + cursor[row]++ ; order[slot] = edge, key[slot] = sort key of the edge
//   4. csr_row_sort_kernel   every row segment is sorted by (key, edge index) - a warp per row in shared memory for rows
//                            of up to 256 edges, csr_long_row_sort_kernel (a CTA per row, in place in global memory) for
//                            the hubs.  Both run the same normalised bitonic network (all compare-exchanges point the same
//                            way, so rows need not be padded to a power of two).
//
// The result is EXACTLY the permutation of the stable sorts: (row asc, key asc, insertion order asc), so edge ids
// (= CSR positions) do not depend on which path built the shard.  Keys: none, timestamp ascending, or weight descending
// (the reference keeps rows weight-ordered for the top-k sampler: graphlearn/src/core/graph/storage/
// memory_topo_storage.cc / topk_sampler.cc; load path graph_store.cc:60-165).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAException.h>
#include "host_utils.h"

namespace glb {

namespace {

constexpr int kSortWarps = 8;
constexpr int kWarpRowMax = 256;        // edges of a row sorted by one warp in shared memory (8 warps x 256 x 16 B = 32 KB static)

__device__ __forceinline__ unsigned long long ts_key(long long ts) { return (unsigned long long)ts ^ 0x8000000000000000ull; }

// weight DESCENDING as an ascending unsigned key (-0.0 and 0.0 compare equal, like a float comparison)
__device__ __forceinline__ unsigned long long weight_desc_key(float w) {
  unsigned b = __float_as_uint(w);
  if ((b << 1) == 0u) b = 0u;                                 // -0.0 -> +0.0
  const unsigned asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return (unsigned long long)(~asc);
}

__global__ void csr_count_kernel(const int64_t* __restrict__ rows, int64_t E, int64_t n_rows, unsigned long long* __restrict__ counts,
                                 int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[i];
    if (r < 0 || r >= n_rows) { *bad = 1; continue; }
    atomicAdd(counts + r, 1ull);
  }
}

// key_mode: 0 none, 1 int64 timestamps ascending, 2 fp32 weights descending
__global__ void csr_fill_kernel(const int64_t* __restrict__ rows, int64_t E, int64_t n_rows, const int64_t* __restrict__ indptr,
                                unsigned long long* __restrict__ cursor, int key_mode, const void* __restrict__ key_in,
                                int64_t* __restrict__ order, unsigned long long* __restrict__ keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[i];
    if (r < 0 || r >= n_rows) continue;
    const int64_t slot = indptr[r] + (int64_t)atomicAdd(cursor + r, 1ull);
    order[slot] = i;
    if (key_mode == 1) keys[slot] = ts_key(reinterpret_cast<const long long*>(key_in)[i]);
    else if (key_mode == 2) keys[slot] = weight_desc_key(reinterpret_cast<const float*>(key_in)[i]);
  }
}

// compare-exchange of positions i < j: afterwards (k[i], v[i]) <= (k[j], v[j]) lexicographically (k == nullptr: v only)
__device__ __forceinline__ void cmpx(unsigned long long* k, long long* v, int i, int j) {
  const long long vi = v[i], vj = v[j];
  bool swap;
  if (k != nullptr) {
    const unsigned long long ki = k[i], kj = k[j];
    swap = ki > kj || (ki == kj && vi > vj);
    if (swap) { k[i] = kj; k[j] = ki; }
  } else {
    swap = vi > vj;
  }
  if (swap) { v[i] = vj; v[j] = vi; }
}

// normalised bitonic sort of n pairs by `nthreads` cooperating threads (a warp or a CTA)
template <bool kCta>
__device__ __forceinline__ void bitonic_pairs(unsigned long long* k, long long* v, int n, int tid, int nthreads) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const int half = np2 >> 1;
  for (int size = 2; size <= np2; size <<= 1) {
    const int hs = size >> 1;
    for (int t = tid; t < half; t += nthreads) {            // mirrored first step of the merge
      const int blk = t / hs, w = t - blk * hs;
      const int i = blk * size + w, j = blk * size + size - 1 - w;
      if (j < n) cmpx(k, v, i, j);
    }
    if (kCta) __syncthreads(); else __syncwarp();
    for (int stride = hs >> 1; stride >= 1; stride >>= 1) {
      for (int t = tid; t < half; t += nthreads) {
        const int i = (t / stride) * (stride << 1) + (t % stride), j = i + stride;
        if (j < n) cmpx(k, v, i, j);
      }
      if (kCta) __syncthreads(); else __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(kSortWarps * 32) csr_row_sort_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, bool has_keys,
                                                                        int64_t* __restrict__ order, unsigned long long* __restrict__ keys) {
  __shared__ unsigned long long sk[kSortWarps][kWarpRowMax];
  __shared__ long long sv[kSortWarps][kWarpRowMax];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * kSortWarps + warp; r < n_rows; r += (int64_t)gridDim.x * kSortWarps) {
    const int64_t lo = indptr[r], hi = indptr[r + 1];
    const int64_t len = hi - lo;
    if (len < 2 || len > kWarpRowMax) continue;              // warp-uniform
    const int n = (int)len;
    for (int e = lane; e < n; e += 32) {
      sv[warp][e] = order[lo + e];
      if (has_keys) sk[warp][e] = keys[lo + e];
    }
    __syncwarp();
    bitonic_pairs<false>(has_keys ? sk[warp] : nullptr, sv[warp], n, lane, 32);
    for (int e = lane; e < n; e += 32) order[lo + e] = sv[warp][e];
    __syncwarp();
  }
}

// hubs: one CTA per listed row, in place in global memory
__global__ void __launch_bounds__(1024) csr_long_row_sort_kernel(const int64_t* __restrict__ indptr, const int64_t* __restrict__ long_rows,
                                                                 bool has_keys, int64_t* __restrict__ order,
                                                                 unsigned long long* __restrict__ keys) {
  const int64_t r = long_rows[blockIdx.x];
  const int64_t lo = indptr[r], hi = indptr[r + 1];
  bitonic_pairs<true>(has_keys ? keys + lo : nullptr, reinterpret_cast<long long*>(order + lo), (int)(hi - lo), threadIdx.x, blockDim.x);
}

}  // namespace

// (indptr int64 [n_rows + 1], order int64 [E]): order lists the edges grouped by source row; inside a row by key
// (key_mode 1: int64 `key` ascending, 2: fp32 `key` descending, 0: none) and then by edge index.
std::vector<at::Tensor> csr_build(const at::Tensor& src_rows, int64_t n_rows, const c10::optional<at::Tensor>& key, int64_t key_mode) {
  check_cuda_i64(src_rows, "src_rows");
  TORCH_CHECK(src_rows.dim() == 1 && src_rows.is_contiguous() && n_rows >= 0);
  TORCH_CHECK(key_mode >= 0 && key_mode <= 2);
  c10::cuda::CUDAGuard guard(src_rows.device());
  const int64_t E = src_rows.numel();
  at::Tensor kin;
  if (key_mode != 0) {
    TORCH_CHECK(key.has_value() && key->defined() && key->is_cuda() && key->numel() == E, "key must be a CUDA tensor with one entry per edge");
    kin = key->contiguous();
    TORCH_CHECK(kin.scalar_type() == (key_mode == 1 ? at::kLong : at::kFloat), "timestamps must be int64, weights fp32");
  }
  auto ol = src_rows.options().dtype(at::kLong);
  auto counts = at::zeros({n_rows + 1}, ol);                 // [0] stays 0: cumsum(counts shifted by one) = indptr
  auto order = at::empty({E}, ol);
  if (E == 0) return {at::zeros({n_rows + 1}, ol), order};
  auto bad = at::zeros({1}, src_rows.options().dtype(at::kInt));
  auto stream = at::cuda::getCurrentCUDAStream();
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((E + threads - 1) / threads, (int64_t)sm_count() * 16);
  csr_count_kernel<<<blocks, threads, 0, stream>>>(src_rows.data_ptr<int64_t>(), E, n_rows,
                                                   reinterpret_cast<unsigned long long*>(counts.data_ptr<int64_t>()) + 1, bad.data_ptr<int>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  auto indptr = at::cumsum(counts, 0);
  auto cursor = at::zeros({std::max<int64_t>(n_rows, 1)}, ol);
  at::Tensor keys;
  if (key_mode != 0) keys = at::empty({E}, ol);
  unsigned long long* kp = key_mode != 0 ? reinterpret_cast<unsigned long long*>(keys.data_ptr<int64_t>()) : nullptr;
  csr_fill_kernel<<<blocks, threads, 0, stream>>>(src_rows.data_ptr<int64_t>(), E, n_rows, indptr.data_ptr<int64_t>(),
                                                  reinterpret_cast<unsigned long long*>(cursor.data_ptr<int64_t>()), (int)key_mode,
                                                  key_mode != 0 ? kin.data_ptr() : nullptr, order.data_ptr<int64_t>(), kp);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  if (n_rows > 0) {
    const int sblocks = (int)std::min<int64_t>((n_rows + kSortWarps - 1) / kSortWarps, (int64_t)sm_count() * 8);
    csr_row_sort_kernel<<<sblocks, kSortWarps * 32, 0, stream>>>(indptr.data_ptr<int64_t>(), n_rows, key_mode != 0, order.data_ptr<int64_t>(), kp);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  // hub rows (one size read-back; a build-time path)
  auto deg = indptr.slice(0, 1, n_rows + 1) - indptr.slice(0, 0, n_rows);
  auto long_rows = at::nonzero(deg > kWarpRowMax).flatten().contiguous();
  TORCH_CHECK(bad.item<int>() == 0, "csr_build: a source row is outside [0, n_rows)");
  if (long_rows.numel() > 0) {
    TORCH_CHECK(deg.max().item<int64_t>() < (int64_t)1 << 30, "a single adjacency row with 2^30 edges or more");
    csr_long_row_sort_kernel<<<(unsigned)long_rows.numel(), 1024, 0, stream>>>(indptr.data_ptr<int64_t>(), long_rows.data_ptr<int64_t>(),
                                                                                key_mode != 0, order.data_ptr<int64_t>(), kp);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  return {indptr, order};
}

}  // namespace glb
