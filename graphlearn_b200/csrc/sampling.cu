// K1: neighbor samplers over a peer-mapped, hash-partitioned CSR.
//
// One thread per OUTPUT SLOT (b, j): the B*k slots give tens of thousands of
// independent 8-byte reads in flight, which is what hides the ~2 us NVLink
// peer-load latency (B300_MICROARCH: peer-LDG ~1.8-2.0k cycles).  The k
// threads of one seed read the same indptr pair, which the LSU coalesces to
// one request.
//
// Semantics follow the reference operators (behaviour, not code):
//   random                      graphlearn/src/core/operator/sampler/random_sampler.cc:53-74
//   random_without_replacement  .../random_without_replacement_sampler.cc:59-68
//   topk                        .../topk_sampler.cc:53-61
//   edge_weight / in_degree     .../edge_weight_sampler.cc:78-92, in_degree_sampler.cc:77-89
//   full                        .../full_sampler.cc:43-88
//   padding                     .../padder/circular_padder.h:36-66, replicate_padder.h:37-56
//   filter                      .../filter.cc:68-93,154-228
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "csr_view.cuh"
#include "host_utils.h"

namespace glb {

enum Strategy : int {
  kRandom = 0,
  kRandomWithoutReplacement = 1,
  kTopk = 2,
  kEdgeWeight = 3,   // also used for in_degree (different cumw array)
};

enum FilterMode : int {
  kNoFilter = 0,
  kFilterIdEqual = 1,        // drop neighbours whose id == filter[b]
  kFilterTsLargerThan = 2,   // keep edges whose timestamp < filter[b] (rows sorted by ts asc)
};

struct SampleParams {
  const int64_t* src;        // [B] vids
  const int64_t* filter;     // [B] or null
  int64_t* out_nbr;          // [B, k]
  int64_t* out_eid;          // [B, k] or null
  const uint64_t* rng;       // {seed, offset}
  int64_t B;
  int k;
  int strategy;
  int filter_mode;
  int padding_circular;      // 1 = circular, 0 = replicate default id
  int retry;                 // SamplingRetryTimes
  int64_t default_id;
  uint32_t salt;
};

// number of leading edges of the (ts-ascending) row with ts < bound: the reference's accelerated
// timestamp filter keeps exactly the edges BEFORE the bound (filter.cc:68-82,193-228), so an event
// never sees itself or a simultaneous edge
__device__ __forceinline__ int64_t ts_prefix(const RowRef& r, int64_t bound) {
  int64_t lo = 0, hi = r.deg;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(r.ts + r.beg + mid) < bound) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// weighted pick in [0, m): smallest i with cumw[i] > u * cumw[m-1]
__device__ __forceinline__ int64_t weighted_pick(const RowRef& r, int64_t m, float u) {
  float total = __ldg(r.cumw + r.beg + m - 1);
  float target = u * total;
  int64_t lo = 0, hi = m - 1;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(r.cumw + r.beg + mid) > target) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
sample_neighbors_kernel(const CsrView g, const SampleParams p) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = p.B * p.k;
  if (t >= total) return;
  int64_t b = t / p.k;
  int j = (int)(t - b * p.k);

  int64_t out_id = p.default_id, out_e = -1;
  RowRef r = csr_row(g, __ldg(p.src + b));
  int64_t m = r.deg;                        // effective row length
  bool reversed = false;
  int64_t fval = 0;
  bool idf = false;
  if (p.filter_mode == kFilterTsLargerThan && r.deg > 0 && r.ts != nullptr) {
    m = ts_prefix(r, __ldg(p.filter + b));
    reversed = true;                        // most recent first (reference returns descending)
  } else if (p.filter_mode == kFilterIdEqual) {
    fval = __ldg(p.filter + b);
    idf = true;
  }

  if (m > 0) {
    uint4 rnd = rng4(p.rng, p.salt, (uint64_t)t);
    int64_t pick = -1;
    if (p.strategy == kRandom || p.strategy == kEdgeWeight) {
      // draws with replacement; re-draw on an id-filter hit up to `retry`
      // times, then fall back to a bounded circular scan for a non-hit.
      uint32_t r0 = rnd.x, r1 = rnd.y;
      int tries = 0;
      while (true) {
        int64_t idx;
        if (p.strategy == kRandom) idx = (int64_t)bounded64(r0, r1, (uint64_t)m);
        else idx = weighted_pick(r, m, u01(r0));
        if (!idf || __ldg(r.indices + r.beg + idx) != fval) { pick = idx; break; }
        if (++tries > p.retry) {
          int64_t lim = m < 1024 ? m : 1024;
          for (int64_t s = 1; s <= lim; ++s) {
            int64_t c = idx + s; if (c >= m) c -= m;
            if (__ldg(r.indices + r.beg + c) != fval) { pick = c; break; }
          }
          break;                            // pick stays -1 when every neighbour is filtered
        }
        uint4 nr = rng4(p.rng, p.salt ^ (0x5bd1e995u * (uint32_t)tries), (uint64_t)t);
        r0 = nr.x; r1 = nr.y;
      }
    } else {
      // ordered strategies: order(i) = perm(i) | i, then padding
      uint32_t k0 = 0, k1 = 0;
      const bool shuffled = (p.strategy == kRandomWithoutReplacement);
      if (shuffled) {
        // one permutation per (step, seed row): key from the row, not the slot
        uint4 rk = rng4(p.rng, p.salt ^ 0xA511E9B3u, (uint64_t)b);
        k0 = rk.x; k1 = rk.y;
      }
      uint32_t m32 = (uint32_t)(m > 0x7fffffff ? 0x7fffffff : m);
      auto order = [&](int64_t i) -> int64_t {
        int64_t o = shuffled ? (int64_t)feistel_perm((uint32_t)i, m32, k0, k1) : i;
        return reversed ? (m - 1 - o) : o;
      };
      if (!idf) {
        if (j < m) pick = order(j);
        else if (p.padding_circular) pick = order(j % m);
      } else {
        // j-th non-hit element of the ordered row
        int64_t cnt = 0, found = -1, scan = m < 65536 ? m : 65536;
        for (int64_t i = 0; i < scan; ++i) {
          int64_t o = order(i);
          if (__ldg(r.indices + r.beg + o) != fval) {
            if (cnt == j) { found = o; break; }
            ++cnt;
          }
        }
        if (found >= 0) pick = found;
        else if (cnt > 0 && p.padding_circular) {
          int64_t want = j % cnt, c2 = 0;
          for (int64_t i = 0; i < scan; ++i) {
            int64_t o = order(i);
            if (__ldg(r.indices + r.beg + o) != fval) {
              if (c2 == want) { pick = o; break; }
              ++c2;
            }
          }
        }
      }
    }
    if (pick >= 0) {
      out_id = __ldg(r.indices + r.beg + pick);
      out_e = r.eids ? __ldg(r.eids + r.beg + pick) : (r.beg + pick);
    }
  }
  p.out_nbr[t] = out_id;
  if (p.out_eid) p.out_eid[t] = out_e;
}

// ---------------------------------------------------------------------------
// degrees and the two-pass "full" sampler
// ---------------------------------------------------------------------------
__global__ void degree_kernel(const CsrView g, const int64_t* __restrict__ src, int64_t B,
                              int64_t cap, int64_t* __restrict__ out) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  RowRef r = csr_row(g, __ldg(src + b));
  int64_t d = r.deg;
  if (cap > 0 && d > cap) d = cap;
  out[b] = d;
}

// one warp per seed copies the (truncated) row into values[offsets[b] ...]
__global__ void __launch_bounds__(256)
full_scatter_kernel(const CsrView g, const int64_t* __restrict__ src, int64_t B, int64_t cap,
                    const int64_t* __restrict__ offsets, int64_t* __restrict__ out_nbr,
                    int64_t* __restrict__ out_eid) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= B) return;
  RowRef r = csr_row(g, __ldg(src + w));
  int64_t d = r.deg;
  if (cap > 0 && d > cap) d = cap;
  int64_t o = __ldg(offsets + w);
  const int64_t room = __ldg(offsets + w + 1) - o;        // sync-free mode clamps the offsets to the output capacity
  if (d > room) d = room;
  for (int64_t i = lane; i < d; i += 32) {
    out_nbr[o + i] = __ldg(r.indices + r.beg + i);
    if (out_eid) out_eid[o + i] = r.eids ? __ldg(r.eids + r.beg + i) : (r.beg + i);
  }
}

// ---------------------------------------------------------------------------
// RNG state maintenance (1 thread): bump the step offset inside a CUDA graph
// ---------------------------------------------------------------------------
__global__ void rng_advance_kernel(uint64_t* state, uint64_t inc) { state[1] += inc; }

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
CsrView csr_from_desc(const at::Tensor& desc) {
  TORCH_CHECK(desc.device().is_cpu() && desc.scalar_type() == at::kLong &&
                  (desc.numel() == 1 + 6 * kMaxWorld || desc.numel() == 1 + 7 * kMaxWorld),
              "csr desc must be a CPU int64 tensor of 49 (or 57: + sorted rows) entries");
  const bool has_sorted = desc.numel() == 1 + 7 * kMaxWorld;
  const int64_t* d = desc.data_ptr<int64_t>();
  CsrView g;
  g.world = (int)d[0];
  TORCH_CHECK(g.world >= 1 && g.world <= kMaxWorld, "bad world size in csr desc");
  for (int r = 0; r < kMaxWorld; ++r) {
    g.nrows[r] = d[1 + r];
    g.indptr.p[r] = reinterpret_cast<const void*>(d[1 + kMaxWorld + r]);
    g.indices.p[r] = reinterpret_cast<const void*>(d[1 + 2 * kMaxWorld + r]);
    g.eids.p[r] = reinterpret_cast<const void*>(d[1 + 3 * kMaxWorld + r]);
    g.cumw.p[r] = reinterpret_cast<const void*>(d[1 + 4 * kMaxWorld + r]);
    g.ts.p[r] = reinterpret_cast<const void*>(d[1 + 5 * kMaxWorld + r]);
    g.sorted.p[r] = has_sorted ? reinterpret_cast<const void*>(d[1 + 6 * kMaxWorld + r]) : nullptr;
  }
  return g;
}

std::vector<at::Tensor> sample_neighbors(const at::Tensor& csr_desc, const at::Tensor& src,
                                         int64_t k, int64_t strategy, int64_t filter_mode,
                                         const c10::optional<at::Tensor>& filter,
                                         bool padding_circular, int64_t retry, int64_t default_id,
                                         const at::Tensor& rng_state, int64_t salt, bool want_eids,
                                         const c10::optional<at::Tensor>& out_buf) {
  check_cuda_i64(src, "src");
  TORCH_CHECK(rng_state.is_cuda() && rng_state.scalar_type() == at::kLong && rng_state.numel() >= 2,
              "rng_state must be a CUDA int64[2] tensor");
  c10::cuda::CUDAGuard guard(src.device());
  CsrView g = csr_from_desc(csr_desc);
  int64_t B = src.numel();
  at::Tensor nbr;
  if (out_buf.has_value() && out_buf->defined()) {
    check_cuda_i64(*out_buf, "out");
    TORCH_CHECK(out_buf->is_contiguous() && out_buf->numel() == B * k, "out must be a contiguous int64 buffer of B*k elements");
    nbr = out_buf->view({B, k});
  } else {
    nbr = at::empty({B, k}, src.options());
  }
  at::Tensor eid;
  if (want_eids) eid = at::empty({B, k}, src.options());
  if (B * k == 0) return {nbr, want_eids ? eid : at::Tensor()};
  SampleParams p;
  auto srcc = src.contiguous();
  p.src = srcc.data_ptr<int64_t>();
  at::Tensor fc;
  p.filter = nullptr;
  if (filter_mode != kNoFilter) {
    TORCH_CHECK(filter.has_value(), "filter values required");
    fc = filter->contiguous();
    check_cuda_i64(fc, "filter");
    TORCH_CHECK(fc.numel() == B, "filter must have one value per source id");
    p.filter = fc.data_ptr<int64_t>();
  }
  p.out_nbr = nbr.data_ptr<int64_t>();
  p.out_eid = want_eids ? eid.data_ptr<int64_t>() : nullptr;
  p.rng = reinterpret_cast<const uint64_t*>(rng_state.data_ptr<int64_t>());
  p.B = B; p.k = (int)k; p.strategy = (int)strategy; p.filter_mode = (int)filter_mode;
  p.padding_circular = padding_circular ? 1 : 0; p.retry = (int)retry;
  p.default_id = default_id; p.salt = (uint32_t)salt;
  if (strategy == kEdgeWeight) {
    TORCH_CHECK(g.cumw.p[0] != nullptr, "weighted sampling needs a prefix-sum array");
  }
  int threads = 256;
  int64_t blocks = (B * k + threads - 1) / threads;
  sample_neighbors_kernel<<<(unsigned)blocks, threads, 0, at::cuda::getCurrentCUDAStream()>>>(g, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {nbr, want_eids ? eid : at::Tensor()};
}

at::Tensor get_degrees(const at::Tensor& csr_desc, const at::Tensor& src, int64_t cap) {
  check_cuda_i64(src, "src");
  c10::cuda::CUDAGuard guard(src.device());
  CsrView g = csr_from_desc(csr_desc);
  auto srcc = src.contiguous();
  auto out = at::empty({src.numel()}, src.options());
  int64_t B = src.numel();
  if (B == 0) return out;
  degree_kernel<<<(unsigned)((B + 255) / 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      g, srcc.data_ptr<int64_t>(), B, cap, out.data_ptr<int64_t>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

// returns (values, eids, offsets[B+1]); sparse output like FullSampler.
// `max_total` > 0: sync-free mode for static-shape / graph-captured callers - the outputs are allocated with that
// capacity (rows are truncated when the data needs more; offsets[B] tells how much was really produced) and no
// device value is read back.  `max_total` <= 0: the exact size is read from the device (one host sync).
std::vector<at::Tensor> sample_full(const at::Tensor& csr_desc, const at::Tensor& src, int64_t cap,
                                    bool want_eids, int64_t max_total) {
  check_cuda_i64(src, "src");
  c10::cuda::CUDAGuard guard(src.device());
  CsrView g = csr_from_desc(csr_desc);
  auto srcc = src.contiguous();
  int64_t B = src.numel();
  auto deg = get_degrees(csr_desc, srcc, cap);
  auto offsets = at::zeros({B + 1}, src.options());
  if (B > 0) offsets.slice(0, 1).copy_(at::cumsum(deg, 0));
  int64_t total;
  if (max_total > 0) {
    total = max_total;                                        // capacity given by the caller: no host sync
    offsets.clamp_max_(max_total);                            // rows past the capacity become empty / truncated
  } else {
    total = B > 0 ? offsets[B].item<int64_t>() : 0;           // eager API: the exact size is data dependent
  }
  auto vals = max_total > 0 ? at::full({total}, (int64_t)-1, src.options()) : at::empty({total}, src.options());
  at::Tensor eids;
  if (want_eids) eids = at::empty({total}, src.options());
  if (total > 0) {
    int64_t blocks = (B * 32 + 255) / 256;
    full_scatter_kernel<<<(unsigned)blocks, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
        g, srcc.data_ptr<int64_t>(), B, cap, offsets.data_ptr<int64_t>(), vals.data_ptr<int64_t>(),
        want_eids ? eids.data_ptr<int64_t>() : nullptr);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  return {vals, want_eids ? eids : at::Tensor(), offsets};
}

void rng_advance(const at::Tensor& rng_state, int64_t inc) {
  TORCH_CHECK(rng_state.is_cuda() && rng_state.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard guard(rng_state.device());
  rng_advance_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>(
      reinterpret_cast<uint64_t*>(rng_state.data_ptr<int64_t>()), (uint64_t)inc);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace glb
