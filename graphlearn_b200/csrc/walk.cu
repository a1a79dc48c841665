// K3: random-walk engine (DeepWalk + node2vec p/q) over the peer-mapped CSR.
//
// The reference performs one distributed Partition->RPC->Stitch round PER STEP
// (graphlearn/src/core/operator/random_walk/random_walk.cc:55-135; a length-40
// walk = 39 sequential all-to-all rounds) and, for node2vec, ships the parent's
// neighbour list (<= DefaultFullNbrNum ids) to the next shard every step
// (random_walk.cc:99-135,188-270).  Here a walker is one thread that stays
// resident for the whole walk: every hop dereferences the adjacency row of the
// current vertex on whichever GPU owns it (ld.global on the IPC-mapped peer
// pointer over NVLink), so there is no per-step collective at all.  The
// second-order bias uses rejection sampling; "is x a neighbour of the parent"
// looks at the parent's first `full_nbr_num` neighbours like the reference: a binary search in the
// id-sorted copy of the row (CsrView::sorted) when the whole row is inside the cap, a capped scan otherwise.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "csr_view.cuh"
#include "host_utils.h"

namespace glb {

CsrView csr_from_desc(const at::Tensor& desc);   // sampling.cu

__global__ void __launch_bounds__(256)
random_walk_kernel(const CsrView g, const int64_t* __restrict__ src, int64_t B, int L, float p, float q,
                   int64_t default_id, int full_nbr_num, const uint64_t* __restrict__ rng, uint32_t salt,
                   int64_t* __restrict__ out) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const bool second_order = !(p == 1.f && q == 1.f);
  const float inv_p = 1.f / p, inv_q = 1.f / q;
  const float wmax = fmaxf(1.f, fmaxf(inv_p, inv_q));
  int64_t cur = __ldg(src + b);
  int64_t prev = -1;
  RowRef prow;
  prow.deg = 0; prow.beg = 0; prow.indices = nullptr; prow.sorted = nullptr;
  for (int step = 0; step < L; ++step) {
    RowRef r = csr_row(g, cur);
    int64_t nxt = default_id;
    if (r.deg > 0) {
      uint4 rnd = rng4(rng, salt + (uint32_t)step * 0x632BE5ABu, (uint64_t)b);
      if (!second_order || prev < 0) {
        int64_t idx = r.cumw ? -1 : (int64_t)bounded64(rnd.x, rnd.y, (uint64_t)r.deg);
        if (idx < 0) {   // weighted first-order walk: inverse CDF on the in-row prefix sums
          float total = __ldg(r.cumw + r.beg + r.deg - 1);
          float target = u01(rnd.x) * total;
          int64_t lo = 0, hi = r.deg - 1;
          while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (__ldg(r.cumw + r.beg + mid) > target) hi = mid; else lo = mid + 1; }
          idx = lo;
        }
        nxt = __ldg(r.indices + r.beg + idx);
      } else {
        // node2vec: accept candidate x with prob w(x)/wmax, w = 1/p (x == parent), 1 (x ~ parent), 1/q otherwise
        uint32_t r0 = rnd.x, r1 = rnd.y, r2 = rnd.z;
        for (int tries = 0; tries < 64; ++tries) {
          int64_t idx = (int64_t)bounded64(r0, r1, (uint64_t)r.deg);
          int64_t x = __ldg(r.indices + r.beg + idx);
          float w;
          if (x == prev) w = inv_p;
          else {
            bool nb = false;
            if (prow.sorted != nullptr && prow.deg > 8 && prow.deg <= full_nbr_num) {
              int64_t lo = 0, hi = prow.deg - 1;          // <= 7 dependent probes instead of up to 100
              while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(prow.sorted + prow.beg + mid) < x) lo = mid + 1; else hi = mid; }
              nb = __ldg(prow.sorted + prow.beg + lo) == x;
            } else {
              int64_t lim = prow.deg < full_nbr_num ? prow.deg : full_nbr_num;
              for (int64_t i = 0; i < lim; ++i)
                if (__ldg(prow.indices + prow.beg + i) == x) { nb = true; break; }
            }
            w = nb ? 1.f : inv_q;
          }
          nxt = x;
          if (u01(r2) * wmax <= w) break;
          uint4 nr = rng4(rng, salt + (uint32_t)step * 0x632BE5ABu + (uint32_t)(tries + 1) * 0x9E3779B9u, (uint64_t)b);
          r0 = nr.x; r1 = nr.y; r2 = nr.z;
        }
      }
    }
    out[b * L + step] = nxt;
    prev = cur;
    prow = r;
    cur = nxt;
  }
}

at::Tensor random_walk(const at::Tensor& csr_desc, const at::Tensor& src, int64_t walk_len, double p, double q,
                       int64_t default_id, int64_t full_nbr_num, const at::Tensor& rng_state, int64_t salt) {
  check_cuda_i64(src, "src");
  c10::cuda::CUDAGuard guard(src.device());
  CsrView g = csr_from_desc(csr_desc);
  auto s = src.contiguous();
  int64_t B = s.numel();
  auto out = at::empty({B, walk_len}, s.options());
  if (B == 0 || walk_len == 0) return out;
  random_walk_kernel<<<(unsigned)((B + 255) / 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      g, s.data_ptr<int64_t>(), B, (int)walk_len, (float)p, (float)q, default_id, (int)full_nbr_num,
      reinterpret_cast<const uint64_t*>(rng_state.data_ptr<int64_t>()), (uint32_t)salt, out.data_ptr<int64_t>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

}  // namespace glb
