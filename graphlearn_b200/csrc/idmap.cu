// Device-resident id translation for NON-dense id spaces (hashed / string-derived / sparse ids).
//
// Every rank keeps the sorted array of the ids it owns in a SYMMETRIC (peer-mapped) buffer; any rank translates
//    id  -> vid = row * world + owner     owner = |id| % world, row = position of id in the owner's sorted array
//    vid -> id                             one peer load
// with a kernel that binary-searches the owner's array over NVLink.  This replaces the collective
// Partition -> all-to-all -> Stitch round trip (IdMap.to_vid / to_id on the portable path) and therefore the forced
// lock-step epochs: non-dense graphs stay on the peer kernels.
// Reference: AutoIndex id -> dense index hash map per server (graphlearn/src/core/graph/storage/auto_indexing.cc:21-33)
// reached through the sharded LookupNodes / Sampling RPCs (hash_partitioner.h:90-92).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "host_utils.h"

namespace glb {

struct IdMapView {
  PeerTable ids;                 // const int64_t* sorted local ids per rank
  int64_t nrows[kMaxWorld];
  int world;
};

static IdMapView idmap_from_desc(const at::Tensor& desc) {
  // desc (CPU int64): [world, nrows[8], ptrs[8]]
  TORCH_CHECK(desc.device().is_cpu() && desc.scalar_type() == at::kLong && desc.numel() == 1 + 2 * kMaxWorld);
  const int64_t* d = desc.data_ptr<int64_t>();
  IdMapView v;
  v.world = (int)d[0];
  TORCH_CHECK(v.world >= 1 && v.world <= kMaxWorld);
  for (int r = 0; r < kMaxWorld; ++r) { v.nrows[r] = d[1 + r]; v.ids.p[r] = reinterpret_cast<const void*>(d[1 + kMaxWorld + r]); }
  return v;
}

__global__ void idmap_to_vid_kernel(const IdMapView m, const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  const int owner = (int)((id < 0 ? -id : id) % m.world);
  const int64_t* arr = reinterpret_cast<const int64_t*>(m.ids.p[owner]);
  int64_t lo = 0, hi = m.nrows[owner];
  while (lo < hi) {                                  // lower bound over the owner's (possibly remote) sorted ids
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(arr + mid) < id) lo = mid + 1; else hi = mid;
  }
  out[i] = (lo < m.nrows[owner] && __ldg(arr + lo) == id) ? lo * m.world + owner : -1;
}

__global__ void idmap_to_id_kernel(const IdMapView m, const int64_t* __restrict__ vids, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = vids[i];
  if (v < 0) { out[i] = -1; return; }
  const int owner = (int)(v % m.world);
  const int64_t row = v / m.world;
  out[i] = row < m.nrows[owner] ? __ldg(reinterpret_cast<const int64_t*>(m.ids.p[owner]) + row) : -1;
}

at::Tensor idmap_translate(const at::Tensor& desc, const at::Tensor& x, bool to_vid) {
  check_cuda_i64(x, "ids");
  c10::cuda::CUDAGuard guard(x.device());
  const IdMapView m = idmap_from_desc(desc);
  auto xc = x.contiguous();
  auto out = at::empty_like(xc);
  const int64_t n = xc.numel();
  if (n == 0) return out;
  auto stream = at::cuda::getCurrentCUDAStream();
  if (to_vid) idmap_to_vid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(m, xc.data_ptr<int64_t>(), n, out.data_ptr<int64_t>());
  else idmap_to_id_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(m, xc.data_ptr<int64_t>(), n, out.data_ptr<int64_t>());
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

}  // namespace glb
