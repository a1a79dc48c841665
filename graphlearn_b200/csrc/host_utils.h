// Host-side helpers shared by the extension translation units.
#pragma once
#include <torch/extension.h>
#include "common.cuh"

namespace glb {

inline void check_cuda_i64(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kLong, name, " must be int64");
}

// Sharded dense table descriptor (CPU int64 tensor):
//   [world, dim, stride_elems, dtype_code(0=f32,1=bf16), nrows[8], ptr[8]]
struct TableView {
  PeerTable base;
  int64_t nrows[kMaxWorld];
  int world;
  int dim;
  int64_t stride;   // elements
  int dtype;        // 0 = fp32, 1 = bf16
};

inline TableView table_from_desc(const at::Tensor& desc) {
  TORCH_CHECK(desc.device().is_cpu() && desc.scalar_type() == at::kLong &&
                  desc.numel() == 4 + 2 * kMaxWorld,
              "table desc must be a CPU int64 tensor of 20 entries");
  const int64_t* d = desc.data_ptr<int64_t>();
  TableView t;
  t.world = (int)d[0];
  t.dim = (int)d[1];
  t.stride = d[2];
  t.dtype = (int)d[3];
  TORCH_CHECK(t.world >= 1 && t.world <= kMaxWorld, "bad world size in table desc");
  for (int r = 0; r < kMaxWorld; ++r) {
    t.nrows[r] = d[4 + r];
    t.base.p[r] = reinterpret_cast<const void*>(d[4 + kMaxWorld + r]);
  }
  return t;
}

}  // namespace glb
