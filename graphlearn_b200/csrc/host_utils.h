// Host-side helpers shared by the extension translation units.
#pragma once
#include <torch/extension.h>
#include "common.cuh"

namespace glb {

int sm_count();   // SMs of the current device (sage_fused.cu)

inline void check_cuda_i64(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kLong, name, " must be int64");
}

// Sharded dense table descriptor (CPU int64 tensor):
//   [world, dim, stride_elems, dtype_code(0=f32,1=bf16,2=fp8 e4m3 block-scaled), nrows[8], ptr[8]]
// fp8 rows (dtype 2, stride in BYTES): [dim e4m3 bytes | pad to 16] [one bf16 scale per 32-element block] [pad]: value = q * scale
// optionally followed by the replica cache of remote rows (N17):
//   [self_rank, cache_map ptr (int32[max_vid+1]: slot or -1), cache rows ptr (same stride/dtype)]
struct TableView {
  PeerTable base;
  int64_t nrows[kMaxWorld];
  int world;
  int dim;
  int64_t stride;   // elements
  int dtype;        // 0 = fp32, 1 = bf16, 2 = fp8 (e4m3, one bf16 scale per 32 elements; stride counts bytes)
  int self;               // rank owning the cache (only meaningful when cmap != nullptr)
  const int32_t* cmap;    // vid -> cache slot, -1 = not cached; nullptr = no cache
  const char* cbase;      // cache rows
};

// bytes per stored element / offset of the scale area of an fp8 row
__host__ __device__ inline int table_esize(int dtype) { return dtype == 0 ? 4 : dtype == 1 ? 2 : 1; }
__host__ __device__ inline int fp8_scale_offset(int dim) { return (dim + 15) / 16 * 16; }

inline TableView table_from_desc(const at::Tensor& desc) {
  TORCH_CHECK(desc.device().is_cpu() && desc.scalar_type() == at::kLong &&
                  (desc.numel() == 4 + 2 * kMaxWorld || desc.numel() == 7 + 2 * kMaxWorld),
              "table desc must be a CPU int64 tensor of 20 (or 23, with cache) entries");
  const int64_t* d = desc.data_ptr<int64_t>();
  TableView t;
  t.world = (int)d[0];
  t.dim = (int)d[1];
  t.stride = d[2];
  t.dtype = (int)d[3];
  TORCH_CHECK(t.world >= 1 && t.world <= kMaxWorld, "bad world size in table desc");
  TORCH_CHECK(t.dtype >= 0 && t.dtype <= 2, "bad dtype code in table desc");
  for (int r = 0; r < kMaxWorld; ++r) {
    t.nrows[r] = d[4 + r];
    t.base.p[r] = reinterpret_cast<const void*>(d[4 + kMaxWorld + r]);
  }
  t.self = 0;
  t.cmap = nullptr;
  t.cbase = nullptr;
  if (desc.numel() == 7 + 2 * kMaxWorld && d[5 + 2 * kMaxWorld] != 0) {
    t.self = (int)d[4 + 2 * kMaxWorld];
    t.cmap = reinterpret_cast<const int32_t*>(d[5 + 2 * kMaxWorld]);
    t.cbase = reinterpret_cast<const char*>(d[6 + 2 * kMaxWorld]);
  }
  return t;
}

}  // namespace glb
