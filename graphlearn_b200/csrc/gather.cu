// K5 / K6: feature-row gather and fused gather-aggregate over peer-resident
// feature tables (hot paths (b) and (c) of BASELINE.json).
//
// A sharded table is addressed as base[owner] + row * stride with owner/row
// derived arithmetically from the vid, so a warp pulls a remote 400-512 B
// feature row with one coalesced request per 16 B lane over NVLink - no
// id partition / all-to-all / stitch like the reference's LookupNodes
// (graphlearn/src/core/graph/local_noder.cc:85-97 behind
// graphlearn/src/core/runner/op_runner.h:87-116) and no host hop.
//
// Aggregation semantics follow graphlearn/src/core/operator/aggregator/*
// (sum / mean / min / max / prod segment reduce of float attributes).
#include <torch/extension.h>
#include <cmath>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cfloat>
#include "host_utils.h"

namespace glb {

enum AggMode : int { kSum = 0, kMean = 1, kMax = 2, kMin = 3, kProd = 4 };

// load 4 consecutive features [4c, 4c+4) of a row as fp32
template <int DTYPE>
__device__ __forceinline__ float4 load4(const void* row, int c, int scale_off = 0) {
  if constexpr (DTYPE == 0) {
    return ld_nc_f4(reinterpret_cast<const float4*>(row) + c);
  } else if constexpr (DTYPE == 2) {
    // fp8 block-scaled row: 4 e4m3 bytes + the bf16 scale of their 32-element block
    const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(row) + c);
    const float s = __uint_as_float(ld_nc_u16(reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(row) + scale_off) + (c >> 3)) << 16);
    const float2 a = e4m3x2_to_float2(q), b = e4m3x2_to_float2(q >> 16);
    return make_float4(a.x * s, a.y * s, b.x * s, b.y * s);
  } else {
    uint2 u = ld_nc_u2(reinterpret_cast<const uint2*>(row) + c);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

__device__ __forceinline__ const void* row_ptr(const TableView& t, int64_t vid) {
  if (vid < 0) return nullptr;
  int owner = (int)(vid % t.world);
  int64_t row = vid / t.world;
  if (row >= t.nrows[owner]) return nullptr;
  size_t esz = (size_t)table_esize(t.dtype);
  if (t.cmap != nullptr && owner != t.self) {   // replica cache of remote rows (read paths only)
    const int s = __ldg(t.cmap + vid);
    if (s >= 0) return t.cbase + (size_t)s * (size_t)t.stride * esz;
  }
  return reinterpret_cast<const char*>(t.base.p[owner]) + (size_t)row * (size_t)t.stride * esz;
}

template <typename OutT>
__device__ __forceinline__ void store4(OutT* out, int c, int dim, float4 v) {
  int f = 4 * c;
  if constexpr (sizeof(OutT) == 4) {
    if (f + 3 < dim && (dim & 3) == 0) {
      reinterpret_cast<float4*>(out)[c] = v;
    } else {
      if (f < dim) out[f] = v.x;
      if (f + 1 < dim) out[f + 1] = v.y;
      if (f + 2 < dim) out[f + 2] = v.z;
      if (f + 3 < dim) out[f + 3] = v.w;
    }
  } else {
    if (f + 3 < dim && (dim & 3) == 0) {
      uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(out)[c] = u;
    } else {
      if (f < dim) out[f] = __float2bfloat16(v.x);
      if (f + 1 < dim) out[f + 1] = __float2bfloat16(v.y);
      if (f + 2 < dim) out[f + 2] = __float2bfloat16(v.z);
      if (f + 3 < dim) out[f + 3] = __float2bfloat16(v.w);
    }
  }
}

// warp per output row
template <int DTYPE, typename OutT>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const TableView t, const int64_t* __restrict__ vids, int64_t n,
                   OutT* __restrict__ out, int64_t out_stride, float fill) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n) return;
  const void* rp = row_ptr(t, __ldg(vids + w));
  int chunks = (t.dim + 3) >> 2;
  OutT* o = out + w * out_stride;
  const int soff = fp8_scale_offset(t.dim);
  for (int c = lane; c < chunks; c += 32) {
    float4 v = rp ? load4<DTYPE>(rp, c, soff) : make_float4(fill, fill, fill, fill);
    store4<OutT>(o, c, t.dim, v);
  }
}

__device__ __forceinline__ float agg_init(int mode) {
  return mode == kMax ? -FLT_MAX : mode == kMin ? FLT_MAX : mode == kProd ? 1.f : 0.f;
}
__device__ __forceinline__ float agg_op(int mode, float a, float b) {
  switch (mode) {
    case kMax: return fmaxf(a, b);
    case kMin: return fminf(a, b);
    case kProd: return a * b;
    default: return a + b;
  }
}

// warp per segment; dense fan-out layout (offsets == nullptr -> segment s is
// vids[s*k .. s*k+k)) or ragged (offsets[S+1]).
template <int DTYPE>
__global__ void __launch_bounds__(256)
gather_agg_kernel(const TableView t, const int64_t* __restrict__ vids,
                  const int64_t* __restrict__ offsets, int64_t S, int k, int mode,
                  float* __restrict__ out, int64_t out_stride) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= S) return;
  int64_t beg = offsets ? __ldg(offsets + w) : w * (int64_t)k;
  int64_t end = offsets ? __ldg(offsets + w + 1) : beg + k;
  int chunks = (t.dim + 3) >> 2;
  for (int c0 = 0; c0 < chunks; c0 += 32) {
    int c = c0 + lane;
    float i0 = agg_init(mode);
    float4 acc = make_float4(i0, i0, i0, i0);
    int64_t cnt = 0;
    // 4 independent row loads in flight per lane
    int64_t i = beg;
    for (; i + 4 <= end; i += 4) {
      const void* r0 = row_ptr(t, __ldg(vids + i));
      const void* r1 = row_ptr(t, __ldg(vids + i + 1));
      const void* r2 = row_ptr(t, __ldg(vids + i + 2));
      const void* r3 = row_ptr(t, __ldg(vids + i + 3));
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v0 = z, v1 = z, v2 = z, v3 = z;
      if (c < chunks) {
        if (r0) v0 = load4<DTYPE>(r0, c);
        if (r1) v1 = load4<DTYPE>(r1, c);
        if (r2) v2 = load4<DTYPE>(r2, c);
        if (r3) v3 = load4<DTYPE>(r3, c);
      }
      acc.x = agg_op(mode, agg_op(mode, agg_op(mode, agg_op(mode, acc.x, v0.x), v1.x), v2.x), v3.x);
      acc.y = agg_op(mode, agg_op(mode, agg_op(mode, agg_op(mode, acc.y, v0.y), v1.y), v2.y), v3.y);
      acc.z = agg_op(mode, agg_op(mode, agg_op(mode, agg_op(mode, acc.z, v0.z), v1.z), v2.z), v3.z);
      acc.w = agg_op(mode, agg_op(mode, agg_op(mode, agg_op(mode, acc.w, v0.w), v1.w), v2.w), v3.w);
      cnt += 4;
    }
    for (; i < end; ++i) {
      const void* r0 = row_ptr(t, __ldg(vids + i));
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 && c < chunks) v0 = load4<DTYPE>(r0, c);
      acc.x = agg_op(mode, acc.x, v0.x); acc.y = agg_op(mode, acc.y, v0.y);
      acc.z = agg_op(mode, acc.z, v0.z); acc.w = agg_op(mode, acc.w, v0.w);
      cnt += 1;
    }
    if (mode == kMean && cnt > 0) {
      float inv = 1.f / (float)cnt;
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    }
    if (cnt == 0) acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < chunks) store4<float>(out + w * out_stride, c, t.dim, acc);
  }
}

// K9 backward / embedding update: rows[vid] += grad (atomics; peer-capable)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const TableView t, const int64_t* __restrict__ vids, int64_t n,
                        const float* __restrict__ grad, int64_t grad_stride, float scale) {
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= n) return;
  float* rp = const_cast<float*>(reinterpret_cast<const float*>(row_ptr(t, __ldg(vids + w))));
  if (!rp) return;
  for (int f = lane; f < t.dim; f += 32) atomicAdd(rp + f, scale * grad[w * grad_stride + f]);
}

// ---------------------------------------------------------------------------
at::Tensor gather_rows(const at::Tensor& table_desc, const at::Tensor& vids, bool out_bf16,
                       double fill) {
  check_cuda_i64(vids, "vids");
  c10::cuda::CUDAGuard guard(vids.device());
  TableView t = table_from_desc(table_desc);
  auto v = vids.contiguous();
  int64_t n = v.numel();
  auto out = at::empty({n, t.dim}, v.options().dtype(out_bf16 ? at::kBFloat16 : at::kFloat));
  if (n == 0) return out;
  auto stream = at::cuda::getCurrentCUDAStream();
  unsigned blocks = (unsigned)((n * 32 + 255) / 256);
  const int64_t* vp = v.data_ptr<int64_t>();
#define LAUNCH(DT, OT, optr) \
  gather_rows_kernel<DT, OT><<<blocks, 256, 0, stream>>>(t, vp, n, optr, (int64_t)t.dim, (float)fill)
  if (t.dtype == 2 && !out_bf16) LAUNCH(2, float, out.data_ptr<float>());
  else if (t.dtype == 2) LAUNCH(2, __nv_bfloat16, reinterpret_cast<__nv_bfloat16*>(out.data_ptr()));
  else if (t.dtype == 0 && !out_bf16) LAUNCH(0, float, out.data_ptr<float>());
  else if (t.dtype == 0 && out_bf16) LAUNCH(0, __nv_bfloat16, reinterpret_cast<__nv_bfloat16*>(out.data_ptr()));
  else if (t.dtype == 1 && !out_bf16) LAUNCH(1, float, out.data_ptr<float>());
  else LAUNCH(1, __nv_bfloat16, reinterpret_cast<__nv_bfloat16*>(out.data_ptr()));
#undef LAUNCH
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

at::Tensor gather_agg(const at::Tensor& table_desc, const at::Tensor& vids,
                      const c10::optional<at::Tensor>& offsets, int64_t k, int64_t mode) {
  check_cuda_i64(vids, "vids");
  c10::cuda::CUDAGuard guard(vids.device());
  TableView t = table_from_desc(table_desc);
  TORCH_CHECK(t.dtype != 2, "gather_agg: fp8 tables are served by gather_rows and the fused layer kernel");
  auto v = vids.contiguous();
  int64_t S;
  const int64_t* op = nullptr;
  at::Tensor oc;
  if (offsets.has_value()) {
    oc = offsets->contiguous();
    check_cuda_i64(oc, "offsets");
    S = oc.numel() - 1;
    op = oc.data_ptr<int64_t>();
  } else {
    TORCH_CHECK(k > 0 && v.numel() % k == 0, "vids must be [S, k]");
    S = v.numel() / k;
  }
  auto out = at::empty({S, t.dim}, v.options().dtype(at::kFloat));
  if (S == 0) return out;
  auto stream = at::cuda::getCurrentCUDAStream();
  unsigned blocks = (unsigned)((S * 32 + 255) / 256);
  if (t.dtype == 0)
    gather_agg_kernel<0><<<blocks, 256, 0, stream>>>(t, v.data_ptr<int64_t>(), op, S, (int)k, (int)mode,
                                                    out.data_ptr<float>(), (int64_t)t.dim);
  else
    gather_agg_kernel<1><<<blocks, 256, 0, stream>>>(t, v.data_ptr<int64_t>(), op, S, (int)k, (int)mode,
                                                    out.data_ptr<float>(), (int64_t)t.dim);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

void scatter_add_rows(const at::Tensor& table_desc, const at::Tensor& vids, const at::Tensor& grad,
                      double scale) {
  check_cuda_i64(vids, "vids");
  c10::cuda::CUDAGuard guard(vids.device());
  TableView t = table_from_desc(table_desc);
  TORCH_CHECK(t.dtype == 0, "scatter_add_rows needs an fp32 table");
  t.cmap = nullptr;   // updates always go to the owner, never to a replica
  auto v = vids.contiguous();
  auto g = grad.contiguous();
  TORCH_CHECK(g.is_cuda() && g.scalar_type() == at::kFloat && g.dim() == 2 && g.size(1) == t.dim &&
              g.size(0) == v.numel(), "grad must be fp32 [n, dim]");
  int64_t n = v.numel();
  if (n == 0) return;
  unsigned blocks = (unsigned)((n * 32 + 255) / 256);
  scatter_add_rows_kernel<<<blocks, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      t, v.data_ptr<int64_t>(), n, g.data_ptr<float>(), (int64_t)t.dim, (float)scale);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------
// K9: sparse Adam on sharded embedding tables.  One warp per touched row (ids unique within the call): the row of the
// weight table and of both moment tables lives on the owner GPU and is updated in place through the peer pointers -
// the asynchronous parameter-server update of the reference's AdamAsyncOptimizer (examples/tf/trainer.py:111-116)
// without a server: no dense gradient, no all-reduce.  Rows touched by two ranks at once race like they do on a PS.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sparse_adam_rows_kernel(const TableView w, const TableView m, const TableView v, const int64_t* __restrict__ vids, int64_t n,
                        const float* __restrict__ grad, int64_t grad_stride, float lr, float b1, float b2, float eps, float bc1, float bc2) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t vid = __ldg(vids + r);
  float* wp = const_cast<float*>(reinterpret_cast<const float*>(row_ptr(w, vid)));
  float* mp = const_cast<float*>(reinterpret_cast<const float*>(row_ptr(m, vid)));
  float* vp = const_cast<float*>(reinterpret_cast<const float*>(row_ptr(v, vid)));
  if (!wp || !mp || !vp) return;
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int f = lane; f < w.dim; f += 32) {
    const float g = grad[r * grad_stride + f];
    const float mm = b1 * mp[f] + (1.f - b1) * g;
    const float vv = b2 * vp[f] + (1.f - b2) * g * g;
    mp[f] = mm; vp[f] = vv;
    wp[f] -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
  }
}

void sparse_adam_rows(const at::Tensor& w_desc, const at::Tensor& m_desc, const at::Tensor& v_desc, const at::Tensor& vids,
                      const at::Tensor& grad, double lr, double b1, double b2, double eps, int64_t step) {
  check_cuda_i64(vids, "vids");
  c10::cuda::CUDAGuard guard(vids.device());
  TableView w = table_from_desc(w_desc), m = table_from_desc(m_desc), v = table_from_desc(v_desc);
  TORCH_CHECK(w.dtype == 0 && m.dtype == 0 && v.dtype == 0 && m.dim == w.dim && v.dim == w.dim, "sparse_adam_rows needs fp32 tables of one width");
  w.cmap = m.cmap = v.cmap = nullptr;        // updates always go to the owner
  auto ids = vids.contiguous();
  auto g = grad.contiguous();
  TORCH_CHECK(g.is_cuda() && g.scalar_type() == at::kFloat && g.dim() == 2 && g.size(1) == w.dim && g.size(0) == ids.numel(),
              "grad must be fp32 [n, dim]");
  const int64_t n = ids.numel();
  if (n == 0) return;
  TORCH_CHECK(step >= 1, "Adam step counts from 1");
  const float bc1 = 1.f - std::pow((float)b1, (float)step), bc2 = 1.f - std::pow((float)b2, (float)step);
  sparse_adam_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      w, m, v, ids.data_ptr<int64_t>(), n, g.data_ptr<float>(), (int64_t)w.dim, (float)lr, (float)b1, (float)b2, (float)eps, bc1, bc2);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------
// Measurement kernel: the random-row read ceiling of the memory system.  One thread per 16-byte chunk of a gathered row
// (no reduction, no staging, millions of independent loads): out[i, :] = table[idx[i], :row_bytes].  The fused layer
// kernels cannot beat this number; tools/bench_gather_floor.py charts it against the row size.
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(256) gather_copy16_kernel(const uint4* __restrict__ table, int64_t stride16, const int64_t* __restrict__ idx,
                                                            int64_t n_rows, int chunks, uint4* __restrict__ out, int read_only) {
  // U independent 16-byte loads in flight per thread (chunks t, t + T, ..., T = total threads)
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n_rows * chunks;
  uint4 v[U];
  int64_t tt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    tt[u] = t0 + u * T;
    v[u] = make_uint4(0u, 0u, 0u, 0u);
    if (tt[u] < total) {
      const int64_t row = tt[u] / chunks;
      const int c = (int)(tt[u] - row * chunks);
      v[u] = ld_nc_u4(table + __ldg(idx + row) * stride16 + c);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (tt[u] < total) {
      if (!read_only) out[tt[u]] = v[u];
      else if ((v[u].x ^ v[u].y ^ v[u].z ^ v[u].w) == 0x9E3779B9u && v[u].x == 0x7F4A7C15u) out[0] = v[u];   // keeps the load alive
    }
  }
}

void gather_copy16(const at::Tensor& table, const at::Tensor& idx, int64_t row_bytes, const at::Tensor& out, bool read_only, int64_t unroll) {
  TORCH_CHECK(table.is_cuda() && table.dim() == 2 && table.is_contiguous() && (table.stride(0) * table.element_size()) % 16 == 0);
  check_cuda_i64(idx, "idx");
  TORCH_CHECK(row_bytes % 16 == 0 && row_bytes <= table.stride(0) * (int64_t)table.element_size());
  const int64_t n = idx.numel();
  TORCH_CHECK(out.is_cuda() && out.is_contiguous() && out.numel() * (int64_t)out.element_size() >= n * row_bytes);
  if (n == 0) return;
  c10::cuda::CUDAGuard guard(table.device());
  const int chunks = (int)(row_bytes / 16);
  const int64_t total = n * chunks;
  const unsigned blocks = (unsigned)((total + 256 * unroll - 1) / (256 * unroll));
  auto stream = at::cuda::getCurrentCUDAStream();
  const uint4* tp = reinterpret_cast<const uint4*>(table.data_ptr());
  const int64_t s16 = table.stride(0) * (int64_t)table.element_size() / 16;
  uint4* op = reinterpret_cast<uint4*>(out.data_ptr());
#define GC(UU) gather_copy16_kernel<UU><<<blocks, 256, 0, stream>>>(tp, s16, idx.data_ptr<int64_t>(), n, chunks, op, read_only ? 1 : 0)
  if (unroll == 1) GC(1); else if (unroll == 2) GC(2); else if (unroll == 4) GC(4); else if (unroll == 8) GC(8);
  else TORCH_CHECK(false, "unroll must be 1, 2, 4 or 8");
#undef GC
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace glb
