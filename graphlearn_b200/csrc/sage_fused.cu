// K6+K7 fused: GraphSAGE / EgoSAGE layer forward in ONE kernel
//
//     out[m, :] = act( [ x_self[m] || agg_j x_nbr[m, j] ] . W^T + b )
//
// * gather + aggregate: 16 warps pull the self row and the k neighbour rows of
//   128 destination nodes straight from the (peer-mapped) feature shards -
//   local HBM or a remote GPU over NVLink, picked per row by vid % world -
//   reduce them in fp32 registers and write the bf16 A tile into shared
//   memory in the UMMA K-major SWIZZLE_128B canonical layout.  The rows never
//   round-trip through HBM ("agg(X).W" instead of materialising [B*k, D]).
// * weights: the pre-swizzled bf16 image of W is fetched by the TMA engine
//   (cp.async.bulk -> UBLKCP) while the gather is in flight.
// * GEMM: one elected thread issues tcgen05.mma (M=128, N<=256, K=16) with
//   the fp32 accumulator in TMEM; completion is tracked with an mbarrier via
//   tcgen05.commit.
// * epilogue: tcgen05.ld -> bias -> ReLU -> bf16/fp32 store.
//
// Math parity: EgoSAGEConv (graphlearn/python/nn/tf/layers/ego_sage_conv.py:71-106):
// agg in {mean,sum} over neighbor.reshape(-1,k,d), out = W.[x || agg]; 'gcn'
// = mean over {x} U nbrs then W.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "host_utils.h"
#include "umma.cuh"

namespace glb {

constexpr int kTileM = 128;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

enum SageMode : int { kConcatMean = 0, kConcatSum = 1, kGcnMean = 2 };

struct SageParams {
  TableView tself;
  TableView tnbr;
  const int64_t* self_vids;   // [M] or null (identity)
  const int64_t* nbr_vids;    // [M, k] or null (identity m*k+j)
  const void* w_img;          // bf16 image, (K_total/64) blocks of [N x 64] SW128
  const float* bias;          // [N] (padded) or null
  void* out;                  // [M, n_out]
  __nv_bfloat16* a_save;      // [M, K_total] or null
  int64_t out_stride;
  int M, k;
  int kp_self, kp_nbr;        // padded K of each half, in {0,64,128,256,512}
  int mode;
  int N;                      // padded output width: multiple of 64, <= 256
  int n_out;                  // real output width
  int relu;
  int out_bf16;
  int tmem_cols;
};

__device__ __forceinline__ float4 load4_rt(const void* row, int c, int dtype) {
  if (dtype == 0) return ld_nc_f4(reinterpret_cast<const float4*>(row) + c);
  uint2 u = ld_nc_u2(reinterpret_cast<const uint2*>(row) + c);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ const char* table_row(const TableView& t, int64_t vid) {
  if (vid < 0) return nullptr;
  int owner = (int)(vid % t.world);
  int64_t row = vid / t.world;
  if (row >= t.nrows[owner]) return nullptr;
  size_t esz = t.dtype == 0 ? 4 : 2;
  return reinterpret_cast<const char*>(t.base.p[owner]) + (size_t)row * (size_t)t.stride * esz;
}

__device__ __forceinline__ float4 mask_tail(float4 v, int f, int dim) {
  if (f + 3 >= dim) {
    if (f >= dim) v.x = 0.f;
    if (f + 1 >= dim) v.y = 0.f;
    if (f + 2 >= dim) v.z = 0.f;
    if (f + 3 >= dim) v.w = 0.f;
  }
  return v;
}

// store 4 consecutive K elements (kcol multiple of 4) of tile row r
__device__ __forceinline__ void put_a(uint8_t* sA, int r, int kcol, float4 v) {
  int kb = kcol >> 6;
  uint32_t off = (uint32_t)kb * (kTileM * 128) + umma::sw128_offset((uint32_t)r, (uint32_t)(kcol & 63));
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(sA + off) = u;
}

// CPL = 16-byte-lane chunks per lane for the neighbour half: kp_nbr = 128*CPL
// (kp_nbr == 64 uses CPL = 1 with the upper half-warp idle).
template <int CPL>
__global__ void __launch_bounds__(kThreads, 1) sage_fused_fwd_kernel(const SageParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int k_total = p.kp_self + p.kp_nbr;
  const int nkb = k_total >> 6;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (size_t)nkb * (kTileM * 128);
  const uint32_t w_kb_bytes = (uint32_t)p.N * 128u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)nkb * w_kb_bytes);
  uint64_t* bar_w = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    umma::mbar_init(bar_w, 1);
    umma::mbar_init(bar_mma, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // --- weights: TMA bulk copies of the pre-swizzled image, overlapped with the gather
  if (tid == 0) {
    umma::mbar_arrive_expect_tx(bar_w, (uint32_t)nkb * w_kb_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w_img);
    for (int kb = 0; kb < nkb; ++kb)
      umma::bulk_g2s(sW + (size_t)kb * w_kb_bytes, src + (size_t)kb * w_kb_bytes, w_kb_bytes, bar_w);
  }

  // --- gather + aggregate -> A tile (bf16, SW128 K-major)
  const int m0 = blockIdx.x * kTileM;
  const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
  const int k = p.k;
  for (int r = warp; r < kTileM; r += kWarps) {
    const int m = m0 + r;
    const bool valid = m < p.M;
    // self row
    const char* sp = nullptr;
    if (valid && (p.kp_self > 0 || p.mode == kGcnMean)) {
      int64_t sv = p.self_vids ? __ldg(p.self_vids + m) : (int64_t)m;
      sp = table_row(p.tself, sv);
    }
    for (int c4 = lane; c4 < (p.kp_self >> 2); c4 += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sp && 4 * c4 < d_self) v = mask_tail(load4_rt(sp, c4, p.tself.dtype), 4 * c4, d_self);
      put_a(sA, r, 4 * c4, v);
      if (p.a_save && valid) {
        uint2 u; u.x = pack_bf16x2(v.x, v.y); u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(p.a_save + (size_t)m * k_total + 4 * c4) = u;
      }
    }
    // neighbour rows
    float4 acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jb = 0; jb < k; jb += 32) {
      const int cnt = min(32, k - jb);
      const char* myp = nullptr;
      if (valid && lane < cnt) {
        int64_t idx = (int64_t)m * k + jb + lane;
        int64_t nv = p.nbr_vids ? __ldg(p.nbr_vids + idx) : idx;
        myp = table_row(p.tnbr, nv);
      }
      for (int j0 = 0; j0 < cnt; j0 += 5) {
        float4 v[5][CPL];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int j = j0 + u;
          const char* rp = reinterpret_cast<const char*>(
              __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(myp), j < cnt ? j : 0));
          if (j >= cnt) rp = nullptr;
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            const int c4 = lane + 32 * i;
            v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rp && 4 * c4 < d_nbr) v[u][i] = load4_rt(rp, c4, p.tnbr.dtype);
          }
        }
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            acc[i].x += v[u][i].x; acc[i].y += v[u][i].y;
            acc[i].z += v[u][i].z; acc[i].w += v[u][i].w;
          }
      }
    }
    float scale = 1.f;
    if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
    else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c4 = lane + 32 * i;
      if (c4 < (p.kp_nbr >> 2)) {
        float4 a = mask_tail(acc[i], 4 * c4, d_nbr);
        if (p.mode == kGcnMean && sp && 4 * c4 < p.tself.dim) {
          float4 s = mask_tail(load4_rt(sp, c4, p.tself.dtype), 4 * c4, p.tself.dim);
          a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        put_a(sA, r, p.kp_self + 4 * c4, a);
        if (p.a_save && valid) {
          uint2 u; u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
          *reinterpret_cast<uint2*>(p.a_save + (size_t)m * k_total + p.kp_self + 4 * c4) = u;
        }
      }
    }
  }
  umma::fence_proxy_async_smem();     // generic-proxy st.shared -> visible to tcgen05 (async proxy)
  __syncthreads();

  // --- GEMM: one thread issues all MMAs; accumulator lives in TMEM
  if (tid == 0) {
    umma::mbar_wait(bar_w, 0);
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc_bf16(kTileM, p.N);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t a_base = umma::smem_u32(sA + (size_t)kb * (kTileM * 128));
      const uint32_t b_base = umma::smem_u32(sW + (size_t)kb * w_kb_bytes);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        umma::mma_bf16_ss(tmem_base, umma::make_desc_sw128(a_base + k4 * 32),
                          umma::make_desc_sw128(b_base + k4 * 32), idesc, (kb | k4) ? 1u : 0u);
      }
    }
    umma::mma_commit(bar_mma);
  }
  __syncwarp();
  umma::mbar_wait(bar_mma, 0);
  umma::tc_fence_after();

  // --- epilogue: TMEM -> registers -> bias/ReLU -> global
  {
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int g = warp >> 2;             // column group
    const int cols_per_group = p.N >> 2; // multiple of 16
    const int row = q * 32 + lane;
    const int m = m0 + row;
    for (int c0 = 0; c0 < cols_per_group; c0 += 16) {
      const int n0 = g * cols_per_group + c0;
      uint32_t v[16];
      umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, v);
      umma::tmem_ld_wait();
      if (m < p.M && n0 < p.n_out) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(v[i]);
          if (p.bias) x += __ldg(p.bias + n0 + i);
          if (p.relu) x = fmaxf(x, 0.f);
          f[i] = x;
        }
        const bool full = (n0 + 16 <= p.n_out);
        if (p.out_bf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.out_stride + n0;
          if (full && (p.out_stride & 7) == 0) {
            uint4 a, b;
            a.x = pack_bf16x2(f[0], f[1]); a.y = pack_bf16x2(f[2], f[3]);
            a.z = pack_bf16x2(f[4], f[5]); a.w = pack_bf16x2(f[6], f[7]);
            b.x = pack_bf16x2(f[8], f[9]); b.y = pack_bf16x2(f[10], f[11]);
            b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
            reinterpret_cast<uint4*>(o)[0] = a;
            reinterpret_cast<uint4*>(o)[1] = b;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + i < p.n_out) o[i] = __float2bfloat16(f[i]);
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + (size_t)m * p.out_stride + n0;
          if (full && (p.out_stride & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + i < p.n_out) o[i] = f[i];
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// bf16 row-major [n_real, k_total] -> SW128 K-major image with N (>= n_real) rows per k-block
__global__ void pack_sw128_kernel(const __nv_bfloat16* __restrict__ w, int n_real, int N, int k_total,
                                  uint8_t* __restrict__ img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte chunk each
  int nkb = k_total >> 6;
  int total = nkb * N * 8;
  if (idx >= total) return;
  int c = idx & 7;
  int n = (idx >> 3) % N;
  int kb = (idx >> 3) / N;
  uint4 val = make_uint4(0, 0, 0, 0);
  if (n < n_real) val = *reinterpret_cast<const uint4*>(w + (size_t)n * k_total + kb * 64 + c * 8);
  size_t off = (size_t)kb * N * 128 + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)((c ^ (n & 7)) * 16);
  *reinterpret_cast<uint4*>(img + off) = val;
}

// ---------------------------------------------------------------------------
static int pad_k(int d) {
  if (d <= 0) return 0;
  if (d <= 64) return 64;
  if (d <= 128) return 128;
  if (d <= 256) return 256;
  TORCH_CHECK(d <= 512, "fused SAGE layer supports feature dims up to 512");
  return 512;
}

int64_t sage_pad_k(int64_t d) { return pad_k((int)d); }

int64_t sage_smem_bytes(int64_t k_total, int64_t N) {
  return 1024 + (k_total / 64) * (kTileM * 128) + (k_total / 64) * N * 128 + 64;
}

at::Tensor pack_weight_sw128(const at::Tensor& w_padded, int64_t N) {
  TORCH_CHECK(w_padded.is_cuda() && w_padded.scalar_type() == at::kBFloat16 && w_padded.dim() == 2 &&
              w_padded.is_contiguous(), "w_padded must be a contiguous CUDA bf16 [n, k_total] tensor");
  int64_t n_real = w_padded.size(0), k_total = w_padded.size(1);
  TORCH_CHECK(k_total % 64 == 0 && N % 8 == 0 && N >= n_real, "bad padded weight shape");
  c10::cuda::CUDAGuard guard(w_padded.device());
  auto img = at::empty({(k_total / 64) * N * 64}, w_padded.options());
  int total = (int)((k_total / 64) * N * 8);
  pack_sw128_kernel<<<(total + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      reinterpret_cast<const __nv_bfloat16*>(w_padded.data_ptr()), (int)n_real, (int)N, (int)k_total,
      reinterpret_cast<uint8_t*>(img.data_ptr()));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return img;
}

std::vector<at::Tensor> sage_fused_forward(const at::Tensor& tself_desc,
                                           const c10::optional<at::Tensor>& self_vids,
                                           const at::Tensor& tnbr_desc,
                                           const c10::optional<at::Tensor>& nbr_vids, int64_t M,
                                           int64_t k, int64_t mode, const at::Tensor& w_img,
                                           const c10::optional<at::Tensor>& bias, int64_t N,
                                           int64_t n_out, bool relu, bool out_bf16, bool save_a) {
  TORCH_CHECK(w_img.is_cuda() && w_img.scalar_type() == at::kBFloat16, "w_img must be CUDA bf16");
  c10::cuda::CUDAGuard guard(w_img.device());
  SageParams p;
  p.tself = table_from_desc(tself_desc);
  p.tnbr = table_from_desc(tnbr_desc);
  p.M = (int)M; p.k = (int)k; p.mode = (int)mode;
  p.kp_self = (mode == kGcnMean) ? 0 : pad_k(p.tself.dim);
  p.kp_nbr = pad_k(p.tnbr.dim);
  if (mode == kGcnMean) TORCH_CHECK(p.tself.dim == p.tnbr.dim, "gcn mode needs equal dims");
  const int k_total = p.kp_self + p.kp_nbr;
  TORCH_CHECK(N % 64 == 0 && N >= 64 && N <= 256, "padded N must be a multiple of 64 in [64, 256]");
  TORCH_CHECK(n_out <= N && n_out >= 1);
  TORCH_CHECK(w_img.numel() == (int64_t)k_total * N, "weight image size mismatch: expected ", k_total * N);
  TORCH_CHECK((reinterpret_cast<uintptr_t>(w_img.data_ptr()) & 15) == 0, "weight image must be 16 B aligned");
  size_t smem = (size_t)sage_smem_bytes(k_total, N);
  TORCH_CHECK(smem <= 232448, "tile does not fit in shared memory (", smem, " B); use the unfused path");
  at::Tensor sv, nv, b;
  p.self_vids = nullptr; p.nbr_vids = nullptr; p.bias = nullptr;
  if (self_vids.has_value()) { sv = self_vids->contiguous(); check_cuda_i64(sv, "self_vids");
    TORCH_CHECK(sv.numel() == M); p.self_vids = sv.data_ptr<int64_t>(); }
  if (nbr_vids.has_value()) { nv = nbr_vids->contiguous(); check_cuda_i64(nv, "nbr_vids");
    TORCH_CHECK(nv.numel() == M * k); p.nbr_vids = nv.data_ptr<int64_t>(); }
  if (bias.has_value()) { b = bias->contiguous();
    TORCH_CHECK(b.is_cuda() && b.scalar_type() == at::kFloat && b.numel() == N, "bias must be fp32 [N]");
    p.bias = b.data_ptr<float>(); }
  auto opts = w_img.options();
  auto out = at::empty({M, n_out}, opts.dtype(out_bf16 ? at::kBFloat16 : at::kFloat));
  at::Tensor a_save;
  p.a_save = nullptr;
  if (save_a) {
    a_save = at::empty({M, (int64_t)k_total}, opts.dtype(at::kBFloat16));
    p.a_save = reinterpret_cast<__nv_bfloat16*>(a_save.data_ptr());
  }
  p.w_img = w_img.data_ptr();
  p.out = out.data_ptr();
  p.out_stride = n_out;
  p.N = (int)N; p.n_out = (int)n_out; p.relu = relu ? 1 : 0; p.out_bf16 = out_bf16 ? 1 : 0;
  p.tmem_cols = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  if (M == 0) return {out, save_a ? a_save : at::Tensor()};
  unsigned grid = (unsigned)((M + kTileM - 1) / kTileM);
  auto stream = at::cuda::getCurrentCUDAStream();
  const int cpl = p.kp_nbr <= 128 ? 1 : p.kp_nbr / 128;
#define LAUNCH(C)                                                                                 \
  do {                                                                                            \
    C10_CUDA_CHECK(cudaFuncSetAttribute(sage_fused_fwd_kernel<C>,                                 \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sage_fused_fwd_kernel<C><<<grid, kThreads, smem, stream>>>(p);                                \
  } while (0)
  if (cpl == 1) LAUNCH(1); else if (cpl == 2) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {out, save_a ? a_save : at::Tensor()};
}

}  // namespace glb
