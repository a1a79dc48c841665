// Backward of the fused GraphSAGE layer on the 5th-gen tensor cores (replaces the library GEMMs of round 1).
//
//   dW_l [n_out, K_total] += dZ_l^T . A_l          (reduction over the 10^4..10^5 destination rows of the layer)
//
// Both operands are row-major with the REDUCTION index outermost, i.e. exactly the "MN-major" operand form
// of tcgen05.mma: a [64 rows x 64 cols] bf16 slab is one SWIZZLE_128B atom column (K = row index), so the
// tiles are staged with plain coalesced 16-byte loads + swizzled 16-byte shared stores - no transpose.
//
// Split-K over the whole chip: CTA (block, split) owns a [128 x <=256] block of dW and a contiguous range
// of rows; it accumulates in TMEM (fp32) over its rows and adds the partial block to the flat gradient
// buffer with vectorised red.global.add (the gradient buffer is zeroed once per step).
//
// The dZ operand is usually not read but COMPUTED on the fly (fused "backward through ReLU + mean
// aggregation", the former sage_bwd_input kernel):
//     dZ_l[r, :] = relu'(H_l[r, :]) * ( dA_{l+1}[self row of r, 0:d]  +  scale * dA_{l+1}[parent of r, kp_self : kp_self + d] )
// so dZ of the big hop segments never exists in HBM for 2-layer models; the bias gradient (column sums of
// dZ) falls out of the same registers.  Deeper models can ask for dZ to be written (needed by dA = dZ.W).
//
// warps 0..15 : producers (loads -> ReLU-grad math -> swizzled st.shared), later the epilogue (tcgen05.ld -> red);
//               warp 0 also owns TMEM and issues the MMAs of a stage once all 16 warps have filled it
//               (3-stage mbarrier ring, operands of stage st+1 already in flight in registers)
//
// Reference semantics: plain autograd of EgoSAGEConv (graphlearn/python/nn/tf/layers/ego_sage_conv.py:71-106).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cstring>
#include "host_utils.h"
#include "umma.cuh"

namespace glb {

constexpr int kDwProdWarps = 16;
constexpr int kDwProdThreads = kDwProdWarps * 32;     // 512
constexpr int kDwThreads = kDwProdThreads;            // warp 0 doubles as the MMA issuer (17 warps would round up to 20
                                                      // in the register allocator and cap the kernel at 96 registers)
constexpr int kDwStages = 3;
constexpr int kDwRows = 64;                           // reduction rows per stage (4 MMA K-steps)
constexpr int kDwZBytes = kDwRows * 128 * 2;          // dZ slab  [64 rows x 128 cols] bf16 = 16 KB
constexpr int kDwBBytes = kDwRows * 256 * 2;          // A slab   [64 rows x 256 cols] bf16 = 32 KB
constexpr int kDwStageBytes = kDwZBytes + kDwBBytes;
constexpr int kDwMaxSegs = 4;

struct BwdSeg {
  int row0, row1;                  // rows [row0, row1) of this layer's activation buffer
  const __nv_bfloat16* self_src;   // dA_{l+1} rows aligned with row0 (row r -> self_src + (r - row0) * ld_da), or null
  const __nv_bfloat16* nbr_src;    // dA_{l+1} rows of the parents (row r -> nbr_src + ((r - row0) / k) * ld_da + nbr_col0), or null
  int k;
  float scale;
};

struct DwParams {
  const __nv_bfloat16* dz; int ld_dz;       // dense dZ (top layer) or null -> computed from the fields below
  const __nv_bfloat16* h; int ld_h;         // activations of this layer (ReLU mask) or null (no mask)
  BwdSeg seg[kDwMaxSegs]; int nseg;
  int ld_da, nbr_col0;
  __nv_bfloat16* dz_out; int ld_dz_out;     // optional materialised dZ
  const __nv_bfloat16* a; int ld_a;         // saved [self || agg] rows, [rows, n_cols]
  float* dw; int ld_dw;                     // fp32 [n_out, n_cols], accumulated
  float* dbias;                             // optional fp32 [n_out], accumulated
  int rows, n_out, dz_cols, n_cols;
  int m_blocks, n_blocks, rows_per_split;
  uint32_t lbo, sbo, kadv;                  // MN-major descriptor fields (bytes)
  int skip_red;                             // debug: skip the red.global epilogue
};

__device__ __forceinline__ uint4 ldg16(const void* p) { return *reinterpret_cast<const uint4*>(p); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// byte offset of the 16-byte chunk (row r in [0,64), 8-column chunk c) inside an MN-major SW128 slab whose
// 64-column groups are 8 KB apart
__device__ __forceinline__ uint32_t mn_off(int r, int c) {
  return (uint32_t)(c >> 3) * 8192u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
}

__global__ void __launch_bounds__(kDwThreads, 1) sage_bwd_dw_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  float* colsum = reinterpret_cast<float*>(stages + kDwStages * kDwStageBytes);   // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(colsum + 128);
  uint64_t* bar_full = bars;                    // [kDwStages] producers -> MMA
  uint64_t* bar_empty = bars + kDwStages;       // [kDwStages] MMA -> producers
  uint64_t* bar_done = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDwStages + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = blockIdx.x % (p.m_blocks * p.n_blocks);
  const int split = blockIdx.x / (p.m_blocks * p.n_blocks);
  const int mb = blk % p.m_blocks, nb = blk / p.m_blocks;
  const int r_begin = split * p.rows_per_split;
  const int r_end = min(p.rows, r_begin + p.rows_per_split);
  const int n_steps = (r_end - r_begin + kDwRows - 1) / kDwRows;
  const int ncols_blk = min(256, p.n_cols - nb * 256);          // multiple of 16

  if (tid == 0) {
    for (int i = 0; i < kDwStages; ++i) { umma::mbar_init(bar_full + i, kDwProdWarps); umma::mbar_init(bar_empty + i, 1); }
    umma::mbar_init(bar_done, 1);
    umma::fence_barrier_init();
  }
  if (tid < 128) colsum[tid] = 0.f;
  if (warp == 0) { umma::tmem_alloc(tmem_slot, 256); umma::tmem_relinquish(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = umma::make_idesc_bf16_major(128, ncols_blk, 1, 1);
  {
    // ------------------------------------------------------------------ producers
    const int pt = tid;                              // 0..511
    const int zc = pt & 15, zr = pt >> 4;            // dZ slab: 16 chunks per row, rows zr and zr + 32
    const int bc = pt & 31, br = pt >> 5;            // A slab : 32 chunks per row, rows br + 16 i
    const int zcol = mb * 128 + zc * 8;              // first dZ column of this thread's chunk
    const int bcol = nb * 256 + bc * 8;
    const bool zcol_ok = zcol < p.dz_cols;
    const bool bcol_ok = bcol < p.n_cols;
    const bool want_out = p.dz_out != nullptr && nb == 0;
    float cs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = 0.f;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    // register-staged operands of one 64-row stage; two sets are alive so that the loads of stage st+1 are in
    // flight while stage st is converted and stored (the kernel is bound by bytes in flight, not by math)
    struct StageRegs { uint4 bv[4], hv[2], sv[2], nv[2]; float nscale[2]; };
    auto issue = [&](int st, StageRegs& g) {
      const int rbase = r_begin + st * kDwRows;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rbase + br + 16 * i;
        g.bv[i] = (r < r_end && bcol_ok) ? ldg16(p.a + (size_t)r * p.ld_a + bcol) : zero4;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = rbase + zr + 32 * i;
        g.hv[i] = zero4; g.sv[i] = zero4; g.nv[i] = zero4; g.nscale[i] = 0.f;
        if (r < r_end && zcol_ok) {
          if (p.dz) {
            g.sv[i] = ldg16(p.dz + (size_t)r * p.ld_dz + zcol);
          } else {
            int sg = 0;
#pragma unroll
            for (int j = 1; j < kDwMaxSegs; ++j) if (j < p.nseg && r >= p.seg[j].row0) sg = j;
            const BwdSeg& q = p.seg[sg];
            const int rl = r - q.row0;
            if (q.self_src) g.sv[i] = ldg16(q.self_src + (size_t)rl * p.ld_da + zcol);
            if (q.nbr_src) { g.nv[i] = ldg16(q.nbr_src + (size_t)(rl / q.k) * p.ld_da + p.nbr_col0 + zcol); g.nscale[i] = q.scale; }
            if (p.h) g.hv[i] = ldg16(p.h + (size_t)r * p.ld_h + zcol);
          }
        }
      }
    };
    auto process = [&](int st, StageRegs& g) {
      const int s = st % kDwStages;
      const int rbase = r_begin + st * kDwRows;
      // math: dZ = relu'(h) * (self + scale * nbr)
      uint4 zv[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (p.dz) { zv[i] = g.sv[i]; continue; }
        const uint32_t* ps = reinterpret_cast<const uint32_t*>(&g.sv[i]);
        const uint32_t* pn = reinterpret_cast<const uint32_t*>(&g.nv[i]);
        const uint32_t* ph = reinterpret_cast<const uint32_t*>(&g.hv[i]);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = unpack_bf16x2(ps[j]), b = unpack_bf16x2(pn[j]);
          float gx = a.x + b.x * g.nscale[i], gy = a.y + b.y * g.nscale[i];
          if (p.h) {
            const float2 hh = unpack_bf16x2(ph[j]);
            if (!(hh.x > 0.f)) gx = 0.f;
            if (!(hh.y > 0.f)) gy = 0.f;
          }
          o[j] = pack_bf16x2(gx, gy);
          cs[2 * j] += gx; cs[2 * j + 1] += gy;
        }
        zv[i] = make_uint4(o[0], o[1], o[2], o[3]);
        const int r = rbase + zr + 32 * i;
        if (want_out && r < r_end && zcol < p.ld_dz_out) *reinterpret_cast<uint4*>(p.dz_out + (size_t)r * p.ld_dz_out + zcol) = zv[i];
      }
      umma::mbar_wait(bar_empty + s, (uint32_t)(((st / kDwStages) & 1) ^ 1));
      uint8_t* zs = stages + (size_t)s * kDwStageBytes;
      uint8_t* bs = zs + kDwZBytes;
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<uint4*>(zs + mn_off(zr + 32 * i, zc)) = zv[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(bs + mn_off(br + 16 * i, bc)) = g.bv[i];
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bar_full + s);
      if (warp == 0) {
        // MMA issue for this stage (the operands of the next stage are already in flight in this warp's registers)
        umma::mbar_wait(bar_full + s, (uint32_t)((st / kDwStages) & 1));
        umma::tc_fence_after();
        if (lane == 0) {
          const uint32_t zb = umma::smem_u32(zs);
          const uint32_t bb = zb + kDwZBytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma::mma_bf16_ss(tmem_base, umma::make_desc_sw128_mn(zb + k4 * p.kadv, p.lbo, p.sbo),
                              umma::make_desc_sw128_mn(bb + k4 * p.kadv, p.lbo, p.sbo), idesc, (st | k4) ? 1u : 0u);
          umma::mma_commit(bar_empty + s);
          if (st == n_steps - 1) umma::mma_commit(bar_done);
        }
        __syncwarp();
      }
    };
    StageRegs ra, rb;
    if (n_steps > 0) issue(0, ra);
    for (int st = 0; st < n_steps; st += 2) {
      if (st + 1 < n_steps) issue(st + 1, rb);
      process(st, ra);
      if (st + 1 < n_steps) {
        if (st + 2 < n_steps) issue(st + 2, ra);
        process(st + 1, rb);
      }
    }
    // bias gradient = column sums of the computed dZ (only the n-block-0 CTAs of an m-block contribute): the two
    // rows of a warp are folded with a shuffle, every warp parks its 128 sums in the (now idle) stage memory and
    // 128 threads add the 16 rows up - no shared-memory atomics (they were 32-way same-address conflicts)
    const bool do_bias = p.dbias && !p.dz && nb == 0 && n_steps > 0;
    if (do_bias) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
      // all MMAs have been issued and committed by the time the last stage was filled; wait for them before the
      // stage memory is reused
      umma::mbar_wait(bar_done, 0);
      float* wsum = reinterpret_cast<float*>(stages) + warp * 128;
      if (lane < 16) {
        *reinterpret_cast<float4*>(wsum + zc * 8) = make_float4(cs[0], cs[1], cs[2], cs[3]);
        *reinterpret_cast<float4*>(wsum + zc * 8 + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
      }
    }
  }
  __syncthreads();
  if (p.dbias && !p.dz && nb == 0 && tid < 128 && mb * 128 + tid < p.n_out && n_steps > 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDwProdWarps; ++w) v += reinterpret_cast<const float*>(stages)[w * 128 + tid];
    atomicAdd(p.dbias + mb * 128 + tid, v);
  }
  // ------------------------------------------------------------------ epilogue: TMEM -> red.global.add into dW
  if (n_steps > 0) {
    umma::mbar_wait(bar_done, 0);
    umma::tc_fence_after();
    const int q = warp & 3;                       // TMEM lane quarter
    const int cq = warp >> 2;                     // column quarter (64 columns each)
    const int m = mb * 128 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = cq * 64; c0 < cq * 64 + 64; c0 += 32) {
      if (c0 >= ncols_blk) break;                 // warp-uniform
      uint32_t v[32];
      umma::tmem_ld32(taddr + (uint32_t)c0, v);
      umma::tmem_ld_wait();
      if (m < p.n_out && !p.skip_red) {
        float* dst = p.dw + (size_t)m * p.ld_dw + nb * 256 + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (c0 + 4 * i < ncols_blk)
            red_add_v4(dst + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                       __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 256);
}

// fp32 master weight [n_out, k_total] -> bf16 K-major SW128 images of W^T for the dA = dZ . W GEMM:
// image h (one per 256-column block of k_total) = (K_pad/64) k-blocks of [rows = 256 (columns 256h.. of W), 64 k (= n_out index)]
__global__ void pack_wt_sw128_kernel(const float* __restrict__ w, int n_out, int k_total, int kpad, int nrows_img,
                                     uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte chunk (8 k values) of one image row
  const int nkb = kpad >> 6;
  const int n_imgs = (k_total + nrows_img - 1) / nrows_img;
  const int total = n_imgs * nkb * nrows_img * 8;
  if (idx >= total) return;
  const int c = idx & 7;
  const int n = (idx >> 3) % nrows_img;
  const int kb = ((idx >> 3) / nrows_img) % nkb;
  const int h = (idx >> 3) / nrows_img / nkb;
  const int col = h * nrows_img + n;                          // column of W = row of W^T
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = kb * 64 + c * 8 + i;                       // row of W
    f[i] = (col < k_total && kk < n_out) ? w[(size_t)kk * k_total + col] : 0.f;
  }
  uint4 val;
  val.x = pack_bf16x2(f[0], f[1]); val.y = pack_bf16x2(f[2], f[3]); val.z = pack_bf16x2(f[4], f[5]); val.w = pack_bf16x2(f[6], f[7]);
  const size_t off = ((size_t)h * nkb + kb) * nrows_img * 128 + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)((c ^ (n & 7)) * 16);
  *reinterpret_cast<uint4*>(img + off) = val;
}

// returns the concatenated images [n_imgs, kpad * nrows_img] (bf16)
at::Tensor pack_weight_t(const at::Tensor& w, int64_t kpad, int64_t nrows_img) {
  TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.dim() == 2 && w.is_contiguous());
  TORCH_CHECK(kpad % 64 == 0 && kpad >= w.size(0) && nrows_img % 8 == 0);
  c10::cuda::CUDAGuard guard(w.device());
  const int64_t n_out = w.size(0), k_total = w.size(1);
  const int64_t n_imgs = (k_total + nrows_img - 1) / nrows_img;
  auto img = at::empty({n_imgs, kpad * nrows_img}, w.options().dtype(at::kBFloat16));
  const int total = (int)(n_imgs * (kpad / 64) * nrows_img * 8);
  pack_wt_sw128_kernel<<<(total + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      w.data_ptr<float>(), (int)n_out, (int)k_total, (int)kpad, (int)nrows_img, reinterpret_cast<uint8_t*>(img.data_ptr()));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return img;
}

static const __nv_bfloat16* bf16_ptr(const at::Tensor& t) { return reinterpret_cast<const __nv_bfloat16*>(t.data_ptr()); }

// dW (+)= dZ^T A with dZ either dense (`dz`) or computed from (h, segments of dA_next).
//   segs: rows [row0, row1) of the layer; self_src[i] / nbr_src[i] are views into dA_next (or None)
void sage_bwd_dw(const c10::optional<at::Tensor>& dz, const c10::optional<at::Tensor>& h,
                 const std::vector<int64_t>& seg_row0, const std::vector<int64_t>& seg_row1,
                 const std::vector<c10::optional<at::Tensor>>& self_src, const std::vector<c10::optional<at::Tensor>>& nbr_src,
                 const std::vector<int64_t>& seg_k, const std::vector<double>& seg_scale, int64_t nbr_col0,
                 const c10::optional<at::Tensor>& dz_out, const at::Tensor& a, const at::Tensor& dw,
                 const c10::optional<at::Tensor>& dbias, int64_t n_out, const std::vector<int64_t>& desc_override) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && a.dim() == 2 && a.stride(1) == 1 && a.stride(0) % 8 == 0);
  TORCH_CHECK(dw.is_cuda() && dw.scalar_type() == at::kFloat && dw.dim() == 2 && dw.stride(1) == 1 && dw.size(0) >= n_out &&
              dw.size(1) == a.size(1) && dw.stride(0) % 4 == 0 && (reinterpret_cast<uintptr_t>(dw.data_ptr()) & 15) == 0,
              "dw must be fp32 [n_out, K_total], 16-byte aligned rows");
  c10::cuda::CUDAGuard guard(a.device());
  DwParams p;
  std::memset(&p, 0, sizeof(p));
  p.rows = (int)a.size(0); p.n_cols = (int)a.size(1); p.n_out = (int)n_out;
  TORCH_CHECK(p.n_cols % 16 == 0, "K_total must be a multiple of 16");
  p.a = bf16_ptr(a); p.ld_a = (int)a.stride(0);
  p.dw = dw.data_ptr<float>(); p.ld_dw = (int)dw.stride(0);
  if (dbias.has_value() && dbias->defined()) {
    TORCH_CHECK(dbias->scalar_type() == at::kFloat && dbias->numel() >= n_out);
    p.dbias = dbias->data_ptr<float>();
  }
  if (dz.has_value() && dz->defined()) {
    TORCH_CHECK(dz->scalar_type() == at::kBFloat16 && dz->dim() == 2 && dz->stride(1) == 1 && dz->size(0) == p.rows &&
                dz->stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(dz->data_ptr()) & 15) == 0);
    p.dz = bf16_ptr(*dz); p.ld_dz = (int)dz->stride(0);
    p.dz_cols = (int)(dz->stride(0) >= (n_out + 7) / 8 * 8 ? (n_out + 7) / 8 * 8 : dz->size(1) / 8 * 8);
    TORCH_CHECK(p.dz_cols >= n_out || dz->size(1) % 8 == 0, "dense dZ rows must be readable in 16-byte chunks");
  } else {
    const int nseg = (int)seg_row0.size();
    TORCH_CHECK(nseg >= 1 && nseg <= kDwMaxSegs && (int)seg_row1.size() == nseg && (int)self_src.size() == nseg &&
                (int)nbr_src.size() == nseg && (int)seg_k.size() == nseg && (int)seg_scale.size() == nseg);
    p.nseg = nseg;
    p.nbr_col0 = (int)nbr_col0;
    TORCH_CHECK(n_out % 8 == 0 && nbr_col0 % 8 == 0, "hidden width must be a multiple of 8");
    p.dz_cols = (int)n_out;
    for (int i = 0; i < nseg; ++i) {
      BwdSeg& g = p.seg[i];
      g.row0 = (int)seg_row0[i]; g.row1 = (int)seg_row1[i]; g.k = (int)std::max<int64_t>(seg_k[i], 1); g.scale = (float)seg_scale[i];
      for (int which = 0; which < 2; ++which) {
        const auto& src = which == 0 ? self_src[i] : nbr_src[i];
        if (!(src.has_value() && src->defined())) continue;
        TORCH_CHECK(src->scalar_type() == at::kBFloat16 && src->dim() == 2 && src->stride(1) == 1 && src->stride(0) % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(src->data_ptr()) & 15) == 0);
        TORCH_CHECK(p.ld_da == 0 || p.ld_da == src->stride(0), "all dA views must share the leading dimension");
        p.ld_da = (int)src->stride(0);
        if (which == 0) { TORCH_CHECK(src->size(0) >= g.row1 - g.row0); g.self_src = bf16_ptr(*src); }
        else { TORCH_CHECK(src->size(0) * g.k >= g.row1 - g.row0); g.nbr_src = bf16_ptr(*src); }
      }
    }
    if (h.has_value() && h->defined()) {
      TORCH_CHECK(h->scalar_type() == at::kBFloat16 && h->dim() == 2 && h->stride(1) == 1 && h->size(0) >= p.rows &&
                  h->size(1) >= n_out && h->stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(h->data_ptr()) & 15) == 0);
      p.h = bf16_ptr(*h); p.ld_h = (int)h->stride(0);
    }
    if (dz_out.has_value() && dz_out->defined()) {
      TORCH_CHECK(dz_out->scalar_type() == at::kBFloat16 && dz_out->dim() == 2 && dz_out->stride(1) == 1 && dz_out->size(0) >= p.rows &&
                  dz_out->stride(0) % 8 == 0 && dz_out->size(1) >= n_out);
      p.dz_out = reinterpret_cast<__nv_bfloat16*>(dz_out->data_ptr()); p.ld_dz_out = (int)dz_out->stride(0);
    }
  }
  if (p.rows == 0) return;
  p.m_blocks = (p.n_out + 127) / 128;
  p.n_blocks = (p.n_cols + 255) / 256;
  const int blocks = p.m_blocks * p.n_blocks;
  const int sms = sm_count();
  const int max_splits = std::max(1, sms / blocks);
  int rps = (p.rows + max_splits - 1) / max_splits;
  rps = std::max(kDwRows, (rps + kDwRows - 1) / kDwRows * kDwRows);
  p.rows_per_split = rps;
  const int splits = (p.rows + rps - 1) / rps;
  p.lbo = 8192; p.sbo = 1024; p.kadv = 2048;
  if (desc_override.size() == 4) p.skip_red = (int)desc_override[3];
  if (desc_override.size() >= 3) { p.lbo = (uint32_t)desc_override[0]; p.sbo = (uint32_t)desc_override[1]; p.kadv = (uint32_t)desc_override[2]; }
  const size_t smem = 1024 + (size_t)kDwStages * kDwStageBytes + 512 + (2 * kDwStages + 2) * 8;
  static bool attr_done = false;
  if (!attr_done) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(sage_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  sage_bwd_dw_kernel<<<blocks * splits, kDwThreads, smem, at::cuda::getCurrentCUDAStream()>>>(p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace glb
