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
#include <cstring>
#include "host_utils.h"
#include "umma.cuh"

namespace glb {

constexpr int kTileM = 128;
constexpr int kThreads = 1024;         // 32 warps: the gather needs warps in flight, not registers
constexpr int kEpiWarps = 16;          // warps that drain TMEM in the epilogue
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
  int rows_per_cta;           // destination rows gathered by one CTA (<= 128)
  const char* zero_row;       // >= 2 KB of zeros: target of the loads of masked lanes / missing rows
  int wshift_self, wshift_nbr; // log2(world) of each table when it is a power of two, else -1
  long long* debug_ts;        // optional [grid][16] phase timestamps (clock64) written by thread 0
  const __nv_bfloat16* a_src; // optional precomputed A [M, K_total] (row-major): skip the gather, just stage it
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

// One 16-byte storage chunk of a feature row: 4 fp32 or 8 bf16 features.  Kept RAW
// (unconverted) so that a whole batch of row loads is issued back to back before the first use.
template <int DT> struct Chunk;
template <> struct Chunk<0> {
  static constexpr int kVec = 4;
  float4 v;
  __device__ __forceinline__ void load(const char* p) { v = ld_nc_f4(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void add_to(float (&a)[4]) const { a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w; }
};
template <> struct Chunk<1> {
  static constexpr int kVec = 8;
  uint4 v;
  __device__ __forceinline__ void load(const char* p) { v = ld_nc_u4(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void add_to(float (&a)[8]) const {
    float2 x;
    x = unpack_bf16x2(v.x); a[0] += x.x; a[1] += x.y;
    x = unpack_bf16x2(v.y); a[2] += x.x; a[3] += x.y;
    x = unpack_bf16x2(v.z); a[4] += x.x; a[5] += x.y;
    x = unpack_bf16x2(v.w); a[6] += x.x; a[7] += x.y;
  }
};

// locator of a table row packed into 32 bits: (row << 3) | owner, 0xFFFFFFFF = missing
__device__ __forceinline__ uint32_t make_loc(const TableView& t, int64_t vid, int wshift) {
  if (vid < 0) return 0xFFFFFFFFu;
  int owner; int64_t row;
  if (wshift >= 0) { owner = (int)(vid & ((1 << wshift) - 1)); row = vid >> wshift; }
  else { owner = (int)(vid % t.world); row = vid / t.world; }
  if (row >= t.nrows[owner]) return 0xFFFFFFFFu;
  return ((uint32_t)row << 3) | (uint32_t)owner;
}
__device__ __forceinline__ const char* loc_ptr(const TableView& t, uint32_t loc, uint32_t row_bytes) {
  return reinterpret_cast<const char*>(t.base.p[loc & 7u]) + (size_t)(loc >> 3) * row_bytes;
}

// row address of a vid: the local replica-cache copy when the row is remote and cached (N17),
// the owner's HBM otherwise; `zero_row` for missing / padded ids
__device__ __forceinline__ const char* vid_ptr(const TableView& t, int64_t vid, int wshift, uint32_t row_bytes,
                                               const char* zero_row) {
  const uint32_t loc = make_loc(t, vid, wshift);
  if (loc == 0xFFFFFFFFu) return zero_row;
  if (t.cmap != nullptr && (int)(loc & 7u) != t.self) {
    const int s = __ldg(t.cmap + vid);
    if (s >= 0) return t.cbase + (size_t)s * row_bytes;
  }
  return loc_ptr(t, loc, row_bytes);
}

// write VEC consecutive K elements of tile row r starting at K column kcol (multiple of VEC)
template <int VEC>
__device__ __forceinline__ void put_chunk(uint8_t* sA, __nv_bfloat16* a_save, size_t a_off, int r, int kcol,
                                          const float (&v)[VEC]) {
  const int kb = kcol >> 6;
  const uint32_t off = (uint32_t)kb * (kTileM * 128) + umma::sw128_offset((uint32_t)r, (uint32_t)(kcol & 63));
  if constexpr (VEC == 4) {
    uint2 u; u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(sA + off) = u;
    if (a_save) *reinterpret_cast<uint2*>(a_save + a_off + kcol) = u;
  } else {
    uint4 u; u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(sA + off) = u;
    if (a_save) *reinterpret_cast<uint4*>(a_save + a_off + kcol) = u;
  }
}

// epilogue shared by both kernel variants: TMEM -> registers -> bias/ReLU -> global
__device__ __forceinline__ void sage_epilogue(const SageParams& p, uint32_t tmem_base, int m0, int R, int warp, int lane) {
  // --- epilogue: TMEM -> registers -> bias/ReLU -> global.  A warp may only touch TMEM lane quarter
  //     (warp % 4); the column range is split over 8 warp groups (N >= 128) or 4 (N = 64).
  const int epi_groups = (p.N % 128 == 0) ? 8 : 4;
  if (warp < 4 * epi_groups) {
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int g = warp >> 2;             // column group
    const int cols_per_group = p.N / epi_groups;   // multiple of 16
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = row < R && m < p.M;
    for (int c0 = 0; c0 < cols_per_group; c0 += 16) {
      const int n0 = g * cols_per_group + c0;
      uint32_t v[16];
      umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, v);
      umma::tmem_ld_wait();
      if (row_ok && n0 < p.n_out) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(v[i]);
          if (p.bias && n0 + i < p.n_out) x += __ldg(p.bias + n0 + i);
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
}

// U  = neighbour row loads kept in flight per lane (one batch)
// DT = storage dtype of the self AND neighbour tables (0 fp32, 1 bf16)
//
// Work decomposition: the CTA owns `rows_per_cta` consecutive destination rows (<= 128; a small M
// is spread over many CTAs - the unused rows of the 128-row MMA tile are never stored).  A row
// needs LPR = kp / VEC lanes (16-byte chunk per lane); when LPR < 32 a warp processes 32/LPR rows
// side by side (sub-warp groups), when LPR > 32 the row is cut into 32-lane slices.  Per item a
// lane group (1) already holds the neighbour locators (prefetched during the previous item),
// (2) issues ALL self + neighbour chunk loads of the batch unconditionally (masked lanes read a
// zero row), (3) reduces in fp32, (4) writes bf16 into the SW128 A tile.
constexpr int kMaxSlots = 64;
constexpr int kProducers = 4;

#define GLB_TS(i) do { if (p.debug_ts && threadIdx.x == 0) p.debug_ts[(size_t)blockIdx.x * 16 + (i)] = clock64(); } while (0)

template <int U, int DT>
__global__ void __launch_bounds__(kThreads, 1) sage_fused_fwd_kernel(const SageParams p) {
  GLB_TS(0);
  // keep every pointer derived from the __shared__ symbol by plain integer offsets so that the
  // compiler emits LDS/STS (a uintptr_t round trip would demote them to generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
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
  GLB_TS(1);

  // --- weights: TMA bulk copies of the pre-swizzled image, overlapped with the gather
  if (tid == 0) {
    umma::mbar_arrive_expect_tx(bar_w, (uint32_t)nkb * w_kb_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w_img);
    for (int kb = 0; kb < nkb; ++kb)
      umma::bulk_g2s(sW + (size_t)kb * w_kb_bytes, src + (size_t)kb * w_kb_bytes, w_kb_bytes, bar_w);
  }

  const int R = p.rows_per_cta;
  const int m0 = blockIdx.x * R;
  if (p.a_src != nullptr) {
    // split path (multi-GPU): A was produced by gather_self_mean_kernel; stage the tile with coalesced
    // 16-byte loads -> swizzled 16-byte shared stores
    const int R_ = p.rows_per_cta;
    const int m0_ = blockIdx.x * R_;
    const int chunks_row = k_total >> 3;                 // 16-byte chunks per A row
    for (int i = tid; i < R_ * chunks_row; i += kThreads) {
      const int r = i / chunks_row, c = i - r * chunks_row;
      const int m = m0_ + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (m < p.M) v = ld_nc_u4(reinterpret_cast<const uint4*>(p.a_src + (size_t)m * k_total) + c);
      const int kcol = c * 8;
      *reinterpret_cast<uint4*>(sA + (size_t)(kcol >> 6) * (kTileM * 128) + umma::sw128_offset((uint32_t)r, (uint32_t)(kcol & 63))) = v;
    }
    GLB_TS(2);
  } else {
  // --- phase 0: resolve the tile's neighbour / self ids into ROW POINTERS staged in shared memory
  //     (coalesced read of nbr_vids[m0*k .. (m0+R)*k); missing rows point at a zero row).  The
  //     hot loop below is then just  LDS.64 + IADD + LDG  per row chunk.
  constexpr int VEC = Chunk<DT>::kVec;
  const int R = p.rows_per_cta;
  const int m0 = blockIdx.x * R;
  const int k = p.k;
  const char** sPtrN = reinterpret_cast<const char**>(bars + 4 + 2 * kMaxSlots);   // [R * k]
  const char** sPtrS = sPtrN + (size_t)R * k;                          // [R]
  const bool need_self = p.kp_self > 0 || p.mode == kGcnMean;
  {
    const int wshift_n = p.wshift_nbr, wshift_s = p.wshift_self;
    const uint32_t nbr_row_bytes = (uint32_t)p.tnbr.stride * (DT == 0 ? 4u : 2u);
    const uint32_t self_row_bytes = (uint32_t)p.tself.stride * (DT == 0 ? 4u : 2u);
    const int64_t base = (int64_t)m0 * k;
    const int64_t lim = (int64_t)p.M * k;
    for (int i = tid; i < R * k; i += kThreads) {
      const int64_t idx = base + i;
      sPtrN[i] = idx < lim ? vid_ptr(p.tnbr, p.nbr_vids ? __ldg(p.nbr_vids + idx) : idx, wshift_n, nbr_row_bytes, p.zero_row) : p.zero_row;
    }
    for (int i = tid; i < R; i += kThreads) {
      const int m = m0 + i;
      sPtrS[i] = (need_self && m < p.M) ? vid_ptr(p.tself, p.self_vids ? __ldg(p.self_vids + m) : (int64_t)m, wshift_s, self_row_bytes, p.zero_row) : p.zero_row;
    }
  }
  __syncthreads();
  GLB_TS(2);

  // --- phase 1: gather + aggregate -> A tile (bf16, SW128 K-major)
  const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
  const int lanes_row = p.kp_nbr / VEC;                 // 16-byte chunks per (half) row: 8..128
  const int lpr = lanes_row < 32 ? lanes_row : 32;      // lanes per row inside a warp
  const int lshift = 31 - __clz(lpr);                   // log2(lpr)
  const int rpi = 32 >> lshift;                         // rows per item
  const int n_slices = lanes_row > 32 ? lanes_row >> 5 : 1;
  const int row_groups = (R + rpi - 1) / rpi;
  const int n_items = row_groups * n_slices;
  const int sub = lane >> lshift;                       // which row of the item this lane works on
  const int lig = lane & (lpr - 1);                     // lane index inside its row group
  const bool has_self = p.kp_self > 0;
  float scale = 1.f;
  if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
  else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);

  for (int item = warp; item < n_items; item += kWarps) {
    const int rg = n_slices == 1 ? item : item / n_slices;
    const int sl = n_slices == 1 ? 0 : item - rg * n_slices;
    const int r = rg * rpi + sub;
    if (r >= R) continue;                                // (only when R is not a multiple of rpi)
    const int m = m0 + r;
    const int chunk = lig + 32 * sl;                     // this lane's 16-byte chunk of the row
    const int f0 = chunk * VEC;                          // first feature of the chunk
    const size_t coff = (size_t)chunk * 16;
    float acc[VEC], sv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { acc[i] = 0.f; sv[i] = 0.f; }
    // lanes whose chunk lies beyond the real feature width skip the loads altogether (their A
    // columns are the zero K-padding); the others issue self + U neighbour loads back to back
    if (f0 < d_nbr) {
      const char* const* ptrs = sPtrN + (size_t)r * k;
      Chunk<DT> sraw;
      const bool self_ld = need_self && f0 < d_self;
      if (self_ld) sraw.load(sPtrS[r] + coff);
      for (int j0 = 0; j0 < k; j0 += U) {
        Chunk<DT> raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u < k ? j0 + u : k - 1;      // tail slots re-read the last row (masked below)
          raw[u].load(ptrs[j] + coff);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (j0 + u < k) raw[u].add_to(acc);
      }
      if (self_ld) sraw.add_to(sv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {                    // tail masks (dims that are not multiples of VEC)
        if (f0 + i >= d_self) sv[i] = 0.f;
        if (f0 + i >= d_nbr) acc[i] = 0.f;
      }
    } else if (need_self && f0 < d_self) {               // self wider than the neighbour half (rare)
      Chunk<DT> sraw;
      sraw.load(sPtrS[r] + coff);
      sraw.add_to(sv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) if (f0 + i >= d_self) sv[i] = 0.f;
    }
    const size_t a_off = (size_t)m * k_total;
    __nv_bfloat16* asave = (p.a_save && m < p.M) ? p.a_save : nullptr;
    if (has_self) put_chunk<VEC>(sA, asave, a_off, r, f0, sv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = (acc[i] + (p.mode == kGcnMean ? sv[i] : 0.f)) * scale;
    put_chunk<VEC>(sA, asave, a_off, r, p.kp_self + f0, acc);
  }
  }  // gather vs. dense-A
  GLB_TS(3);
  umma::fence_proxy_async_smem();     // generic-proxy st.shared -> visible to tcgen05 (async proxy)
  __syncthreads();
  GLB_TS(4);

  // --- GEMM: one thread issues all MMAs; accumulator lives in TMEM
  if (tid == 0) {
    umma::mbar_wait(bar_w, 0);
    GLB_TS(5);
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
    GLB_TS(6);
  }
  __syncwarp();
  umma::mbar_wait(bar_mma, 0);
  GLB_TS(7);
  umma::tc_fence_after();

  sage_epilogue(p, tmem_base, m0, R, warp, lane);
  GLB_TS(8);
  umma::tc_fence_before();
  __syncthreads();
  GLB_TS(9);
  if (warp == 1) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// TMA-gather variant.  Same math and same tile / MMA / epilogue as above, but the feature rows are
// pulled by the TMA engine (cp.async.bulk, one bulk copy per row, completion on mbarriers) into a
// shared-memory ring instead of through registers: ~128 KB of row fetches stay in flight per SM
// with zero register cost, which is what hides HBM *and* NVLink latency (peer rows cost ~2 us).
// The ring lives in the weight region: W is only needed after the gather, so it is fetched (one
// more bulk copy, from L2) once the ring has drained.
//   warp 0            producer: per destination row issues (1 + k) row copies into the next slot
//   warps 1..31       consumers: wait for a slot, reduce the k neighbour rows (fp32), write the bf16
//                     A-tile chunks (+ the row-major copy saved for backward), release the slot
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kThreads, 1) sage_fused_tma_kernel(const SageParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  GLB_TS(0);
  const int k_total = p.kp_self + p.kp_nbr;
  const int nkb = k_total >> 6;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (size_t)nkb * (kTileM * 128);                     // ring during the gather, W afterwards
  const uint32_t w_kb_bytes = (uint32_t)p.N * 128u;
  const uint32_t w_bytes = (uint32_t)nkb * w_kb_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)w_bytes);
  uint64_t* bar_w = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  uint64_t* full = bars + 4;                                           // [kMaxSlots]
  uint64_t* empty = full + kMaxSlots;                                  // [kMaxSlots]
  const char** sPtrN = reinterpret_cast<const char**>(empty + kMaxSlots);
  const int R = p.rows_per_cta;
  const int k = p.k;
  const char** sPtrS = sPtrN + (size_t)R * k;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  constexpr int VEC = Chunk<DT>::kVec;
  const bool need_self = p.kp_self > 0 || p.mode == kGcnMean;
  const uint32_t nbr_row_bytes = (uint32_t)p.tnbr.stride * (DT == 0 ? 4u : 2u);
  const uint32_t self_row_bytes = (uint32_t)p.tself.stride * (DT == 0 ? 4u : 2u);
  // bytes actually copied per row: the real features rounded up to 16 B (never more than the stride)
  const uint32_t nbr_copy = min(nbr_row_bytes, (uint32_t)((p.tnbr.dim * (DT == 0 ? 4 : 2) + 15) & ~15));
  const uint32_t self_copy = need_self ? min(self_row_bytes, (uint32_t)((p.tself.dim * (DT == 0 ? 4 : 2) + 15) & ~15)) : 0u;
  const uint32_t slot_bytes = self_copy + (uint32_t)k * nbr_copy;
  const int S = min((int)(w_bytes / slot_bytes), kMaxSlots);

  if (tid == 0) {
    umma::mbar_init(bar_w, 1);
    umma::mbar_init(bar_mma, 1);
    for (int s = 0; s < S; ++s) { umma::mbar_init(full + s, 1); umma::mbar_init(empty + s, 1); }
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    umma::tmem_relinquish();
  }
  // resolve ids -> row pointers (coalesced), as in the register variant
  const int m0 = blockIdx.x * R;
  {
    const int64_t base = (int64_t)m0 * k;
    const int64_t lim = (int64_t)p.M * k;
    for (int i = tid; i < R * k; i += kThreads) {
      const int64_t idx = base + i;
      sPtrN[i] = idx < lim ? vid_ptr(p.tnbr, p.nbr_vids ? __ldg(p.nbr_vids + idx) : idx, p.wshift_nbr, nbr_row_bytes, p.zero_row) : p.zero_row;
    }
    for (int i = tid; i < R; i += kThreads) {
      const int m = m0 + i;
      sPtrS[i] = (need_self && m < p.M) ? vid_ptr(p.tself, p.self_vids ? __ldg(p.self_vids + m) : (int64_t)m, p.wshift_self, self_row_bytes, p.zero_row) : p.zero_row;
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GLB_TS(2);

  const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
  const bool has_self = p.kp_self > 0;
  float scale = 1.f;
  if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
  else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);

  if (warp == 0) {
    // ===== producer: one slot per destination row =====
    for (int i = 0; i < R; ++i) {
      const int s = i % S, turn = i / S;
      if (turn > 0) umma::mbar_wait(empty + s, (uint32_t)((turn - 1) & 1));
      uint8_t* slot = sW + (size_t)s * slot_bytes;
      if (lane == 0) umma::mbar_arrive_expect_tx(full + s, slot_bytes);
      __syncwarp();
      if (need_self && lane == 0) umma::bulk_g2s(slot, sPtrS[i], self_copy, full + s);
      for (int j = lane; j < k; j += 32)
        umma::bulk_g2s(slot + self_copy + (size_t)j * nbr_copy, sPtrN[(size_t)i * k + j], nbr_copy, full + s);
    }
  } else {
    // ===== consumers =====
    // Slot s is always drained by the SAME consumer warp (s % C) in increasing turn order: an
    // mbarrier only distinguishes the parity of a phase, so two warps waiting for different turns
    // of one slot would alias.
    const int nchunks = p.kp_nbr / VEC;                  // 16-byte chunks per K half
    const int C = kWarps - 1, c = warp - 1;
    for (int turn = 0; turn * S < R; ++turn)
    for (int s = c; s < S; s += C) {
      const int i = turn * S + s;
      if (i >= R) break;
      umma::mbar_wait(full + s, (uint32_t)(turn & 1));
      const uint8_t* slot = sW + (size_t)s * slot_bytes;
      const int m = m0 + i;
      const size_t a_off = (size_t)m * k_total;
      __nv_bfloat16* asave = (p.a_save && m < p.M) ? p.a_save : nullptr;
      for (int chunk = lane; chunk < nchunks; chunk += 32) {
        const int f0 = chunk * VEC;
        float acc[VEC], sv[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) { acc[q] = 0.f; sv[q] = 0.f; }
        if (f0 < d_nbr) {
          const uint8_t* np = slot + self_copy + (size_t)chunk * 16;
          for (int j = 0; j < k; ++j) {
            Chunk<DT> c;
            c.v = *reinterpret_cast<const decltype(c.v)*>(np + (size_t)j * nbr_copy);
            c.add_to(acc);
          }
        }
        if (need_self && f0 < d_self) {
          Chunk<DT> c;
          c.v = *reinterpret_cast<const decltype(c.v)*>(slot + (size_t)chunk * 16);
          c.add_to(sv);
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          if (f0 + q >= d_self) sv[q] = 0.f;
          if (f0 + q >= d_nbr) acc[q] = 0.f;
        }
        if (has_self) put_chunk<VEC>(sA, asave, a_off, i, f0, sv);
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = (acc[q] + (p.mode == kGcnMean ? sv[q] : 0.f)) * scale;
        put_chunk<VEC>(sA, asave, a_off, i, p.kp_self + f0, acc);
      }
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(empty + s);
    }
  }
  GLB_TS(3);
  umma::fence_proxy_async_smem();     // A-tile st.shared visible to tcgen05; ring reads ordered before the W copy
  __syncthreads();
  GLB_TS(4);

  // --- weights into the (now idle) ring region, then the MMAs
  if (tid == 0) {
    umma::mbar_arrive_expect_tx(bar_w, w_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w_img);
    for (int kb = 0; kb < nkb; ++kb)
      umma::bulk_g2s(sW + (size_t)kb * w_kb_bytes, src + (size_t)kb * w_kb_bytes, w_kb_bytes, bar_w);
    umma::mbar_wait(bar_w, 0);
    GLB_TS(5);
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc_bf16(kTileM, p.N);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t a_base = umma::smem_u32(sA + (size_t)kb * (kTileM * 128));
      const uint32_t b_base = umma::smem_u32(sW + (size_t)kb * w_kb_bytes);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma::mma_bf16_ss(tmem_base, umma::make_desc_sw128(a_base + k4 * 32),
                          umma::make_desc_sw128(b_base + k4 * 32), idesc, (kb | k4) ? 1u : 0u);
    }
    umma::mma_commit(bar_mma);
    GLB_TS(6);
  }
  __syncwarp();
  umma::mbar_wait(bar_mma, 0);
  GLB_TS(7);
  umma::tc_fence_after();
  sage_epilogue(p, tmem_base, m0, R, warp, lane);
  GLB_TS(8);
  umma::tc_fence_before();
  __syncthreads();
  GLB_TS(9);
  if (warp == 1) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// cp.async-gather variant (the multi-GPU / multi-wave default).  Same ring-in-the-W-region structure
// as the TMA variant, but the rows are moved by per-lane 16-byte cp.async (LDGSTS): normal LSU issue
// rate (the TMA unit needs ~110 cycles per small bulk copy), zero registers held while in flight,
// and the ring keeps ~128 KB of row fetches outstanding per SM - enough to cover NVLink peer
// latency.  kProducers warps issue copies, the remaining warps reduce + write the A tile.
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kThreads, 1) sage_fused_async_kernel(const SageParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  GLB_TS(0);
  const int k_total = p.kp_self + p.kp_nbr;
  const int nkb = k_total >> 6;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (size_t)nkb * (kTileM * 128);                     // ring during the gather, W afterwards
  const uint32_t w_kb_bytes = (uint32_t)p.N * 128u;
  const uint32_t w_bytes = (uint32_t)nkb * w_kb_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)w_bytes);
  uint64_t* bar_w = bars;
  uint64_t* bar_mma = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  uint64_t* full = bars + 4;                                           // [kMaxSlots]
  uint64_t* empty = full + kMaxSlots;                                  // [kMaxSlots]
  const char** sPtrN = reinterpret_cast<const char**>(empty + kMaxSlots);
  const int R = p.rows_per_cta;
  const int k = p.k;
  const char** sPtrS = sPtrN + (size_t)R * k;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  constexpr int VEC = Chunk<DT>::kVec;
  const bool need_self = p.kp_self > 0 || p.mode == kGcnMean;
  const uint32_t nbr_row_bytes = (uint32_t)p.tnbr.stride * (DT == 0 ? 4u : 2u);
  const uint32_t self_row_bytes = (uint32_t)p.tself.stride * (DT == 0 ? 4u : 2u);
  // bytes actually copied per row: the real features rounded up to 16 B (never more than the stride)
  const uint32_t nbr_copy = min(nbr_row_bytes, (uint32_t)((p.tnbr.dim * (DT == 0 ? 4 : 2) + 15) & ~15));
  const uint32_t self_copy = need_self ? min(self_row_bytes, (uint32_t)((p.tself.dim * (DT == 0 ? 4 : 2) + 15) & ~15)) : 0u;
  const uint32_t slot_bytes = self_copy + (uint32_t)k * nbr_copy;
  constexpr int P = kProducers;                                         // producer warps
  const int S = (min((int)(w_bytes / slot_bytes), kMaxSlots) / P) * P;   // slot s belongs to producer s % P

  if (tid == 0) {
    umma::mbar_init(bar_w, 1);
    umma::mbar_init(bar_mma, 1);
    for (int s = 0; s < S; ++s) { umma::mbar_init(full + s, 32); umma::mbar_init(empty + s, 1); }
    umma::fence_barrier_init();
  }
  if (warp == 1) {
    umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    umma::tmem_relinquish();
  }
  // resolve ids -> row pointers (coalesced), as in the register variant
  const int m0 = blockIdx.x * R;
  {
    const int64_t base = (int64_t)m0 * k;
    const int64_t lim = (int64_t)p.M * k;
    for (int i = tid; i < R * k; i += kThreads) {
      const int64_t idx = base + i;
      sPtrN[i] = idx < lim ? vid_ptr(p.tnbr, p.nbr_vids ? __ldg(p.nbr_vids + idx) : idx, p.wshift_nbr, nbr_row_bytes, p.zero_row) : p.zero_row;
    }
    for (int i = tid; i < R; i += kThreads) {
      const int m = m0 + i;
      sPtrS[i] = (need_self && m < p.M) ? vid_ptr(p.tself, p.self_vids ? __ldg(p.self_vids + m) : (int64_t)m, p.wshift_self, self_row_bytes, p.zero_row) : p.zero_row;
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GLB_TS(2);

  const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
  const bool has_self = p.kp_self > 0;
  float scale = 1.f;
  if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
  else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);

  if (warp < P) {
    // ===== producers: warp p fills the slots s = p (mod P); every lane moves 16-byte chunks with
    //       cp.async (LDGSTS) and the slot's mbarrier fires when all 32 lanes' copies have landed =====
    const uint32_t cps = self_copy >> 4, cpn = nbr_copy >> 4;          // 16-byte chunks per row
    const uint32_t total = cps + (uint32_t)k * cpn;
    for (int turn = 0; turn * S < R; ++turn)
    for (int s = warp; s < S; s += P) {
      const int i = turn * S + s;
      if (i >= R) break;
      if (turn > 0) umma::mbar_wait(empty + s, (uint32_t)((turn - 1) & 1));
      uint8_t* slot = sW + (size_t)s * slot_bytes;
      const char* sp = sPtrS[i];
      const char* const* np = sPtrN + (size_t)i * k;
      for (uint32_t c = lane; c < total; c += 32) {
        const char* src;
        if (c < cps) src = sp + (size_t)c * 16;
        else { const uint32_t cc = c - cps; const uint32_t j = cc / cpn; src = np[j] + (size_t)(cc - j * cpn) * 16; }
        umma::cp_async16(slot + (size_t)c * 16, src);
      }
      umma::cp_async_mbar_arrive_noinc(full + s);
    }
  } else {
    // ===== consumers =====
    // Slot s is always drained by the SAME consumer warp (s % C) in increasing turn order: an
    // mbarrier only distinguishes the parity of a phase, so two warps waiting for different turns
    // of one slot would alias.
    const int nchunks = p.kp_nbr / VEC;                  // 16-byte chunks per K half
    const int C = kWarps - P, c = warp - P;
    for (int turn = 0; turn * S < R; ++turn)
    for (int s = c; s < S; s += C) {
      const int i = turn * S + s;
      if (i >= R) break;
      umma::mbar_wait(full + s, (uint32_t)(turn & 1));
      const uint8_t* slot = sW + (size_t)s * slot_bytes;
      const int m = m0 + i;
      const size_t a_off = (size_t)m * k_total;
      __nv_bfloat16* asave = (p.a_save && m < p.M) ? p.a_save : nullptr;
      for (int chunk = lane; chunk < nchunks; chunk += 32) {
        const int f0 = chunk * VEC;
        float acc[VEC], sv[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) { acc[q] = 0.f; sv[q] = 0.f; }
        if (f0 < d_nbr) {
          const uint8_t* np = slot + self_copy + (size_t)chunk * 16;
          for (int j = 0; j < k; ++j) {
            Chunk<DT> c;
            c.v = *reinterpret_cast<const decltype(c.v)*>(np + (size_t)j * nbr_copy);
            c.add_to(acc);
          }
        }
        if (need_self && f0 < d_self) {
          Chunk<DT> c;
          c.v = *reinterpret_cast<const decltype(c.v)*>(slot + (size_t)chunk * 16);
          c.add_to(sv);
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          if (f0 + q >= d_self) sv[q] = 0.f;
          if (f0 + q >= d_nbr) acc[q] = 0.f;
        }
        if (has_self) put_chunk<VEC>(sA, asave, a_off, i, f0, sv);
#pragma unroll
        for (int q = 0; q < VEC; ++q) acc[q] = (acc[q] + (p.mode == kGcnMean ? sv[q] : 0.f)) * scale;
        put_chunk<VEC>(sA, asave, a_off, i, p.kp_self + f0, acc);
      }
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(empty + s);
    }
  }
  GLB_TS(3);
  umma::fence_proxy_async_smem();     // A-tile st.shared visible to tcgen05; ring reads ordered before the W copy
  __syncthreads();
  GLB_TS(4);

  // --- weights into the (now idle) ring region, then the MMAs
  if (tid == 0) {
    umma::mbar_arrive_expect_tx(bar_w, w_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w_img);
    for (int kb = 0; kb < nkb; ++kb)
      umma::bulk_g2s(sW + (size_t)kb * w_kb_bytes, src + (size_t)kb * w_kb_bytes, w_kb_bytes, bar_w);
    umma::mbar_wait(bar_w, 0);
    GLB_TS(5);
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc_bf16(kTileM, p.N);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t a_base = umma::smem_u32(sA + (size_t)kb * (kTileM * 128));
      const uint32_t b_base = umma::smem_u32(sW + (size_t)kb * w_kb_bytes);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma::mma_bf16_ss(tmem_base, umma::make_desc_sw128(a_base + k4 * 32),
                          umma::make_desc_sw128(b_base + k4 * 32), idesc, (kb | k4) ? 1u : 0u);
    }
    umma::mma_commit(bar_mma);
    GLB_TS(6);
  }
  __syncwarp();
  umma::mbar_wait(bar_mma, 0);
  GLB_TS(7);
  umma::tc_fence_after();
  sage_epilogue(p, tmem_base, m0, R, warp, lane);
  GLB_TS(8);
  umma::tc_fence_before();
  __syncthreads();
  GLB_TS(9);
  if (warp == 1) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// EXPERIMENTAL (gather_mode = 5, not measured in round 1 - see DESIGN.md "Round-2 plan"): the same tile with
// TWO resident CTAs per SM.  ncu shows the default kernel at 23 % of the copy roofline with 40 % long-scoreboard
// and 25 % barrier stalls: one 1024-thread CTA per SM serialises id staging -> gather -> MMA -> epilogue, and
// while it is outside the gather no loads are in flight on that SM.  Here the CTA has 512 threads and keeps only
// the A tile (nkb x 16 KB) plus a 2-deep ring of W slices of kNS = 32 output columns (nkb x 4 KB each) in
// shared memory (<= ~110 KB for K = 256), so two CTAs fit and their phases interleave.  Thread 0 streams the W
// slices with bulk copies and issues M128 x N32 x K16 MMAs slice by slice into disjoint TMEM column ranges
// (2 x 256 columns for the two CTAs = the whole TMEM).  Every mbarrier has exactly one waiter (thread 0).
// ---------------------------------------------------------------------------------------------
constexpr int kOccThreads = 512;
constexpr int kOccWarps = kOccThreads / 32;
constexpr int kNS = 32;                 // output columns per W slice

template <int U, int DT>
__global__ void __launch_bounds__(kOccThreads, 2) sage_fused_occ2_kernel(const SageParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int k_total = p.kp_self + p.kp_nbr;
  const int nkb = k_total >> 6;
  const int n_slices_w = p.N / kNS;
  uint8_t* sA = smem;
  const uint32_t slice_kb_bytes = (uint32_t)kNS * 128u;                 // one k-block of one slice: 4 KB
  const uint32_t slice_bytes = (uint32_t)nkb * slice_kb_bytes;
  uint8_t* sW = sA + (size_t)nkb * (kTileM * 128);                      // ring: 2 slices
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 2 * (size_t)slice_bytes);
  uint64_t* full = bars;                 // [2]
  uint64_t* empty = bars + 2;            // [2]
  uint64_t* bar_mma = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const char** sPtrN = reinterpret_cast<const char**>(bars + 8);        // [R * k]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tid == 0) {
    umma::mbar_init(full + 0, 1); umma::mbar_init(full + 1, 1);
    umma::mbar_init(empty + 0, 1); umma::mbar_init(empty + 1, 1);
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

  const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w_img);
  const uint32_t w_kb_bytes = (uint32_t)p.N * 128u;                     // one k-block of the FULL image
  auto load_slice = [&](int s, int b) {                                 // thread 0 only
    umma::mbar_arrive_expect_tx(full + b, slice_bytes);
    for (int kb = 0; kb < nkb; ++kb)
      umma::bulk_g2s(sW + (size_t)b * slice_bytes + (size_t)kb * slice_kb_bytes,
                     wsrc + (size_t)kb * w_kb_bytes + (size_t)s * slice_kb_bytes, slice_kb_bytes, full + b);
  };
  if (tid == 0) {                        // the first two slices stream in behind the gather
    load_slice(0, 0);
    if (n_slices_w > 1) load_slice(1, 1);
  }

  // --- phase 0: ids -> row pointers (same as the default kernel)
  constexpr int VEC = Chunk<DT>::kVec;
  const int R = p.rows_per_cta;
  const int m0 = blockIdx.x * R;
  const int k = p.k;
  const char** sPtrS = sPtrN + (size_t)R * k;
  const bool need_self = p.kp_self > 0 || p.mode == kGcnMean;
  {
    const uint32_t nbr_row_bytes = (uint32_t)p.tnbr.stride * (DT == 0 ? 4u : 2u);
    const uint32_t self_row_bytes = (uint32_t)p.tself.stride * (DT == 0 ? 4u : 2u);
    const int64_t base = (int64_t)m0 * k;
    const int64_t lim = (int64_t)p.M * k;
    for (int i = tid; i < R * k; i += kOccThreads) {
      const int64_t idx = base + i;
      sPtrN[i] = idx < lim ? vid_ptr(p.tnbr, p.nbr_vids ? __ldg(p.nbr_vids + idx) : idx, p.wshift_nbr, nbr_row_bytes, p.zero_row) : p.zero_row;
    }
    for (int i = tid; i < R; i += kOccThreads) {
      const int m = m0 + i;
      sPtrS[i] = (need_self && m < p.M) ? vid_ptr(p.tself, p.self_vids ? __ldg(p.self_vids + m) : (int64_t)m, p.wshift_self, self_row_bytes, p.zero_row) : p.zero_row;
    }
  }
  __syncthreads();

  // --- phase 1: gather + aggregate -> A tile (identical math / layout to the default kernel)
  {
    const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
    const int lanes_row = p.kp_nbr / VEC;
    const int lpr = lanes_row < 32 ? lanes_row : 32;
    const int lshift = 31 - __clz(lpr);
    const int rpi = 32 >> lshift;
    const int n_sl = lanes_row > 32 ? lanes_row >> 5 : 1;
    const int row_groups = (R + rpi - 1) / rpi;
    const int n_items = row_groups * n_sl;
    const int sub = lane >> lshift;
    const int lig = lane & (lpr - 1);
    const bool has_self = p.kp_self > 0;
    float scale = 1.f;
    if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
    else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);
    for (int item = warp; item < n_items; item += kOccWarps) {
      const int rg = n_sl == 1 ? item : item / n_sl;
      const int sl = n_sl == 1 ? 0 : item - rg * n_sl;
      const int r = rg * rpi + sub;
      if (r >= R) continue;
      const int m = m0 + r;
      const int chunk = lig + 32 * sl;
      const int f0 = chunk * VEC;
      const size_t coff = (size_t)chunk * 16;
      float acc[VEC], sv[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) { acc[i] = 0.f; sv[i] = 0.f; }
      if (f0 < d_nbr) {
        const char* const* ptrs = sPtrN + (size_t)r * k;
        Chunk<DT> sraw;
        const bool self_ld = need_self && f0 < d_self;
        if (self_ld) sraw.load(sPtrS[r] + coff);
        for (int j0 = 0; j0 < k; j0 += U) {
          Chunk<DT> raw[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = j0 + u < k ? j0 + u : k - 1;
            raw[u].load(ptrs[j] + coff);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (j0 + u < k) raw[u].add_to(acc);
        }
        if (self_ld) sraw.add_to(sv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (f0 + i >= d_self) sv[i] = 0.f;
          if (f0 + i >= d_nbr) acc[i] = 0.f;
        }
      } else if (need_self && f0 < d_self) {
        Chunk<DT> sraw;
        sraw.load(sPtrS[r] + coff);
        sraw.add_to(sv);
#pragma unroll
        for (int i = 0; i < VEC; ++i) if (f0 + i >= d_self) sv[i] = 0.f;
      }
      const size_t a_off = (size_t)m * k_total;
      __nv_bfloat16* asave = (p.a_save && m < p.M) ? p.a_save : nullptr;
      if (has_self) put_chunk<VEC>(sA, asave, a_off, r, f0, sv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = (acc[i] + (p.mode == kGcnMean ? sv[i] : 0.f)) * scale;
      put_chunk<VEC>(sA, asave, a_off, r, p.kp_self + f0, acc);
    }
  }
  umma::fence_proxy_async_smem();
  __syncthreads();

  // --- GEMM: slice s of W (kNS output columns) x the whole A tile -> TMEM columns [s*kNS, (s+1)*kNS)
  if (tid == 0) {
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc_bf16(kTileM, kNS);
    for (int s = 0; s < n_slices_w; ++s) {
      const int b = s & 1;
      umma::mbar_wait(full + b, (uint32_t)((s >> 1) & 1));
      umma::tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t a_base = umma::smem_u32(sA + (size_t)kb * (kTileM * 128));
        const uint32_t b_base = umma::smem_u32(sW + (size_t)b * slice_bytes + (size_t)kb * slice_kb_bytes);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma::mma_bf16_ss(tmem_base + (uint32_t)(s * kNS), umma::make_desc_sw128(a_base + k4 * 32),
                            umma::make_desc_sw128(b_base + k4 * 32), idesc, (kb | k4) ? 1u : 0u);
      }
      umma::mma_commit(empty + b);                       // buffer b is free once these MMAs have read it
      if (s >= 1 && s + 1 < n_slices_w) {                 // refill the buffer slice s-1 used with slice s+1
        const int pb = (s - 1) & 1;
        umma::mbar_wait(empty + pb, (uint32_t)(((s - 1) >> 1) & 1));
        load_slice(s + 1, pb);
      }
    }
    umma::mma_commit(bar_mma);
  }
  __syncwarp();
  umma::mbar_wait(bar_mma, 0);
  umma::tc_fence_after();

  // --- epilogue: 16 warps = 4 TMEM lane quarters x 4 column groups
  {
    const int q = warp & 3, g = warp >> 2;
    const int cols_per_group = p.N / 4;                   // multiple of 16 (N is a multiple of 64)
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = row < R && m < p.M;
    for (int c0 = 0; c0 < cols_per_group; c0 += 16) {
      const int n0 = g * cols_per_group + c0;
      uint32_t v[16];
      umma::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, v);
      umma::tmem_ld_wait();
      if (row_ok && n0 < p.n_out) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (n0 + i >= p.n_out) continue;
          float x = __uint_as_float(v[i]);
          if (p.bias) x += __ldg(p.bias + n0 + i);
          if (p.relu) x = fmaxf(x, 0.f);
          if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[(size_t)m * p.out_stride + n0 + i] = __float2bfloat16(x);
          else reinterpret_cast<float*>(p.out)[(size_t)m * p.out_stride + n0 + i] = x;
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// Split path, stage 1 (used when rows live on peer GPUs): a small, maximum-occupancy kernel
// (256 threads, ~40 registers -> 48-64 resident warps per SM) that only gathers + aggregates and
// writes the bf16 A matrix [M, K_total] = [ self || agg(nbrs) ] (zero K-padding included).  NVLink
// peer reads need far more requests in flight than the 1-CTA/SM fused kernel can keep up
// (measured: 580-640 GB/s remote for this shape vs ~170 GB/s inside the fused kernel); stage 2 is
// the same tcgen05 kernel fed from this (L2-resident) A instead of gathering itself.
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256, 5) gather_self_mean_kernel(const SageParams p) {
  constexpr int VEC = Chunk<DT>::kVec;
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int k = p.k;
  const int k_total = p.kp_self + p.kp_nbr;
  const int lanes_row = p.kp_nbr / VEC;
  const int lpr = lanes_row < 32 ? lanes_row : 32;
  const int lshift = 31 - __clz(lpr);
  const int rpi = 32 >> lshift;
  const int n_slices = lanes_row > 32 ? lanes_row >> 5 : 1;
  const int sub = lane >> lshift, lig = lane & (lpr - 1);
  const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
  const bool has_self = p.kp_self > 0;
  const bool need_self = has_self || p.mode == kGcnMean;
  const uint32_t nbr_row_bytes = (uint32_t)p.tnbr.stride * (DT == 0 ? 4u : 2u);
  const uint32_t self_row_bytes = (uint32_t)p.tself.stride * (DT == 0 ? 4u : 2u);
  float scale = 1.f;
  if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
  else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);
  const int64_t n_items = ((int64_t)p.M + rpi - 1) / rpi * n_slices;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += warps) {
    const int64_t rg = item / n_slices;
    const int sl = (int)(item - rg * n_slices);
    const int64_t m = rg * rpi + sub;
    if (m >= p.M) continue;
    const int chunk = lig + 32 * sl;
    const int f0 = chunk * VEC;
    const size_t coff = (size_t)chunk * 16;
    float acc[VEC], sv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { acc[i] = 0.f; sv[i] = 0.f; }
    if (f0 < d_nbr || (need_self && f0 < d_self)) {
      Chunk<DT> sraw;
      const bool self_ld = need_self && f0 < d_self;
      if (self_ld) {
        sraw.load(vid_ptr(p.tself, p.self_vids ? __ldg(p.self_vids + m) : m, p.wshift_self, self_row_bytes, p.zero_row) + coff);
      }
      if (f0 < d_nbr) {
        const int64_t base = m * k;
        for (int j0 = 0; j0 < k; j0 += U) {
          Chunk<DT> raw[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = j0 + u < k ? j0 + u : k - 1;
            raw[u].load(vid_ptr(p.tnbr, p.nbr_vids ? __ldg(p.nbr_vids + base + j) : base + j, p.wshift_nbr, nbr_row_bytes, p.zero_row) + coff);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (j0 + u < k) raw[u].add_to(acc);
        }
      }
      if (self_ld) sraw.add_to(sv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (f0 + i >= d_self) sv[i] = 0.f;
        if (f0 + i >= d_nbr) acc[i] = 0.f;
      }
    }
    __nv_bfloat16* out = p.a_save + (size_t)m * k_total;
    auto store = [&](int kcol, const float (&v)[VEC]) {
      if constexpr (VEC == 4) {
        uint2 u2; u2.x = pack_bf16x2(v[0], v[1]); u2.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(out + kcol) = u2;
      } else {
        uint4 u4; u4.x = pack_bf16x2(v[0], v[1]); u4.y = pack_bf16x2(v[2], v[3]);
        u4.z = pack_bf16x2(v[4], v[5]); u4.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(out + kcol) = u4;
      }
    };
    if (has_self) store(f0, sv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = (acc[i] + (p.mode == kGcnMean ? sv[i] : 0.f)) * scale;
    store(p.kp_self + f0, acc);
  }
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

// fp32 row-major padded weight [n_real, k_total] -> (SW128 bf16 image, bf16 row-major copy) in one pass
__global__ void pack_sw128_f32_kernel(const float* __restrict__ w, int n_real, int N, int k_total,
                                      uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ w16) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte (8 element) chunk each
  int nkb = k_total >> 6;
  int total = nkb * N * 8;
  if (idx >= total) return;
  int c = idx & 7;
  int n = (idx >> 3) % N;
  int kb = (idx >> 3) / N;
  uint4 val = make_uint4(0, 0, 0, 0);
  if (n < n_real) {
    const float4* src = reinterpret_cast<const float4*>(w + (size_t)n * k_total + kb * 64 + c * 8);
    float4 a = src[0], b = src[1];
    val.x = pack_bf16x2(a.x, a.y); val.y = pack_bf16x2(a.z, a.w);
    val.z = pack_bf16x2(b.x, b.y); val.w = pack_bf16x2(b.z, b.w);
    if (w16) *reinterpret_cast<uint4*>(w16 + (size_t)n * k_total + kb * 64 + c * 8) = val;
  }
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

// returns (image, bf16 row-major copy)
std::vector<at::Tensor> pack_weight_f32(const at::Tensor& w_padded, int64_t N, bool want_rowmajor) {
  TORCH_CHECK(w_padded.is_cuda() && w_padded.scalar_type() == at::kFloat && w_padded.dim() == 2 &&
              w_padded.is_contiguous(), "w_padded must be a contiguous CUDA fp32 [n, k_total] tensor");
  int64_t n_real = w_padded.size(0), k_total = w_padded.size(1);
  TORCH_CHECK(k_total % 64 == 0 && N % 8 == 0 && N >= n_real, "bad padded weight shape");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(w_padded.data_ptr()) & 15) == 0, "weight must be 16 B aligned");
  c10::cuda::CUDAGuard guard(w_padded.device());
  auto o16 = w_padded.options().dtype(at::kBFloat16);
  auto img = at::empty({(k_total / 64) * N * 64}, o16);
  at::Tensor w16;
  if (want_rowmajor) w16 = at::empty({n_real, k_total}, o16);
  int total = (int)((k_total / 64) * N * 8);
  pack_sw128_f32_kernel<<<(total + 255) / 256, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      w_padded.data_ptr<float>(), (int)n_real, (int)N, (int)k_total, reinterpret_cast<uint8_t*>(img.data_ptr()),
      want_rowmajor ? reinterpret_cast<__nv_bfloat16*>(w16.data_ptr()) : nullptr);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {img, want_rowmajor ? w16 : at::Tensor()};
}

std::vector<at::Tensor> sage_fused_forward(const at::Tensor& tself_desc,
                                           const c10::optional<at::Tensor>& self_vids,
                                           const at::Tensor& tnbr_desc,
                                           const c10::optional<at::Tensor>& nbr_vids, int64_t M,
                                           int64_t k, int64_t mode, const at::Tensor& w_img,
                                           const c10::optional<at::Tensor>& bias, int64_t N,
                                           int64_t n_out, bool relu, bool out_bf16, bool save_a, int64_t rows_per_cta,
                                           const c10::optional<at::Tensor>& out_buf,
                                           const c10::optional<at::Tensor>& a_buf,
                                           const c10::optional<at::Tensor>& debug_ts, int64_t gather_mode) {
  // gather_mode 4 = split path: gather_self_mean_kernel -> A (global) -> tcgen05 kernel staged from A
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
  size_t smem = (size_t)sage_smem_bytes(k_total, N);   // + locator staging, added below
  TORCH_CHECK(smem <= 232448, "tile does not fit in shared memory (", smem, " B); use the unfused path");
  at::Tensor sv, nv, b;
  p.self_vids = nullptr; p.nbr_vids = nullptr; p.bias = nullptr;
  if (self_vids.has_value()) { sv = self_vids->contiguous(); check_cuda_i64(sv, "self_vids");
    TORCH_CHECK(sv.numel() == M); p.self_vids = sv.data_ptr<int64_t>(); }
  if (nbr_vids.has_value()) { nv = nbr_vids->contiguous(); check_cuda_i64(nv, "nbr_vids");
    TORCH_CHECK(nv.numel() == M * k); p.nbr_vids = nv.data_ptr<int64_t>(); }
  if (bias.has_value()) { b = bias->contiguous();
    TORCH_CHECK(b.is_cuda() && b.scalar_type() == at::kFloat && b.numel() >= n_out, "bias must be fp32 [>= n_out]");
    p.bias = b.data_ptr<float>(); }
  auto opts = w_img.options();
  at::Tensor out;
  if (out_buf.has_value()) {
    out = *out_buf;
    TORCH_CHECK(out.is_cuda() && out.dim() == 2 && out.size(0) == M && out.size(1) == n_out && out.stride(1) == 1,
                "out_buf must be [M, n_out] with unit inner stride");
    TORCH_CHECK(out.scalar_type() == (out_bf16 ? at::kBFloat16 : at::kFloat), "out_buf dtype mismatch");
  } else {
    out = at::empty({M, n_out}, opts.dtype(out_bf16 ? at::kBFloat16 : at::kFloat));
  }
  at::Tensor a_save;
  p.a_save = nullptr;
  if (save_a) {
    if (a_buf.has_value()) {
      a_save = *a_buf;
      TORCH_CHECK(a_save.scalar_type() == at::kBFloat16 && a_save.dim() == 2 && a_save.size(0) == M &&
                  a_save.size(1) == k_total && a_save.stride(1) == 1 && a_save.stride(0) == k_total,
                  "a_buf must be a row-contiguous bf16 [M, K_total] view");
    } else {
      a_save = at::empty({M, (int64_t)k_total}, opts.dtype(at::kBFloat16));
    }
    p.a_save = reinterpret_cast<__nv_bfloat16*>(a_save.data_ptr());
  }
  p.a_src = nullptr;
  p.w_img = w_img.data_ptr();
  p.out = out.data_ptr();
  p.out_stride = out.stride(0);
  p.N = (int)N; p.n_out = (int)n_out; p.relu = relu ? 1 : 0; p.out_bf16 = out_bf16 ? 1 : 0;
  p.tmem_cols = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  if (M == 0) return {out, save_a ? a_save : at::Tensor()};
  // spread small M over the whole chip: aim for >= 2 CTAs' worth of work per SM-wave but never
  // more than 128 rows per CTA; keep R a multiple of 8 (one 1024-byte swizzle atom)
  int R = kTileM;
  if (rows_per_cta > 0) R = (int)std::min<int64_t>(kTileM, std::max<int64_t>(8, (rows_per_cta + 7) / 8 * 8));
  else {
    // balance whole waves: w = number of 148-CTA waves needed at <= 128 rows per CTA, then the
    // smallest R (multiple of 8) that still fits M into w waves
    const int64_t sms = 148;
    int64_t waves = (M + sms * kTileM - 1) / (sms * kTileM);
    int64_t want = (M + sms * waves - 1) / (sms * waves);
    R = (int)std::min<int64_t>(kTileM, std::max<int64_t>(8, (want + 7) / 8 * 8));
  }
  // shared-memory row-pointer staging: 8 B per (row, neighbour) + 8 B per row; shrink R until it fits
  const size_t bar_bytes = 32 + 2 * kMaxSlots * 8;          // control words + TMA ring barriers
  while (R > 8 && smem + bar_bytes + (size_t)R * (k + 1) * 8 > 232448) R -= 8;
  TORCH_CHECK(smem + bar_bytes + (size_t)R * (k + 1) * 8 <= 232448, "fan-out too large for the fused kernel's id staging");
  smem += bar_bytes + (size_t)R * (k + 1) * 8;
  p.rows_per_cta = R;
  p.debug_ts = nullptr;
  if (debug_ts.has_value()) {
    TORCH_CHECK(debug_ts->scalar_type() == at::kLong && debug_ts->numel() >= (int64_t)((M + R - 1) / R) * 16);
    p.debug_ts = reinterpret_cast<long long*>(debug_ts->data_ptr<int64_t>());
  }
  TORCH_CHECK(p.tself.dtype == p.tnbr.dtype, "fused SAGE layer needs self / neighbour tables of the same dtype");
  TORCH_CHECK(p.kp_self == 0 || p.kp_self == p.kp_nbr, "fused SAGE layer needs equally padded self / neighbour dims");
  auto log2_or_neg = [](int w) { int s = 0; while ((1 << s) < w) ++s; return (1 << s) == w ? s : -1; };
  p.wshift_self = log2_or_neg(p.tself.world);
  p.wshift_nbr = log2_or_neg(p.tnbr.world);
  TORCH_CHECK((p.tself.stride * (p.tself.dtype == 0 ? 4 : 2)) % 16 == 0 && (p.tnbr.stride * (p.tnbr.dtype == 0 ? 4 : 2)) % 16 == 0,
              "table rows must be 16-byte aligned (stride multiple of 4 fp32 / 8 bf16 elements)");
  {
    static std::vector<at::Tensor> zero_rows(64);
    int dev = w_img.get_device();
    if (!zero_rows[dev].defined()) zero_rows[dev] = at::zeros({1024}, opts.dtype(at::kFloat));
    p.zero_row = reinterpret_cast<const char*>(zero_rows[dev].data_ptr());
  }
  unsigned grid = (unsigned)((M + R - 1) / R);
  auto stream = at::cuda::getCurrentCUDAStream();
  const int u = k <= 4 ? 4 : (k % 5 == 0 || k > 12) ? 5 : 6;
  const int dt = p.tnbr.dtype;
  if (gather_mode == 5) {
    // EXPERIMENTAL two-CTA-per-SM variant (see sage_fused_occ2_kernel): falls through to the default kernel when
    // the tile does not fit twice into an SM.
    const size_t nkb5 = (size_t)k_total / 64;
    const size_t base5 = nkb5 * (kTileM * 128) + 2 * nkb5 * (kNS * 128) + 64 + 1024;     // A + W ring + barriers + alignment
    const int64_t per_wave = 148 * 2;
    const int64_t waves5 = (M + per_wave * kTileM - 1) / (per_wave * kTileM);
    int R5 = (int)std::min<int64_t>(kTileM, std::max<int64_t>(8, ((M + per_wave * waves5 - 1) / (per_wave * waves5) + 7) / 8 * 8));
    if (rows_per_cta > 0) R5 = R;
    const size_t limit5 = 112 * 1024;      // 2 x (dynamic + 1 KB static + 1 KB reserved) <= 228 KB per SM
    while (R5 > 8 && base5 + (size_t)R5 * (k + 1) * 8 > limit5) R5 -= 8;
    if (base5 + (size_t)R5 * (k + 1) * 8 <= limit5 && !p.debug_ts) {
      const size_t smem5 = base5 + (size_t)R5 * (k + 1) * 8;
      p.rows_per_cta = R5;
      const unsigned grid5 = (unsigned)((M + R5 - 1) / R5);
#define LAUNCH5(UU, DD)                                                                           \
  do {                                                                                            \
    C10_CUDA_CHECK(cudaFuncSetAttribute(sage_fused_occ2_kernel<UU, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit5)); \
    sage_fused_occ2_kernel<UU, DD><<<grid5, kOccThreads, smem5, stream>>>(p);                     \
  } while (0)
      if (dt == 0) { if (u == 4) LAUNCH5(4, 0); else if (u == 5) LAUNCH5(5, 0); else LAUNCH5(6, 0); }
      else         { if (u == 4) LAUNCH5(4, 1); else if (u == 5) LAUNCH5(5, 1); else LAUNCH5(6, 1); }
#undef LAUNCH5
      C10_CUDA_KERNEL_LAUNCH_CHECK();
      return {out, save_a ? a_save : at::Tensor()};
    }
    p.rows_per_cta = R;
  }
  // MEASURED (2 GPUs, fp32 rows): fused register path 7.7k steps/s vs split 6.9k - the remote rows are
  // NVLink-bandwidth bound either way (394 GB/s achieved vs 580 GB/s best random-row rate), so the
  // split path stays opt-in.
  const bool split = gather_mode == 4;
  if (split) {
    if (!a_save.defined()) {
      a_save = at::empty({M, (int64_t)k_total}, opts.dtype(at::kBFloat16));
      p.a_save = reinterpret_cast<__nv_bfloat16*>(a_save.data_ptr());
    }
    const int lanes_row = p.kp_nbr / (dt == 0 ? 4 : 8);
    const int rpi = lanes_row < 32 ? 32 / lanes_row : 1;
    const int n_slices = lanes_row > 32 ? lanes_row / 32 : 1;
    const int64_t items = (M + rpi - 1) / rpi * n_slices;
    const unsigned gblocks = (unsigned)std::min<int64_t>((items + 7) / 8, 148 * 8);
    if (dt == 0) gather_self_mean_kernel<0><<<gblocks, 256, 0, stream>>>(p);
    else         gather_self_mean_kernel<1><<<gblocks, 256, 0, stream>>>(p);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    p.a_src = p.a_save;
    p.a_save = nullptr;
  }
  // gather_mode: 1 = register-staged loads, 2 = TMA bulk copies into the shared-memory ring,
  // 3 = per-lane cp.async into the same ring; 0 = auto = cp.async ring for multi-wave launches and
  // whenever rows may live on a peer GPU, register-staged otherwise.  MEASURED (profiles/): per-row cp.async.bulk copies of 400 B cost
  // ~110 cycles each in the TMA unit (1408 copies -> 153 k cycles per 128-row tile, 0.76 TB/s), 4x
  // slower than the register path (36 k cycles) - the ring variant is kept for wide rows only.
  const size_t slot_bytes = (size_t)(k + 1) * (size_t)p.tnbr.stride * (dt == 0 ? 4 : 2);
  const bool tma_ok = k >= 1 && slot_bytes <= (size_t)k_total / 64 * N * 128;
  bool use_tma = gather_mode == 2;
  // cp.async ring (mode 3): needs >= kProducers slots
  const bool async_ok = k >= 1 && slot_bytes * kProducers <= (size_t)k_total / 64 * N * 128;
  bool use_async = gather_mode == 3 && async_ok;   // MEASURED slower than the register path (2-GPU: 5.4k vs 7.7k steps/s)
  use_tma = use_tma && tma_ok;
#define SET_ATTR(KERNEL)                                                                          \
  do {                                                                                            \
    static bool attr_done = false;                                                                \
    if (!attr_done) {                                                                             \
      C10_CUDA_CHECK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)); \
      attr_done = true;                                                                           \
    }                                                                                             \
  } while (0)
#define LAUNCH(UU, DD)                                                                            \
  do {                                                                                            \
    SET_ATTR((sage_fused_fwd_kernel<UU, DD>));                                                    \
    sage_fused_fwd_kernel<UU, DD><<<grid, kThreads, smem, stream>>>(p);                           \
  } while (0)
  if (use_async) {
    if (dt == 0) { SET_ATTR(sage_fused_async_kernel<0>); sage_fused_async_kernel<0><<<grid, kThreads, smem, stream>>>(p); }
    else         { SET_ATTR(sage_fused_async_kernel<1>); sage_fused_async_kernel<1><<<grid, kThreads, smem, stream>>>(p); }
  } else if (use_tma) {
    if (dt == 0) { SET_ATTR(sage_fused_tma_kernel<0>); sage_fused_tma_kernel<0><<<grid, kThreads, smem, stream>>>(p); }
    else         { SET_ATTR(sage_fused_tma_kernel<1>); sage_fused_tma_kernel<1><<<grid, kThreads, smem, stream>>>(p); }
  } else if (dt == 0) { if (u == 4) LAUNCH(4, 0); else if (u == 5) LAUNCH(5, 0); else LAUNCH(6, 0); }
  else                { if (u == 4) LAUNCH(4, 1); else if (u == 5) LAUNCH(5, 1); else LAUNCH(6, 1); }
#undef LAUNCH
#undef SET_ATTR
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {out, (save_a || split) ? a_save : at::Tensor()};
}

// K7 standalone: out = act(A . W^T + b) for a dense bf16 A [M, K] (K multiple of 64, <= 512; N <= 256)
// on the same tcgen05 / TMEM tile kernel (A staged from global memory instead of gathered).
at::Tensor tc_linear_forward(const at::Tensor& a, const at::Tensor& w_img, const c10::optional<at::Tensor>& bias,
                             int64_t N, int64_t n_out, bool relu, bool out_bf16) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && a.dim() == 2 && a.is_contiguous(),
              "A must be a contiguous CUDA bf16 [M, K] matrix");
  const int64_t M = a.size(0), K = a.size(1);
  TORCH_CHECK(K % 64 == 0 && K >= 64 && K <= 512, "K must be a multiple of 64 in [64, 512]");
  TORCH_CHECK(N % 64 == 0 && N >= 64 && N <= 256 && n_out <= N && n_out >= 1);
  TORCH_CHECK(w_img.numel() == K * N && w_img.scalar_type() == at::kBFloat16, "weight image size mismatch");
  c10::cuda::CUDAGuard guard(a.device());
  SageParams p;
  std::memset(&p, 0, sizeof(p));
  p.tself.world = 1; p.tnbr.world = 1;
  p.tself.dtype = 1; p.tnbr.dtype = 1;
  p.M = (int)M; p.k = 0; p.mode = kConcatMean;
  p.kp_self = 0; p.kp_nbr = (int)K;
  at::Tensor b;
  if (bias.has_value()) { b = bias->contiguous();
    TORCH_CHECK(b.is_cuda() && b.scalar_type() == at::kFloat && b.numel() >= n_out); p.bias = b.data_ptr<float>(); }
  auto out = at::empty({M, n_out}, a.options().dtype(out_bf16 ? at::kBFloat16 : at::kFloat));
  p.w_img = w_img.data_ptr(); p.out = out.data_ptr(); p.out_stride = n_out;
  p.N = (int)N; p.n_out = (int)n_out; p.relu = relu ? 1 : 0; p.out_bf16 = out_bf16 ? 1 : 0;
  p.tmem_cols = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  p.a_src = reinterpret_cast<const __nv_bfloat16*>(a.data_ptr());
  if (M == 0) return out;
  size_t smem = (size_t)sage_smem_bytes(K, N);
  TORCH_CHECK(smem <= 232448, "tile does not fit in shared memory");
  const int64_t sms = 148;
  int64_t waves = (M + sms * kTileM - 1) / (sms * kTileM);
  int64_t want = (M + sms * waves - 1) / (sms * waves);
  int R = (int)std::min<int64_t>(kTileM, std::max<int64_t>(8, (want + 7) / 8 * 8));
  p.rows_per_cta = R;
  smem += 32 + 2 * kMaxSlots * 8 + 64;
  unsigned grid = (unsigned)((M + R - 1) / R);
  static bool attr_done = false;
  if (!attr_done) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(sage_fused_fwd_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done = true;
  }
  sage_fused_fwd_kernel<4, 1><<<grid, kThreads, smem, at::cuda::getCurrentCUDAStream()>>>(p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return out;
}

}  // namespace glb
