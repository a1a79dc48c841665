// K6+K7 fused: GraphSAGE / EgoSAGE layer forward as ONE persistent, warp-specialised kernel
//
//     out[m, :] = act( [ x_self[m] || agg_j x_nbr[m, j] ] . W^T + b )
//
// One CTA per SM stays resident and walks a static list of row tiles (<= 64 destination rows each)
// that may span SEVERAL segments of the ego graph (e.g. seeds<-hop1 with k=25 and hop1<-hop2 with
// k=10 of a 2-layer GraphSAGE): the tile height of every segment is chosen on the host so that all
// tiles cost about the same number of row fetches and the tile count is a multiple of the SM count.
// Inside the CTA the phases of a tile overlap instead of running back to back:
//
//   gather warps (GW, default 19)   each warp owns work items (1-4 destination rows x one 512-byte slice): it
//                             translates the item's ids into row pointers (local HBM, a peer GPU's HBM over
//                             NVLink, or the local replica cache) in a private SMEM scratch - the ids of the
//                             NEXT item are prefetched while the current rows are in flight -, pulls the self
//                             row and the k neighbour rows with batches of 16-byte loads, reduces in fp32
//                             registers and writes the bf16 A tile in the UMMA K-major SWIZZLE_128B layout
//                             (and row-major to global for the backward pass); items are dealt round-robin
//                             ACROSS tile boundaries so no warp idles; the A tile is double buffered; rows may be
//                             fp32, bf16 or block-scaled fp8 (dequantised while accumulating)
//   MMA warp (1)              W image fetched ONCE per CTA by the TMA engine (cp.async.bulk); one
//                             elected thread issues tcgen05.mma (M=64, N<=256, K=16) into one of two
//                             TMEM accumulators; tcgen05.commit frees the A tile and publishes the
//                             accumulator through mbarriers
//   epilogue warps (EW, default 4)  tcgen05.ld -> bias -> ReLU -> bf16/fp32 store of tile t while the gather
//                             warps are already loading tile t+1; for the top layer a second, coalesced
//                             phase (one warp per row, one lane per class) computes the softmax
//                             cross-entropy loss, dlogits and the bias gradient (fused CE)
//
// The rows never round-trip through HBM ("agg(X).W" instead of materialising [B*k, D]).
//
// Math parity: EgoSAGEConv (graphlearn/python/nn/tf/layers/ego_sage_conv.py:71-106):
// agg in {mean,sum} over neighbor.reshape(-1,k,d), out = W.[x || agg]; 'gcn' = mean over {x} U nbrs
// then W.  Loss parity: graphlearn/python/nn/tf/loss.py (softmax cross entropy).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cfloat>
#include <cstring>
#include "host_utils.h"
#include "umma.cuh"
#include "gather_common.cuh"

namespace glb {

constexpr int kTileM = 64;                         // destination rows per tile = UMMA M (two A tiles fit next to the W image)
constexpr int kABufs = 2;
// Role layout (template parameters EW / GW of the kernel): warps [0, EW) epilogue (TMEM lane quarter = warp & 3, 32-column
// chunks interleaved by warp >> 2), warp EW = MMA issuer, warps [EW + 1, EW + 1 + GW) gather.  The register file is
// split evenly over the CTA's threads, so the shape trades gather warps against row loads in flight per lane:
//   (8, 23) 1024 threads x 64 regs     (4, 19) 768 threads x 80 regs     (4, 11) 512 threads x 128 regs
constexpr int kMaxGatherWarps = 23;                 // sizes the pointer scratch
constexpr int kMaxSegs = 4;
constexpr int kScrCap = 64;                        // row pointers per gather warp (private scratch): rows_per_item * (k + 1) <= 64
constexpr size_t kSmemLimit = 232448;

enum SageMode : int { kConcatMean = 0, kConcatSum = 1, kGcnMean = 2 };

struct SageSeg {
  const int64_t* self_vids;   // [M] or null (identity: self_base + m)
  const int64_t* nbr_vids;    // [M, k] or null (identity: nbr_base + m*k + j)
  int64_t self_base, nbr_base;
  char* out;                  // [M, n_out] rows of this segment
  __nv_bfloat16* a_save;      // [M, K_total] or null
  int M, k;
  int R;                      // destination rows per tile (<= 64, multiple of 8)
  int tile0;                  // index of the segment's first tile
};

struct SageParams {
  TableView tself;
  TableView tnbr;
  SageSeg seg[kMaxSegs];
  int nseg, total_tiles;
  const void* w_img;          // bf16 image, (K_total/64) blocks of [N x 64] SW128
  const float* bias;          // [bias_len] or null
  int bias_len;
  int64_t out_stride;         // elements
  int kp_self, kp_nbr;        // padded K of each half, in {0,64,128,256,512}
  int mode;
  int N;                      // padded output width: multiple of 16, <= 256
  int n_out;                  // real output width
  int relu;
  int out_bf16;
  int tmem_cols;
  const char* zero_row;       // >= 2 KB of zeros: target of the loads of masked lanes / missing rows
  int wshift_self, wshift_nbr; // log2(world) of each table when it is a power of two, else -1
  // fused softmax cross-entropy (top layer): active when ce_labels != nullptr
  const int64_t* ce_labels;   // label table
  const int64_t* ce_seeds;    // [M] vids (label = ce_labels[seed / ce_world]) or null (label = ce_labels[m])
  int ce_world;
  float* ce_loss;             // accumulated mean loss
  __nv_bfloat16* ce_dlogits;  // [M, ce_dl_stride], columns >= n_out are written as zero
  int ce_dl_stride;
  float* ce_dbias;            // [n_out] accumulated, or null
  float ce_inv_b;
  // several weight images in one launch (dA = dZ . W with K_total > 256): CTA b uses image b % n_imgs, walks the tiles
  // b / n_imgs + j * (gridDim.x / n_imgs) and writes output columns [img * N, img * N + n_out)
  int n_imgs;
  long long* dbg;             // optional [grid][64] clock64 trace (tools/persist_timeline.py), lane 0 of one warp per role
};

// write VEC consecutive K elements of tile row r starting at K column kcol (multiple of VEC)
template <int VEC>
__device__ __forceinline__ void put_chunk(uint8_t* sA, __nv_bfloat16* a_row, int r, int kcol, const float (&v)[VEC]) {
  const int kb = kcol >> 6;
  const uint32_t off = (uint32_t)kb * (kTileM * 128) + umma::sw128_offset((uint32_t)r, (uint32_t)(kcol & 63));
  if constexpr (VEC == 4) {
    uint2 u; u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(sA + off) = u;
    if (a_row) *reinterpret_cast<uint2*>(a_row + kcol) = u;
  } else {
    uint4 u; u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(sA + off) = u;
    if (a_row) *reinterpret_cast<uint4*>(a_row + kcol) = u;
  }
}

__device__ __forceinline__ int find_seg(const SageParams& p, int t) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < kMaxSegs; ++i)
    if (i < p.nseg && t >= p.seg[i].tile0) s = i;
  return s;
}

#define GLB_DBG(slot) do { if (p.dbg && lane == 0) p.dbg[(size_t)blockIdx.x * 64 + (slot)] = clock64(); } while (0)

// shared-memory control block (after the pointer buffers)
enum : int { kBarW = 0, kBarAFull = 1, kBarAFree = 3, kBarTFull = 5, kBarTFree = 7, kNumBars = 9 };

// ------------------------------------------------------------------------------------------------
// epilogue of one tile: TMEM -> registers -> bias/ReLU -> global (warps 0..3, one row per thread)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_cols32(const SageParams& p, const SageSeg& sg, int m, int n0, const float (&f)[32]) {
  if (n0 >= p.n_out) return;
  const bool full = (n0 + 32 <= p.n_out);
  if (p.out_bf16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(sg.out) + (size_t)m * p.out_stride + n0;
    if (full && (p.out_stride & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 a;
        a.x = pack_bf16x2(f[8 * i], f[8 * i + 1]); a.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
        a.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]); a.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
        reinterpret_cast<uint4*>(o)[i] = a;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.n_out) o[i] = __float2bfloat16(f[i]);
    }
  } else {
    float* o = reinterpret_cast<float*>(sg.out) + (size_t)m * p.out_stride + n0;
    if (full && (p.out_stride & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<float4*>(o)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < p.n_out) o[i] = f[i];
    }
  }
}

// bias vector staged once per CTA in shared memory (fp32 [N], zero beyond n_out)
__device__ __forceinline__ void load_cols32(const SageParams& p, const float* sBias, uint32_t taddr, int n0, float (&f)[32]) {
  uint32_t v[32];
  umma::tmem_ld32(taddr + (uint32_t)n0, v);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = *reinterpret_cast<const float4*>(sBias + n0 + 4 * i);     // same address for all lanes: broadcast
    float x0 = __uint_as_float(v[4 * i]) + b.x, x1 = __uint_as_float(v[4 * i + 1]) + b.y;
    float x2 = __uint_as_float(v[4 * i + 2]) + b.z, x3 = __uint_as_float(v[4 * i + 3]) + b.w;
    if (p.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
    f[4 * i] = x0; f[4 * i + 1] = x1; f[4 * i + 2] = x2; f[4 * i + 3] = x3;
  }
}

// phase A of the epilogue: TMEM -> registers -> bias/ReLU -> global.  Eight warps: lane quarter q = warp & 3,
// 32-column chunks interleaved between the two warps (warp >> 2) that share a quarter.
__device__ __forceinline__ void sage_epilogue_tile(const SageParams& p, const SageSeg& sg, const float* sBias, uint32_t tmem_acc,
                                                   int m0, int rows, int warp, int lane, int epi_warps) {
  // UMMA M = 64: accumulator row r lives in TMEM lane (r / 16) * 32 + (r % 16), i.e. lanes 0..15 of every quarter
  const int q = warp & 3, half = warp >> 2;
  if (q * 16 >= rows) return;                          // warp-uniform: this lane quarter holds no valid row
  const int row = q * 16 + lane;
  const int m = m0 + row;
  const bool row_ok = lane < 16 && row < rows;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  for (int n0 = half * 32; n0 < p.N; n0 += 8 * epi_warps) {
    if (n0 >= p.n_out) break;                          // padded output columns are never read back
    float f[32];
    load_cols32(p, sBias, taddr, n0, f);
    if (row_ok && sg.out) store_cols32(p, sg, m, n0, f);
  }
}

// phase B (top layer only): softmax cross-entropy of the tile's rows, one warp per row, lanes = classes
// (n_out <= 64: columns lane and lane + 32), reading the logits phase A has just written (L2).
__device__ __forceinline__ void sage_ce_tile(const SageParams& p, const SageSeg& sg, int m0, int rows, int warp, int lane, int epi_warps) {
  const float* logits = reinterpret_cast<const float*>(sg.out);
  const int c0 = lane, c1 = lane + 32;
  float db0 = 0.f, db1 = 0.f, lsum = 0.f;
  for (int r = warp; r < rows; r += epi_warps) {
    const int m = m0 + r;
    const float* lrow = logits + (size_t)m * p.out_stride;
    const float x0 = c0 < p.n_out ? __ldcg(lrow + c0) : -FLT_MAX;
    const float x1 = c1 < p.n_out ? __ldcg(lrow + c1) : -FLT_MAX;
    long long y = -1;
    if (lane == 0) y = p.ce_seeds ? __ldg(p.ce_labels + __ldg(p.ce_seeds + m) / p.ce_world) : __ldg(p.ce_labels + m);
    y = __shfl_sync(0xffffffffu, y, 0);
    const bool valid = y >= 0 && y < p.n_out;
    float mx = fmaxf(x0, x1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e0 = c0 < p.n_out ? __expf(x0 - mx) : 0.f, e1 = c1 < p.n_out ? __expf(x1 - mx) : 0.f;
    float se = e0 + e1;
    float xy = (c0 == (int)y ? x0 : 0.f) + (c1 == (int)y ? x1 : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); xy += __shfl_xor_sync(0xffffffffu, xy, o); }
    const float inv = 1.f / se;
    const float g0 = (valid && c0 < p.n_out) ? (e0 * inv - (c0 == (int)y ? 1.f : 0.f)) * p.ce_inv_b : 0.f;
    const float g1 = (valid && c1 < p.n_out) ? (e1 * inv - (c1 == (int)y ? 1.f : 0.f)) * p.ce_inv_b : 0.f;
    __nv_bfloat16* drow = p.ce_dlogits + (size_t)m * p.ce_dl_stride;
    if (c0 < p.ce_dl_stride) drow[c0] = __float2bfloat16(g0);
    if (c1 < p.ce_dl_stride) drow[c1] = __float2bfloat16(g1);
    db0 += g0; db1 += g1;
    if (valid) lsum += (mx + __logf(se) - xy) * p.ce_inv_b;
  }
  if (p.ce_dbias) {
    if (c0 < p.n_out && db0 != 0.f) atomicAdd(p.ce_dbias + c0, db0);
    if (c1 < p.n_out && db1 != 0.f) atomicAdd(p.ce_dbias + c1, db1);
  }
  if (lane == 0 && lsum != 0.f) atomicAdd(p.ce_loss, lsum);
}

// ------------------------------------------------------------------------------------------------
// gather-warp helpers
// ------------------------------------------------------------------------------------------------
struct TileCtx {       // kept small on purpose: the gather loop holds two of these next to its row loads
  int s;               // segment index
  int m0, rows;
};

struct GatherGeom {
  int rpi, n_slices, lshift, sub, lig;
  int tile_first, tile_stride;     // this CTA's tiles: tile_first + it * tile_stride
};

// position (it, item) -> first tile at or after `it` in which this warp still has an item; false past the last tile
__device__ __forceinline__ bool seek_item(const SageParams& p, const GatherGeom& gg, int& it, int& item, TileCtx& c) {
  for (;;) {
    const int t = gg.tile_first + it * gg.tile_stride;
    if (t >= p.total_tiles) return false;
    const int s = find_seg(p, t);
    const int m0 = (t - p.seg[s].tile0) * p.seg[s].R;
    const int rows = min(p.seg[s].R, p.seg[s].M - m0);
    const int n_items = ((rows + gg.rpi - 1) >> (5 - gg.lshift)) * gg.n_slices;     // rpi == 32 >> lshift
    if (item < n_items) { c.s = s; c.m0 = m0; c.rows = rows; return true; }
    item -= n_items;
    ++it;
  }
}

// An item holds rpi * (k + 1) ids, flat index f = sub_row * (k + 1) + j (j == k: the self id); a lane owns f = lane and
// f = lane + 32.  The (sub_row, j) split of those two slots only depends on the segment's k: it is cached in
// one byte per slot, (sub_row << 6) | j (0xFF: slot unused; rpi <= 4 and k < 64), both slots packed into one register
// and recomputed when the warp crosses into another segment, so the steady-state loop has no integer division.
__device__ __forceinline__ uint32_t decompose_slots(const GatherGeom& gg, int k, int lane) {
  const int k1 = k + 1;
  const int n_ids = gg.rpi * k1;
  uint32_t dec = 0;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int f = lane + 32 * u;
    const int sr = f / k1;
    dec |= (f < n_ids ? (uint32_t)((sr << 6) | (f - sr * k1)) : 0xFFu) << (8 * u);
  }
  return dec;
}

__device__ __forceinline__ void load_item_ids(const SageParams& p, const GatherGeom& gg, const TileCtx& c, int item, bool need_self,
                                              uint32_t dec, int64_t (&ids)[2]) {
  const SageSeg& sg = p.seg[c.s];
  const int rg = gg.n_slices == 1 ? item : item / gg.n_slices;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    ids[u] = -1;
    const int d = (int)((dec >> (8 * u)) & 0xFFu);
    if (d != 0xFF) {
      const int j = d & 63;
      const int r = rg * gg.rpi + (d >> 6);
      if (r < c.rows) {
        const int64_t m = c.m0 + r;
        if (j < sg.k) ids[u] = sg.nbr_vids ? __ldg(sg.nbr_vids + m * sg.k + j) : sg.nbr_base + m * sg.k + j;
        else if (need_self) ids[u] = sg.self_vids ? __ldg(sg.self_vids + m) : sg.self_base + m;
      }
    }
  }
}

__device__ __forceinline__ void write_item_ptrs(const SageParams& p, const TileCtx& c, const int64_t (&ids)[2], uint32_t dec,
                                                int lane, const char** scr, uint32_t self_row_bytes, uint32_t nbr_row_bytes) {
  const int k = p.seg[c.s].k;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int d = (int)((dec >> (8 * u)) & 0xFFu);
    if (d != 0xFF) {
      scr[lane + 32 * u] = (d & 63) < k ? vid_ptr(p.tnbr, ids[u], p.wshift_nbr, nbr_row_bytes, p.zero_row)
                                             : vid_ptr(p.tself, ids[u], p.wshift_self, self_row_bytes, p.zero_row);
    }
  }
}

// U  = neighbour row loads kept in flight per lane (one batch)
// DT = storage dtype of the self AND neighbour tables (0 fp32, 1 bf16)
//
// Gather decomposition: a row needs LPR = kp / VEC lanes (16-byte chunk per lane); when LPR < 32 a
// warp processes 32/LPR rows side by side (sub-warp groups), when LPR > 32 the row is cut into
// 32-lane slices.  Per item a lane group issues ALL self + neighbour chunk loads of a batch
// unconditionally (masked lanes read a zero row), reduces in fp32 and writes bf16 into the SW128
// A tile.
template <int U, int DT, int EW, int GW>
__global__ void __launch_bounds__((EW + 1 + GW) * 32, 1) sage_persist_kernel(const __grid_constant__ SageParams p) {
  constexpr int kEpiWarps = EW, kMmaWarp = EW, kGatherWarp0 = EW + 1, kGatherWarps = GW;
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[(size_t)blockIdx.x * 64 + 60] = (long long)gt;
    p.dbg[(size_t)blockIdx.x * 64 + 61] = clock64();
  }
  // keep every pointer derived from the __shared__ symbol by plain integer offsets so that the
  // compiler emits LDS/STS (a uintptr_t round trip would demote them to generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int k_total = p.kp_self + p.kp_nbr;
  const int nkb = k_total >> 6;
  uint8_t* sA = smem;
  const uint32_t a_buf_bytes = (uint32_t)nkb * (kTileM * 128);
  uint8_t* sW = sA + (size_t)kABufs * a_buf_bytes;
  const uint32_t w_kb_bytes = (uint32_t)p.N * 128u;
  const char** sScr = reinterpret_cast<const char**>(sW + (size_t)nkb * w_kb_bytes);   // [kGatherWarps][kScrCap]
  float* sBias = reinterpret_cast<float*>(sScr + kMaxGatherWarps * kScrCap);           // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    umma::mbar_init(bars + kBarW, 1);
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(bars + kBarAFull + i, kGatherWarps);
      umma::mbar_init(bars + kBarAFree + i, 1);
      umma::mbar_init(bars + kBarTFull + i, 1);
      umma::mbar_init(bars + kBarTFree + i, kEpiWarps);
    }
    umma::fence_barrier_init();
  }
  const int img = (int)blockIdx.x % p.n_imgs;
  // (multi-image launches: image `img` owns the output columns [img * N, img * N + N) and the matching bias slice)
  if (tid < 256) sBias[tid] = (p.bias && tid < p.n_out && img * p.N + tid < p.bias_len) ? __ldg(p.bias + img * p.N + tid) : 0.f;
  if (warp == kMmaWarp) {
    umma::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tile_first = (int)blockIdx.x / p.n_imgs, tile_stride = (int)gridDim.x / p.n_imgs;
  if (warp == 0) GLB_DBG(0);

  if (warp < kEpiWarps) {
    // =================================================================== epilogue warps
    int it = 0;
    for (int t = tile_first; t < p.total_tiles; t += tile_stride, ++it) {
      const int acc = it & 1;
      umma::mbar_wait(bars + kBarTFull + acc, (uint32_t)((it >> 1) & 1));
      umma::tc_fence_after();
      if (warp == 0 && it < 4) GLB_DBG(4 + it * 2);
      const int s = find_seg(p, t);
      SageSeg sg = p.seg[s];
      if (sg.out) sg.out += (size_t)img * p.N * (p.out_bf16 ? 2 : 4);
      const int m0 = (t - sg.tile0) * sg.R;
      const int rows = min(sg.R, sg.M - m0);
      sage_epilogue_tile(p, sg, sBias, tmem_base + (uint32_t)(acc * p.N), m0, rows, warp, lane, kEpiWarps);
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bars + kBarTFree + acc);
      if (p.ce_labels != nullptr) {
        __threadfence_block();
        asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");     // logits of the tile are written
        sage_ce_tile(p, sg, m0, rows, warp, lane, kEpiWarps);
      }
      if (warp == 0 && it < 4) GLB_DBG(5 + it * 2);
    }
  } else if (warp == kMmaWarp) {
    // =================================================================== MMA issuer (+ W image via TMA)
    if (lane == 0) {
      umma::mbar_arrive_expect_tx(bars + kBarW, (uint32_t)nkb * w_kb_bytes);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w_img) + (size_t)img * nkb * w_kb_bytes;
      for (int kb = 0; kb < nkb; ++kb)
        umma::bulk_g2s(sW + (size_t)kb * w_kb_bytes, src + (size_t)kb * w_kb_bytes, w_kb_bytes, bars + kBarW);
    }
    __syncwarp();
    umma::mbar_wait(bars + kBarW, 0);
    GLB_DBG(1);
    const uint32_t idesc = umma::make_idesc_bf16(kTileM, p.N);
    int it = 0;
    for (int t = tile_first; t < p.total_tiles; t += tile_stride, ++it) {
      const int acc = it & 1;
      umma::mbar_wait(bars + kBarAFull + acc, (uint32_t)((it >> 1) & 1));
      umma::mbar_wait(bars + kBarTFree + acc, (uint32_t)(((it >> 1) & 1) ^ 1));
      umma::tc_fence_after();
      if (it < 4) GLB_DBG(12 + it * 2);
      if (lane == 0) {
        const uint32_t d = tmem_base + (uint32_t)(acc * p.N);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t a_base = umma::smem_u32(sA + (size_t)acc * a_buf_bytes + (size_t)kb * (kTileM * 128));
          const uint32_t b_base = umma::smem_u32(sW + (size_t)kb * w_kb_bytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma::mma_bf16_ss(d, umma::make_desc_sw128(a_base + k4 * 32), umma::make_desc_sw128(b_base + k4 * 32), idesc,
                              (kb | k4) ? 1u : 0u);
        }
        umma::mma_commit(bars + kBarAFree + acc);     // this A buffer may be overwritten
        umma::mma_commit(bars + kBarTFull + acc);     // the accumulator may be drained
      }
      __syncwarp();
      if (it < 4) GLB_DBG(13 + it * 2);
    }
  } else {
    // =================================================================== gather warps
    constexpr int VEC = Chunk<DT>::kVec;
    const int gw = warp - kGatherWarp0;
    const int kp = p.kp_nbr > 0 ? p.kp_nbr : p.kp_self;
    const int d_self = p.tself.dim, d_nbr = p.tnbr.dim;
    const int lanes_row = kp / VEC;                       // 16-byte chunks per (half) row: 8..128
    const int lpr = lanes_row < 32 ? lanes_row : 32;      // lanes per row inside a warp
    GatherGeom gg;
    gg.lshift = 31 - __clz(lpr);                          // log2(lpr)
    gg.rpi = 32 >> gg.lshift;                             // rows per item
    gg.n_slices = lanes_row > 32 ? lanes_row >> 5 : 1;
    gg.sub = lane >> gg.lshift;                           // which row of the item this lane works on
    gg.lig = lane & (lpr - 1);                            // lane index inside its row group
    gg.tile_first = tile_first; gg.tile_stride = tile_stride;
    const bool has_self = p.kp_self > 0;
    const bool has_nbr = p.kp_nbr > 0;
    const bool need_self = has_self || p.mode == kGcnMean;
    constexpr uint32_t kEsz = DT == 0 ? 4u : DT == 1 ? 2u : 1u;
    const uint32_t nbr_row_bytes = (uint32_t)p.tnbr.stride * kEsz;
    const uint32_t self_row_bytes = (uint32_t)p.tself.stride * kEsz;
    const int soff_self = fp8_scale_offset(d_self), soff_nbr = fp8_scale_offset(d_nbr);     // fp8 rows: where the block scales start
    const char** scr = sScr + gw * kScrCap;
    const int n_tiles_cta = (p.total_tiles - tile_first + tile_stride - 1) / tile_stride;

    int arrived = 0;                                      // tiles this warp has already arrived for on a_full
    // arrive for every tile < upto (tiles in which this warp has no item left); an arrival for tile j (A buffer
    // j & 1) is only legal once the MMA of tile j-2 has been committed, otherwise it would be counted in that
    // tile's phase (a fresh barrier passes the parity-1 wait, so tiles 0 and 1 need no special case)
    auto arrive_upto = [&](int upto) {
      while (arrived < upto) {
        umma::mbar_wait(bars + kBarAFree + (arrived & 1), (uint32_t)(((arrived >> 1) & 1) ^ 1));
        if (lane == 0) umma::mbar_arrive(bars + kBarAFull + (arrived & 1));
        ++arrived;
      }
    };

    int it = 0, item = gw;
    TileCtx cur;
    bool valid = seek_item(p, gg, it, item, cur);
    uint32_t dec = 0xFFFFu;                               // id-slot split of the segment `dec_seg`
    int dec_seg = -1;
    if (valid) {
      int64_t ids[2];
      dec_seg = cur.s;
      dec = decompose_slots(gg, p.seg[cur.s].k, lane);
      load_item_ids(p, gg, cur, item, need_self, dec, ids);
      write_item_ptrs(p, cur, ids, dec, lane, scr, self_row_bytes, nbr_row_bytes);
      __syncwarp();
    }
    if (gw == 0) GLB_DBG(28);
    while (valid) {
      // ---- prefetch the ids of this warp's NEXT item (their latency hides behind the row loads below)
      int nit = it, nitem = item + kGatherWarps;
      TileCtx nxt;
      const bool nvalid = seek_item(p, gg, nit, nitem, nxt);
      int64_t nids[2];
      if (nvalid) {
        // (the current item's pointers are already staged: `dec` only serves the NEXT item from here on)
        if (nxt.s != dec_seg) { dec = decompose_slots(gg, p.seg[nxt.s].k, lane); dec_seg = nxt.s; }   // rare
        load_item_ids(p, gg, nxt, nitem, need_self, dec, nids);
      }
      // ---- current item
      const SageSeg& sg = p.seg[cur.s];
      const int k = sg.k;
      const int rg = gg.n_slices == 1 ? item : item / gg.n_slices;
      const int sl = gg.n_slices == 1 ? 0 : item - rg * gg.n_slices;
      const int r = rg * gg.rpi + gg.sub;
      const bool row_ok = r < cur.rows;
      const int chunk = gg.lig + 32 * sl;                 // this lane's 16-byte chunk of the row
      const int f0 = chunk * VEC;                         // first feature of the chunk
      float acc[VEC], sv[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) { acc[i] = 0.f; sv[i] = 0.f; }
      if (row_ok) {
        // lanes whose chunk lies beyond the real feature width skip the loads altogether (their A
        // columns are the zero K-padding); the others issue self + U neighbour loads back to back
        const char* const* ptrs = scr + gg.sub * (k + 1);
        const bool self_ld = need_self && f0 < d_self;
        Chunk<DT> sraw;
        if (self_ld) load_chunk<DT>(sraw, ptrs[k], chunk, soff_self);
        if (has_nbr && f0 < d_nbr) {
          int j0 = 0;
          for (; j0 + U <= k; j0 += U) {                  // full batches: U unconditional loads, then U adds
            Chunk<DT> raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) load_chunk<DT>(raw[u], ptrs[j0 + u], chunk, soff_nbr);
#pragma unroll
            for (int u = 0; u < U; ++u) raw[u].add_to(acc);
          }
          if (j0 < k) {                                   // remainder batch (k % U rows)
            Chunk<DT> raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) load_chunk<DT>(raw[u], ptrs[j0 + u < k ? j0 + u : k - 1], chunk, soff_nbr);
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (j0 + u < k) raw[u].add_to(acc);
          }
        }
        if (self_ld) sraw.add_to(sv);
        if (f0 + VEC > d_self || f0 + VEC > d_nbr) {       // only the chunk that straddles the real width needs masking
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            if (f0 + i >= d_self) sv[i] = 0.f;
            if (f0 + i >= d_nbr) acc[i] = 0.f;
          }
        }
      }
      // ---- A buffer it & 1 may only be written once the MMA of tile it-2 has consumed it
      arrive_upto(it);
      umma::mbar_wait(bars + kBarAFree + (it & 1), (uint32_t)(((it >> 1) & 1) ^ 1));
      uint8_t* sAt = sA + (size_t)(it & 1) * a_buf_bytes;
      if (row_ok) {
        const int m = cur.m0 + r;
        __nv_bfloat16* a_row = (sg.a_save && img == 0) ? sg.a_save + (size_t)m * k_total : nullptr;
        if (has_self) put_chunk<VEC>(sAt, a_row, r, f0, sv);
        if (has_nbr) {
          float scale = 1.f;
          if (p.mode == kConcatMean) scale = k > 0 ? 1.f / (float)k : 0.f;
          else if (p.mode == kGcnMean) scale = 1.f / (float)(k + 1);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = (acc[i] + (p.mode == kGcnMean ? sv[i] : 0.f)) * scale;
          put_chunk<VEC>(sAt, a_row, r, p.kp_self + f0, acc);
        }
      }
      if (!nvalid || nit != it) {                         // last item of this warp in tile `it`
        umma::fence_proxy_async_smem();                   // generic-proxy st.shared -> visible to tcgen05 (async proxy)
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bars + kBarAFull + (it & 1));
        arrived = it + 1;
        if (gw == 0 && it < 4) GLB_DBG(30 + it * 3);
      }
      __syncwarp();                                       // every lane is done reading the scratch pointers
      if (nvalid) {
        write_item_ptrs(p, nxt, nids, dec, lane, scr, self_row_bytes, nbr_row_bytes);
        __syncwarp();
      }
      it = nit; item = nitem; cur = nxt; valid = nvalid;
    }
    arrive_upto(n_tiles_cta);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) GLB_DBG(2);
  if (warp == kMmaWarp) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[(size_t)blockIdx.x * 64 + 62] = (long long)gt;
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

// dynamic shared memory of the persistent kernel: A tile + W image + per-warp pointer scratch + bias + control + alignment slack
int64_t sage_smem_bytes(int64_t k_total, int64_t N) {
  return 1024 + kABufs * (k_total / 64) * (kTileM * 128) + (k_total / 64) * N * 128 + kMaxGatherWarps * kScrCap * 8 + 256 * 4 + (kNumBars + 2) * 8;
}

int sm_count() {
  static int n = 0;
  if (n == 0) n = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  return n;
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

static long long* g_dbg_ptr = nullptr;
static int g_max_ctas = 0;     // 0 = all SMs; engines cap a launch that should co-run with another whole-SM kernel
void sage_set_max_ctas(int64_t n) { g_max_ctas = (int)n; }
static int usable_sms() { const int s = sm_count(); return g_max_ctas > 0 ? std::min(s, g_max_ctas) : s; }
void sage_set_debug_trace(const c10::optional<at::Tensor>& t) {
  g_dbg_ptr = (t.has_value() && t->defined()) ? reinterpret_cast<long long*>(t->data_ptr<int64_t>()) : nullptr;
}

static int log2_or_neg(int w) { int s = 0; while ((1 << s) < w) ++s; return (1 << s) == w ? s : -1; }

static const char* zero_row_for(const at::TensorOptions& opts, int dev) {
  static std::vector<at::Tensor> zero_rows(64);
  if (!zero_rows[dev].defined()) zero_rows[dev] = at::zeros({1024}, opts.dtype(at::kFloat));
  return reinterpret_cast<const char*>(zero_rows[dev].data_ptr());
}

// Balanced tiling: every tile should cost about the same number of row fetches ((k + 1) per destination
// row, + 2 for the epilogue / A stores) and the tile count should just fill `waves` tiles per SM.
static void plan_tiles(SageParams& p, int64_t rows_forced) {
  const int sms = std::max(1, usable_sms() / p.n_imgs);
  int64_t n128 = 0;
  double total_cost = 0.0;
  for (int s = 0; s < p.nseg; ++s) {
    n128 += (p.seg[s].M + kTileM - 1) / kTileM;
    total_cost += (double)p.seg[s].M * (p.seg[s].k + 3);
  }
  const int64_t waves = std::max<int64_t>(1, (n128 + sms - 1) / sms);
  const double target = total_cost / (double)(sms * waves);
  int tiles = 0;
  for (int s = 0; s < p.nseg; ++s) {
    SageSeg& sg = p.seg[s];
    int64_t R = (int64_t)std::ceil(target / (double)(sg.k + 3));
    R = (R + 7) / 8 * 8;
    if (rows_forced > 0) R = (rows_forced + 7) / 8 * 8;
    R = std::min<int64_t>(kTileM, std::max<int64_t>(8, R));
    sg.R = (int)R;
    sg.tile0 = tiles;
    tiles += sg.M > 0 ? (int)((sg.M + R - 1) / R) : 0;
  }
  p.total_tiles = tiles;
}

static int g_variant = 2;      // role layout of the persistent kernel (see the table at the top): (4, 19) measured fastest
                               // at 1 and 2 GPUs (profiles/r2_role_layouts.txt); sage_set_variant / GLB_SAGE_VARIANT override
void sage_set_variant(int64_t v) { TORCH_CHECK(v >= 0 && v <= 2, "variant 0..2"); g_variant = (int)v; }

static void launch_persist(SageParams& p, size_t smem, cudaStream_t stream) {
  if (p.total_tiles == 0) return;
  const unsigned grid = (unsigned)(p.n_imgs * std::min<int>(std::max(1, usable_sms() / p.n_imgs), p.total_tiles));
  int kmax = 1;
  for (int s = 0; s < p.nseg; ++s) kmax = std::max(kmax, p.seg[s].k);
  const int dt = p.tnbr.dtype;
#define LAUNCH(UU, DD, EW, GW)                                                                             \
  do {                                                                                                     \
    static bool attr_done = false;                                                                         \
    if (!attr_done) {                                                                                      \
      C10_CUDA_CHECK(cudaFuncSetAttribute(sage_persist_kernel<UU, DD, EW, GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit)); \
      attr_done = true;                                                                                    \
    }                                                                                                      \
    sage_persist_kernel<UU, DD, EW, GW><<<grid, (EW + 1 + GW) * 32, smem, stream>>>(p);                    \
  } while (0)
#define LAUNCH_DT(UU, EW, GW) do { if (dt == 0) LAUNCH(UU, 0, EW, GW); else LAUNCH(UU, 1, EW, GW); } while (0)
#define LAUNCH_DT3(UU, EW, GW) do { if (dt == 2) LAUNCH(UU, 2, EW, GW); else LAUNCH_DT(UU, EW, GW); } while (0)
  TORCH_CHECK(dt != 2 || g_variant == 2, "fp8 feature rows are built for the default role layout only");
  if (g_variant == 1) {
    // 512 threads x 128 registers: a whole neighbour list (or half of a long one) in flight per lane
    const int u = kmax <= 5 ? 5 : kmax <= 10 ? 10 : 13;
    if (u == 5) LAUNCH_DT(5, 4, 11); else if (u == 10) LAUNCH_DT(10, 4, 11); else LAUNCH_DT(13, 4, 11);
  } else if (g_variant == 2) {
    const int u = kmax <= 4 ? 4 : (kmax % 5 == 0 || kmax > 12) ? 5 : 6;
    if (u == 4) LAUNCH_DT3(4, 4, 19); else if (u == 5) LAUNCH_DT3(5, 4, 19); else LAUNCH_DT3(6, 4, 19);
  } else {
    const int u = kmax <= 4 ? 4 : (kmax % 5 == 0 || kmax > 12) ? 5 : 6;
    if (u == 4) LAUNCH_DT(4, 8, 23); else if (u == 5) LAUNCH_DT(5, 8, 23); else LAUNCH_DT(6, 8, 23);
  }
#undef LAUNCH_DT3
#undef LAUNCH_DT
#undef LAUNCH
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// Multi-segment entry point.  All segments share the tables, the weight image, the bias and the output
// layout; segment i reads ids (self_vids[i], nbr_vids[i]) - or identity ids self_base[i] + m /
// nbr_base[i] + m*k + j when the tensors are None - and writes outs[i] ([M_i, n_out]) and optionally
// a_saves[i] ([M_i, K_total] bf16, the [self || agg] rows for the backward pass).
// ce = optional fused softmax cross-entropy on the output rows of segment 0:
//   [labels_table(int64), seeds(int64)|undefined, loss(fp32 >=1), dlogits(bf16 [M, stride]), dbias(fp32)|undefined]
void sage_fused_multi(const at::Tensor& tself_desc, const at::Tensor& tnbr_desc,
                      const std::vector<c10::optional<at::Tensor>>& self_vids,
                      const std::vector<c10::optional<at::Tensor>>& nbr_vids,
                      const std::vector<int64_t>& self_base, const std::vector<int64_t>& nbr_base,
                      const std::vector<int64_t>& Ms, const std::vector<int64_t>& ks,
                      const std::vector<at::Tensor>& outs, const std::vector<c10::optional<at::Tensor>>& a_saves,
                      int64_t mode, const at::Tensor& w_img, const c10::optional<at::Tensor>& bias, int64_t N,
                      int64_t n_out, bool relu, bool out_bf16, int64_t rows_forced,
                      const std::vector<c10::optional<at::Tensor>>& ce, int64_t ce_world, int64_t n_imgs) {
  TORCH_CHECK(w_img.is_cuda() && w_img.scalar_type() == at::kBFloat16, "w_img must be CUDA bf16");
  c10::cuda::CUDAGuard guard(w_img.device());
  const int nseg = (int)Ms.size();
  TORCH_CHECK(nseg >= 1 && nseg <= kMaxSegs, "1..", kMaxSegs, " segments per launch");
  TORCH_CHECK((int)ks.size() == nseg && (int)outs.size() == nseg && (int)a_saves.size() == nseg &&
              (int)self_vids.size() == nseg && (int)nbr_vids.size() == nseg && (int)self_base.size() == nseg &&
              (int)nbr_base.size() == nseg, "per-segment argument lists must have equal lengths");
  SageParams p;
  std::memset(&p, 0, sizeof(p));
  p.tself = table_from_desc(tself_desc);
  p.tnbr = table_from_desc(tnbr_desc);
  p.nseg = nseg;
  p.mode = (int)mode;
  bool any_nbr = false;
  for (int s = 0; s < nseg; ++s) any_nbr = any_nbr || ks[s] > 0;
  p.kp_self = (mode == kGcnMean) ? 0 : pad_k(p.tself.dim);
  p.kp_nbr = any_nbr ? pad_k(p.tnbr.dim) : 0;
  if (mode == kGcnMean) TORCH_CHECK(p.tself.dim == p.tnbr.dim, "gcn mode needs equal dims");
  const int k_total = p.kp_self + p.kp_nbr;
  TORCH_CHECK(k_total >= 64, "empty layer");
  TORCH_CHECK(N % 32 == 0 && N >= 32 && N <= 256, "padded N must be a multiple of 32 in [32, 256]");
  TORCH_CHECK(n_out <= N && n_out >= 1);
  TORCH_CHECK(n_imgs >= 1 && n_imgs <= 8 && (n_imgs == 1 || n_out == N), "bad image count");
  p.n_imgs = (int)n_imgs;
  TORCH_CHECK(w_img.numel() == n_imgs * (int64_t)k_total * N, "weight image size mismatch: expected ", n_imgs * k_total * N, " got ", w_img.numel());
  TORCH_CHECK((reinterpret_cast<uintptr_t>(w_img.data_ptr()) & 15) == 0, "weight image must be 16 B aligned");
  const size_t smem = (size_t)sage_smem_bytes(k_total, N);
  TORCH_CHECK(smem <= kSmemLimit, "tile does not fit in shared memory (", smem, " B); use the unfused path");
  TORCH_CHECK(p.tself.dtype == p.tnbr.dtype, "fused SAGE layer needs self / neighbour tables of the same storage dtype");
  TORCH_CHECK(p.kp_self == 0 || p.kp_nbr == 0 || p.kp_self == p.kp_nbr, "fused SAGE layer needs equally padded self / neighbour dims");
  TORCH_CHECK((p.tself.stride * table_esize(p.tself.dtype)) % 16 == 0 && (p.tnbr.stride * table_esize(p.tnbr.dtype)) % 16 == 0,
              "table rows must be 16-byte aligned (stride multiple of 4 fp32 / 8 bf16 elements)");
  {
    const int vec = p.tnbr.dtype == 0 ? 4 : 8;
    const int lanes_row = (p.kp_nbr > 0 ? p.kp_nbr : p.kp_self) / vec;
    const int rpi = lanes_row < 32 ? 32 / lanes_row : 1;
    for (int s = 0; s < nseg; ++s)
      TORCH_CHECK(rpi * (ks[s] + 1) <= kScrCap, "fan-out too large for the fused kernel's pointer scratch: k <= ", kScrCap / rpi - 1);
  }
  std::vector<at::Tensor> keep;
  at::Tensor b;
  if (bias.has_value() && bias->defined()) {
    b = bias->contiguous();
    TORCH_CHECK(b.is_cuda() && b.scalar_type() == at::kFloat && (n_imgs > 1 || b.numel() >= n_out), "bias must be fp32 [>= n_out]");
    p.bias = b.data_ptr<float>();
    p.bias_len = (int)b.numel();
  }
  p.out_stride = -1;
  for (int s = 0; s < nseg; ++s) {
    SageSeg& sg = p.seg[s];
    sg.M = (int)Ms[s]; sg.k = (int)ks[s];
    sg.self_base = self_base[s]; sg.nbr_base = nbr_base[s];
    if (self_vids[s].has_value() && self_vids[s]->defined()) {
      at::Tensor t = self_vids[s]->contiguous(); check_cuda_i64(t, "self_vids");
      TORCH_CHECK(t.numel() == Ms[s]); keep.push_back(t); sg.self_vids = t.data_ptr<int64_t>();
    }
    if (nbr_vids[s].has_value() && nbr_vids[s]->defined()) {
      at::Tensor t = nbr_vids[s]->contiguous(); check_cuda_i64(t, "nbr_vids");
      TORCH_CHECK(t.numel() == Ms[s] * ks[s]); keep.push_back(t); sg.nbr_vids = t.data_ptr<int64_t>();
    }
    const at::Tensor& out = outs[s];
    if (out.defined()) {
      TORCH_CHECK(out.is_cuda() && out.dim() == 2 && out.size(0) == Ms[s] && out.size(1) == n_out && out.stride(1) == 1,
                  "outs[i] must be [M_i, n_out] with unit inner stride");
      TORCH_CHECK(out.scalar_type() == (out_bf16 ? at::kBFloat16 : at::kFloat), "out dtype mismatch");
      TORCH_CHECK(p.out_stride < 0 || p.out_stride == out.stride(0), "all outputs must share the row stride");
      p.out_stride = out.stride(0);
      sg.out = reinterpret_cast<char*>(out.data_ptr());
    }
    if (a_saves[s].has_value() && a_saves[s]->defined()) {
      const at::Tensor& a = *a_saves[s];
      TORCH_CHECK(a.scalar_type() == at::kBFloat16 && a.dim() == 2 && a.size(0) == Ms[s] && a.size(1) == k_total &&
                  a.stride(1) == 1 && a.stride(0) == k_total, "a_saves[i] must be a row-contiguous bf16 [M_i, K_total] view");
      sg.a_save = reinterpret_cast<__nv_bfloat16*>(a.data_ptr());
    }
  }
  if (p.out_stride < 0) p.out_stride = n_out;
  p.w_img = w_img.data_ptr();
  p.N = (int)N; p.n_out = (int)n_out; p.relu = relu ? 1 : 0; p.out_bf16 = out_bf16 ? 1 : 0;
  const int need_cols = 2 * (int)N;
  p.tmem_cols = need_cols <= 32 ? 32 : need_cols <= 64 ? 64 : need_cols <= 128 ? 128 : need_cols <= 256 ? 256 : 512;
  p.wshift_self = log2_or_neg(p.tself.world);
  p.wshift_nbr = log2_or_neg(p.tnbr.world);
  p.zero_row = zero_row_for(w_img.options(), w_img.get_device());
  if (!ce.empty()) {
    TORCH_CHECK(ce.size() == 5 && nseg == 1 && N <= 64, "fused CE needs one segment and n_out <= 64");
    TORCH_CHECK(p.seg[0].out != nullptr && !out_bf16, "fused CE needs the fp32 logits output (its second phase re-reads it)");
    const at::Tensor& labels = *ce[0];
    TORCH_CHECK(labels.is_cuda() && labels.scalar_type() == at::kLong);
    p.ce_labels = labels.data_ptr<int64_t>();
    if (ce[1].has_value() && ce[1]->defined()) {
      TORCH_CHECK(ce[1]->scalar_type() == at::kLong && ce[1]->numel() == Ms[0] && ce[1]->is_contiguous());
      p.ce_seeds = ce[1]->data_ptr<int64_t>();
    } else {
      TORCH_CHECK(labels.numel() >= Ms[0]);
    }
    p.ce_world = (int)std::max<int64_t>(ce_world, 1);
    const at::Tensor& loss = *ce[2];
    TORCH_CHECK(loss.is_cuda() && loss.scalar_type() == at::kFloat && loss.numel() >= 1);
    p.ce_loss = loss.data_ptr<float>();
    const at::Tensor& dl = *ce[3];
    TORCH_CHECK(dl.is_cuda() && dl.scalar_type() == at::kBFloat16 && dl.dim() == 2 && dl.size(0) == Ms[0] && dl.stride(1) == 1 &&
                dl.stride(0) >= n_out, "dlogits must be bf16 [M, >= n_out]");
    p.ce_dlogits = reinterpret_cast<__nv_bfloat16*>(dl.data_ptr());
    p.ce_dl_stride = (int)dl.stride(0);
    if (ce[4].has_value() && ce[4]->defined()) {
      TORCH_CHECK(ce[4]->scalar_type() == at::kFloat && ce[4]->numel() >= n_out);
      p.ce_dbias = ce[4]->data_ptr<float>();
    }
    p.ce_inv_b = 1.f / (float)Ms[0];
  }
  p.dbg = g_dbg_ptr;
  plan_tiles(p, rows_forced);
  launch_persist(p, smem, at::cuda::getCurrentCUDAStream());
}

// single-segment convenience wrapper (autograd op in ops/sage.py)
std::vector<at::Tensor> sage_fused_forward(const at::Tensor& tself_desc,
                                           const c10::optional<at::Tensor>& self_vids,
                                           const at::Tensor& tnbr_desc,
                                           const c10::optional<at::Tensor>& nbr_vids, int64_t M,
                                           int64_t k, int64_t mode, const at::Tensor& w_img,
                                           const c10::optional<at::Tensor>& bias, int64_t N,
                                           int64_t n_out, bool relu, bool out_bf16, bool save_a, int64_t rows_per_cta,
                                           const c10::optional<at::Tensor>& out_buf,
                                           const c10::optional<at::Tensor>& a_buf) {
  c10::cuda::CUDAGuard guard(w_img.device());
  auto opts = w_img.options();
  at::Tensor out = out_buf.has_value() && out_buf->defined() ? *out_buf
                                                             : at::empty({M, n_out}, opts.dtype(out_bf16 ? at::kBFloat16 : at::kFloat));
  at::Tensor a_save;
  if (save_a) {
    const TableView ts = table_from_desc(tself_desc), tn = table_from_desc(tnbr_desc);
    const int64_t k_total = ((mode == kGcnMean) ? 0 : pad_k(ts.dim)) + (k > 0 ? pad_k(tn.dim) : 0);
    a_save = a_buf.has_value() && a_buf->defined() ? *a_buf : at::empty({M, k_total}, opts.dtype(at::kBFloat16));
  }
  if (M == 0) return {out, a_save};
  sage_fused_multi(tself_desc, tnbr_desc, {self_vids}, {nbr_vids}, {0}, {0}, {M}, {k}, {out},
                   {save_a ? c10::optional<at::Tensor>(a_save) : c10::nullopt}, mode, w_img, bias, N, n_out, relu, out_bf16,
                   rows_per_cta, {}, 1, 1);
  return {out, a_save};
}

// K7 standalone: out = act(A . W^T + b) for a dense bf16 A [M, K] (K multiple of 64, <= 512; N <= 256)
// on the same persistent tcgen05 / TMEM kernel (the "gather" degenerates to staging A's rows).
at::Tensor tc_linear_forward(const at::Tensor& a, const at::Tensor& w_img, const c10::optional<at::Tensor>& bias,
                             int64_t N, int64_t n_out, bool relu, bool out_bf16) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && a.dim() == 2 && a.stride(1) == 1,
              "A must be a CUDA bf16 [M, K] matrix with unit inner stride");
  const int64_t M = a.size(0), K = a.size(1);
  TORCH_CHECK(K % 8 == 0 && K >= 8 && K <= 512 && a.stride(0) % 8 == 0, "K must be a multiple of 8 in [8, 512]");
  c10::cuda::CUDAGuard guard(a.device());
  auto out = at::empty({M, n_out}, a.options().dtype(out_bf16 ? at::kBFloat16 : at::kFloat));
  if (M == 0) return out;
  std::vector<int64_t> d(4 + 2 * kMaxWorld, 0);
  d[0] = 1; d[1] = K; d[2] = a.stride(0); d[3] = 1; d[4] = M; d[4 + kMaxWorld] = reinterpret_cast<int64_t>(a.data_ptr());
  at::Tensor desc = at::from_blob(d.data(), {(int64_t)d.size()}, at::kLong).clone();
  sage_fused_multi(desc, desc, {c10::nullopt}, {c10::nullopt}, {0}, {0}, {M}, {0}, {out}, {c10::nullopt}, kConcatMean, w_img, bias,
                   N, n_out, relu, out_bf16, 0, {}, 1, 1);
  return out;
}

}  // namespace glb
