// Peer-memory runtime + K8 (one-shot gradient all-reduce fused with cast/scale)
// + fused Adam over the flat parameter buffer.
//
// The reference moves every byte through gRPC/protobuf
// (graphlearn/src/service/dist/grpc_channel.cc:81-90) and relies on PyTorch DDP
// for the gradient all-reduce (graphlearn/examples/pytorch/gcn/train.py:192-193).
// Here every rank cudaMalloc's "symmetric" buffers, exchanges CUDA IPC handles
// once through torch.distributed and afterwards kernels simply dereference
// peer pointers over NVLink / NVSwitch.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <cstring>
#include "host_utils.h"

namespace glb {

// ---------------------------------------------------------------------------
// symmetric heap / CUDA IPC
// ---------------------------------------------------------------------------
int64_t symm_alloc(int64_t nbytes, int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  void* p = nullptr;
  size_t n = (size_t)((nbytes + 511) / 512 * 512);
  if (n == 0) n = 512;
  C10_CUDA_CHECK(cudaMalloc(&p, n));
  C10_CUDA_CHECK(cudaMemset(p, 0, n));
  C10_CUDA_CHECK(cudaDeviceSynchronize());
  return reinterpret_cast<int64_t>(p);
}

void symm_free(int64_t ptr, int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  C10_CUDA_CHECK(cudaFree(reinterpret_cast<void*>(ptr)));
}

py::bytes ipc_get_handle(int64_t ptr, int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  cudaIpcMemHandle_t h;
  C10_CUDA_CHECK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(ptr)));
  return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
}

int64_t ipc_open_handle(const std::string& handle, int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  C10_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  return reinterpret_cast<int64_t>(p);
}

void ipc_close_handle(int64_t ptr, int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  C10_CUDA_CHECK(cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)));
}

at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, int64_t dtype_code, int64_t device) {
  at::ScalarType st = dtype_code == 0 ? at::kFloat : dtype_code == 1 ? at::kBFloat16
                    : dtype_code == 2 ? at::kLong : dtype_code == 3 ? at::kInt : at::kByte;
  auto opts = at::TensorOptions().dtype(st).device(at::kCUDA, (c10::DeviceIndex)device);
  return at::from_blob(reinterpret_cast<void*>(ptr), sizes, opts);
}

// ---------------------------------------------------------------------------
// K8: one-shot all-reduce over peer memory
// ---------------------------------------------------------------------------
constexpr int kArMaxBlocks = 64;

struct AllReduceParams {
  PeerTableMut stage;        // per rank: float [2][n_pad]   (double buffered staging, symmetric)
  PeerTableMut flags;        // per rank: uint32 [kArMaxBlocks][kMaxWorld]  (symmetric, zero-init)
  unsigned long long* epochs;  // local: [kArMaxBlocks]
  int* error_flag;           // local: set to 1 on barrier timeout
  float* grad;               // local flat gradient (in/out)
  int64_t n;                 // elements (multiple of 4)
  int64_t n_pad;
  float scale;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512) allreduce_oneshot_kernel(const AllReduceParams p) {
  const int b = blockIdx.x;
  const unsigned long long e = p.epochs[b] + 1;      // this call's epoch (same on every rank)
  const int parity = (int)(e & 1);
  const int64_t per_block = ((p.n / 4 + gridDim.x - 1) / gridDim.x) * 4;
  const int64_t lo = (int64_t)b * per_block;
  const int64_t hi = min(p.n, lo + per_block);

  // phase 0: publish my slice (scaled) into my staging buffer
  float* my_stage = reinterpret_cast<float*>(p.stage.p[p.rank]) + (size_t)parity * p.n_pad;
  for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi; i += 4 * (int64_t)blockDim.x) {
    float4 g = *reinterpret_cast<const float4*>(p.grad + i);
    g.x *= p.scale; g.y *= p.scale; g.z *= p.scale; g.w *= p.scale;
    *reinterpret_cast<float4*>(my_stage + i) = g;
  }
  __threadfence_system();
  __syncthreads();

  // phase 1: block-level cross-GPU barrier (block b of every rank)
  if (threadIdx.x < p.world) {
    const int r = threadIdx.x;
    unsigned int* remote = reinterpret_cast<unsigned int*>(p.flags.p[r]) + b * kMaxWorld + p.rank;
    st_release_sys(remote, (unsigned int)e);
    const unsigned int* mine = reinterpret_cast<const unsigned int*>(p.flags.p[p.rank]) + b * kMaxWorld + r;
    long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - (unsigned int)e) < 0) {
      if (clock64() - t0 > 20000000000LL) { *p.error_flag = 1; break; }   // ~10 s watchdog
    }
  }
  __syncthreads();

  // phase 2: pull + reduce every rank's slice straight from peer HBM
  for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi; i += 4 * (int64_t)blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r) {
      if (r < p.world) {
        const float4* src = reinterpret_cast<const float4*>(
            reinterpret_cast<const float*>(p.stage.p[r]) + (size_t)parity * p.n_pad + i);
        float4 v = __ldcv(src);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    *reinterpret_cast<float4*>(p.grad + i) = acc;
  }
  if (threadIdx.x == 0) p.epochs[b] = e;
}

// desc (CPU int64): [rank, world, n_pad, stage_ptr[8], flags_ptr[8]]
void allreduce_oneshot(const at::Tensor& desc, const at::Tensor& grad, const at::Tensor& epochs,
                       const at::Tensor& error_flag, double scale) {
  TORCH_CHECK(desc.device().is_cpu() && desc.scalar_type() == at::kLong && desc.numel() == 3 + 2 * kMaxWorld);
  TORCH_CHECK(grad.is_cuda() && grad.scalar_type() == at::kFloat && grad.is_contiguous());
  TORCH_CHECK(epochs.is_cuda() && epochs.scalar_type() == at::kLong && epochs.numel() >= kArMaxBlocks);
  TORCH_CHECK(error_flag.is_cuda() && error_flag.scalar_type() == at::kInt);
  c10::cuda::CUDAGuard guard(grad.device());
  const int64_t* d = desc.data_ptr<int64_t>();
  AllReduceParams p;
  p.rank = (int)d[0]; p.world = (int)d[1]; p.n_pad = d[2];
  for (int r = 0; r < kMaxWorld; ++r) {
    p.stage.p[r] = reinterpret_cast<void*>(d[3 + r]);
    p.flags.p[r] = reinterpret_cast<void*>(d[3 + kMaxWorld + r]);
  }
  p.n = grad.numel();
  TORCH_CHECK(p.n % 4 == 0 && p.n <= p.n_pad, "flat grad must be padded to a multiple of 4 and fit the staging buffer");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(grad.data_ptr()) & 15) == 0);
  p.grad = grad.data_ptr<float>();
  p.epochs = reinterpret_cast<unsigned long long*>(epochs.data_ptr<int64_t>());
  p.error_flag = error_flag.data_ptr<int>();
  p.scale = (float)scale;
  // grid MUST be identical on every rank (slices are matched by block index)
  int blocks = (int)std::min<int64_t>(kArMaxBlocks, std::max<int64_t>(1, p.n / (4 * 512)));
  allreduce_oneshot_kernel<<<blocks, 512, 0, at::cuda::getCurrentCUDAStream()>>>(p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------
// fused Adam over a flat buffer; step count lives on the device (graph safe)
// ---------------------------------------------------------------------------
__global__ void step_advance_kernel(uint64_t* rng_state, long long* opt_step, uint64_t rng_inc) {
  if (rng_state) rng_state[1] += rng_inc;
  if (opt_step) opt_step[0] += 1;
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, int64_t n, const long long* __restrict__ step_ptr, float lr,
                 float beta1, float beta2, float eps, float weight_decay) {
  const float step = (float)step_ptr[0];
  const float bc1 = 1.f - powf(beta1, step);
  const float bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float wi = w[i];
    if (weight_decay != 0.f) gi += weight_decay * wi;
    float mi = beta1 * m[i] + (1.f - beta1) * gi;
    float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    w[i] = wi - step_size * mi / denom;
  }
}

void step_advance(const c10::optional<at::Tensor>& rng_state, const c10::optional<at::Tensor>& opt_step,
                  int64_t rng_inc) {
  uint64_t* rp = nullptr; long long* sp = nullptr;
  c10::Device dev(at::kCUDA, 0);
  if (rng_state.has_value()) { rp = reinterpret_cast<uint64_t*>(rng_state->data_ptr<int64_t>()); dev = rng_state->device(); }
  if (opt_step.has_value()) { sp = reinterpret_cast<long long*>(opt_step->data_ptr<int64_t>()); dev = opt_step->device(); }
  c10::cuda::CUDAGuard guard(dev);
  step_advance_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>(rp, sp, (uint64_t)rng_inc);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void adam_flat(const at::Tensor& w, const at::Tensor& g, const at::Tensor& m, const at::Tensor& v,
               const at::Tensor& step, double lr, double beta1, double beta2, double eps,
               double weight_decay) {
  TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.is_contiguous());
  TORCH_CHECK(g.numel() == w.numel() && m.numel() == w.numel() && v.numel() == w.numel());
  TORCH_CHECK(step.is_cuda() && step.scalar_type() == at::kLong);
  c10::cuda::CUDAGuard guard(w.device());
  int64_t n = w.numel();
  if (n == 0) return;
  int blocks = (int)std::min<int64_t>(148 * 4, (n + 255) / 256);
  adam_flat_kernel<<<blocks, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      w.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), n,
      reinterpret_cast<const long long*>(step.data_ptr<int64_t>()), (float)lr, (float)beta1, (float)beta2,
      (float)eps, (float)weight_decay);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace glb
