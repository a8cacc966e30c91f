// Peer-memory gradient exchange (data-parallel step): IPC arena management, the "ready" signal and the owner's
// reduce pass.  Replaces the NCCL all-reduce of the raw bf16 weight gradients: no extra SM-hungry kernels compete
// with the persistent GEMM, nothing waits for a free SM, and the all-gather half of the exchange is folded into the
// gradient-finishing pass (elementwise.cu: grad_finish_peer_kernel), which pulls each slice from its owner.
#include "../../include/ed2b200.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "peer.cuh"
#include "status.h"

namespace ed2 {
namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) peer_signal_kernel(PeerExchange x, const float* __restrict__ small_src) {
  __shared__ unsigned long long epoch_s;
  if (threadIdx.x == 0) {
    unsigned long long* w = peer::local_words(x);
    epoch_s = w[0] + 1;
    w[0] = epoch_s;
  }
  __syncthreads();
  // companion vector -> arena slot (epoch & 1): a peer that is still in its finishing pass of epoch e reads slot e & 1
  // while this rank may already publish epoch e + 1 into the other slot (it cannot reach e + 2 before every peer has
  // finished e: its own finishing pass of e + 1 waits for every owner's reduced flag of e + 1)
  float* dst = const_cast<float*>(x.small[x.rank]) + (epoch_s & 1ull) * x.small_n;
  for (long long i = threadIdx.x; i < x.small_n; i += kBlock) dst[i] = small_src[i];
  __syncthreads();  // orders every thread's `small` stores before the releasing threads (CTA scope, cumulative)
  // the bucket (written by the previous kernel) and `small` become visible to whoever acquires the flag
  if (threadIdx.x < x.world) {
    __threadfence_system();
    peer::st_release_sys(peer::ready_flags(x, threadIdx.x) + x.rank, epoch_s);
  }
}

// peer_signal_kernel without a companion vector: epoch += 1, ready flag on every rank.  One warp, <= 32 registers, so it
// fits beside a persistent GEMM CTA (see peer_wait_kernel).
__global__ void __maxnreg__(32) peer_signal_small_kernel(PeerExchange x) {
  unsigned long long* w = peer::local_words(x);
  unsigned long long epoch = 0;
  if (threadIdx.x == 0) {
    epoch = w[0] + 1;
    w[0] = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if (threadIdx.x < x.world) {
    __threadfence_system();
    peer::st_release_sys(peer::ready_flags(x, threadIdx.x) + x.rank, epoch);
  }
}

// publish the CURRENT epoch into flag section `which` (0: ready, 1: reduced) of every rank — used by the copy-engine
// exchange, where the data movement is a DMA the flag must follow in stream order
__global__ void peer_flag_kernel(PeerExchange x, int which) {
  if (threadIdx.x < x.world) {
    const unsigned long long epoch = peer::local_words(x)[0];
    __threadfence_system();
    peer::st_release_sys(x.flags[threadIdx.x] + which * x.world + x.rank, epoch);
  }
}

// Wait until flag section `which` of THIS rank carries the current epoch from every rank.  One warp, <= 32 registers:
// it fits beside a persistent GEMM CTA that owns the rest of the SM, so waiting for a peer never keeps a GEMM CTA from
// becoming resident (the heavy reduce / convert kernels are enqueued behind it and start only when their data is here).
__global__ void __maxnreg__(32) peer_wait_kernel(PeerExchange x, int which) {
  unsigned long long* lw = peer::local_words(x);
  const unsigned long long epoch = lw[0];
  if (threadIdx.x < x.world)
    peer::wait_flag(x.flags[x.rank] + which * x.world + threadIdx.x, epoch, lw + 2, 0x400 + which * 0x100 + threadIdx.x);
}

__device__ __forceinline__ void acc8(float (&a)[8], const uint4& v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    a[2 * j] += f.x;
    a[2 * j + 1] += f.y;
  }
}

template <int kWorld>  // 0: run-time world size
__global__ void __launch_bounds__(kBlock) peer_reduce_slice_kernel(PeerExchange x) {
  constexpr int kW = kWorld > 0 ? kWorld : kMaxPeers;
  const int world = kWorld > 0 ? kWorld : x.world;
  unsigned long long* lw = peer::local_words(x);
  const unsigned long long epoch = lw[0];
  if (threadIdx.x < world) peer::wait_flag(peer::ready_flags(x, x.rank) + threadIdx.x, epoch, lw + 2, 0x100 + threadIdx.x);
  __syncthreads();
  const long long lo = x.rank * x.chunk, hi = min(x.n, lo + x.chunk);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(x.reduced[x.rank]);
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  // two 16-byte vectors per trip: 2 * world NVLink loads in flight per thread (the kernel is latency-bound)
  for (long long i = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < hi; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < hi;
    uint4 v[kW], w[kW];
#pragma unroll
    for (int r = 0; r < kW; ++r)
      if (r < world) {
        v[r] = peer::ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(x.bucket[r]) + i);
        if (two) w[r] = peer::ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(x.bucket[r]) + i2);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < kW; ++r)  // fixed rank order: every rank gets the same bits
        if (r < world) acc8(a, h == 0 ? v[r] : w[r]);
      uint4 o;
      __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(a[2 * j], a[2 * j + 1]);
      *reinterpret_cast<uint4*>(out + (h == 0 ? i : i2)) = o;
    }
  }
  // last block out publishes the slice
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long t = atomicAdd(lw + 1, 1ull);
    if (t == gridDim.x - 1) {
      lw[1] = 0;
      __threadfence_system();  // every block's slice stores (observed through the counter) before the flags
      for (int q = 0; q < world; ++q) peer::st_release_sys(peer::reduced_flags(x, q) + x.rank, epoch);
    }
  }
}

// Plain finishing pass of the exchange for gradients that need no further arithmetic (deterministic kernels, or fp32
// gradients that were cast into the bucket): out[i] = float(reduced[owner(i)][i]), every slice read from its owner.
__global__ void __launch_bounds__(kBlock) peer_gather_f32_kernel(PeerExchange x, float* __restrict__ out) {
  unsigned long long* lw = peer::local_words(x);
  const unsigned long long epoch = lw[0];
  if (threadIdx.x < x.world) peer::wait_flag(peer::reduced_flags(x, x.rank) + threadIdx.x, epoch, lw + 2, 0x300 + threadIdx.x);
  __syncthreads();
  const long long nvec = x.n >> 3;
  const unsigned chunk8 = static_cast<unsigned>(x.chunk >> 3);
  for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
    const int owner = static_cast<int>(static_cast<unsigned>(v) / chunk8);
    const uint4 raw = peer::ld_sys_v4(reinterpret_cast<const __nv_bfloat16*>(x.reduced[owner]) + (v << 3));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]),
                 d = __bfloat1622float2(h[3]);
    reinterpret_cast<float4*>(out)[2 * v] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(out)[2 * v + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
}

// Same shared-memory carveout as the persistent GEMM these kernels run beside: an SM cannot change its L1 / shared
// split while blocks of the other configuration are resident, which would serialise the two kernels.
void set_carveout() {
  static const bool once = [] {
    cudaFuncSetAttribute(peer_signal_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_signal_small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_flag_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_gather_f32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_reduce_slice_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_reduce_slice_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_reduce_slice_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(peer_reduce_slice_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return true;
  }();
  (void)once;
}

int check_exchange(const PeerExchange& x) {
  if (x.world < 2 || x.world > kMaxPeers || x.rank < 0 || x.rank >= x.world)
    return set_error(ED2_ERR_BAD_ARG, "peer exchange: world %d (2..%d) / rank %d", x.world, kMaxPeers, x.rank);
  if (x.n <= 0 || x.n % 8 != 0 || x.chunk <= 0 || x.chunk % 8 != 0 || x.chunk * x.world < x.n)
    return set_error(ED2_ERR_BAD_SHAPE, "peer exchange: n %lld and chunk %lld must be multiples of 8 covering n", x.n, x.chunk);
  for (int r = 0; r < x.world; ++r)
    if (x.bucket[r] == nullptr || x.reduced[r] == nullptr || x.flags[r] == nullptr || (x.small_n > 0 && x.small[r] == nullptr))
      return set_error(ED2_ERR_BAD_ARG, "peer exchange: null pointer for rank %d", r);
  return ED2_OK;
}

}  // namespace

int k_peer_signal(const PeerExchange& x, const float* small_src, cudaStream_t s) {
  int st = check_exchange(x);
  if (st != ED2_OK) return st;
  if (x.small_n > 0 && small_src == nullptr) return set_error(ED2_ERR_BAD_ARG, "peer signal: small_src is null");
  set_carveout();
  // without a companion vector the kernel only bumps the epoch and writes <= 8 flags: one warp, which fits beside a
  // persistent GEMM CTA that owns the rest of the SM (a 256-thread block would wait for that GEMM's tail)
  if (x.small_n > 0) peer_signal_kernel<<<1, kBlock, 0, s>>>(x, small_src);
  else peer_signal_small_kernel<<<1, 32, 0, s>>>(x);
  count_launch();
  return check_launch("peer_signal");
}

int k_peer_flag(const PeerExchange& x, int which, cudaStream_t s) {
  int st = check_exchange(x);
  if (st != ED2_OK) return st;
  if (which != 0 && which != 1) return set_error(ED2_ERR_BAD_ARG, "peer flag: which must be 0 (ready) or 1 (reduced)");
  peer_flag_kernel<<<1, 32, 0, s>>>(x, which);
  count_launch();
  return check_launch("peer_flag");
}

int k_peer_wait(const PeerExchange& x, int which, cudaStream_t s) {
  int st = check_exchange(x);
  if (st != ED2_OK) return st;
  if (which != 0 && which != 1) return set_error(ED2_ERR_BAD_ARG, "peer wait: which must be 0 (ready) or 1 (reduced)");
  static const bool once = [] {
    cudaFuncSetAttribute(peer_wait_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return true;
  }();
  (void)once;
  peer_wait_kernel<<<1, 32, 0, s>>>(x, which);
  count_launch();
  return check_launch("peer_wait");
}

int k_peer_reduce_slice(const PeerExchange& x, int light, cudaStream_t s) {
  int st = check_exchange(x);
  if (st != ED2_OK) return st;
  set_carveout();
  // The kernel runs on a side stream and mostly waits on NVLink.  Default footprint: 2 blocks x 256 threads per SM
  // (beside the persistent GEMM, which leaves > 32 K registers); light: 1 block x 128 threads per SM (4 K registers),
  // which fits beside the gradient-finishing pass at its full occupancy.
  const int threads = light ? 128 : kBlock;
  const long long vec = (x.chunk / 8 + threads - 1) / threads;
  const int g = static_cast<int>(std::max<long long>(1, std::min<long long>(vec, light ? 148 : 148 * 2)));
  switch (x.world) {
    case 2: peer_reduce_slice_kernel<2><<<g, threads, 0, s>>>(x); break;
    case 4: peer_reduce_slice_kernel<4><<<g, threads, 0, s>>>(x); break;
    case 8: peer_reduce_slice_kernel<8><<<g, threads, 0, s>>>(x); break;
    default: peer_reduce_slice_kernel<0><<<g, threads, 0, s>>>(x); break;
  }
  count_launch();
  return check_launch("peer_reduce_slice");
}

int k_peer_gather_f32(const PeerExchange& x, float* out, cudaStream_t s) {
  int st = check_exchange(x);
  if (st != ED2_OK) return st;
  if (out == nullptr || (reinterpret_cast<uintptr_t>(out) & 15) != 0)
    return set_error(ED2_ERR_BAD_ARG, "peer gather: out must be a 16-byte aligned device pointer");
  if (x.n >= (1LL << 34)) return set_error(ED2_ERR_BAD_SHAPE, "peer gather: n too large");
  set_carveout();
  const long long blocks = (x.n / 8 + kBlock - 1) / kBlock;
  const int g = static_cast<int>(std::max<long long>(1, std::min<long long>(blocks, 148 * 3)));
  peer_gather_f32_kernel<<<g, kBlock, 0, s>>>(x, out);
  count_launch();
  return check_launch("peer_gather_f32");
}

}  // namespace ed2

using namespace ed2;

namespace {
PeerExchange to_internal(const ed2_peer_exchange* e) {
  PeerExchange x{};
  x.world = e->world;
  x.rank = e->rank;
  x.n = e->n;
  x.chunk = e->chunk;
  x.small_n = e->small_n;
  for (int r = 0; r < kMaxPeers && r < e->world; ++r) {
    x.bucket[r] = e->bucket[r];
    x.reduced[r] = e->reduced[r];
    x.small[r] = e->small_vec[r];
    x.flags[r] = reinterpret_cast<unsigned long long*>(e->flags[r]);
  }
  return x;
}
}  // namespace

extern "C" {

int ed2_peer_alloc(size_t bytes, void** ptr_out, void* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == ED2_IPC_HANDLE_BYTES, "IPC handle size");
  if (bytes == 0 || ptr_out == nullptr || handle_out == nullptr) return set_error(ED2_ERR_BAD_ARG, "ed2_peer_alloc: bad argument");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "ed2_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_error(ED2_ERR_CUDA, "ed2_peer_alloc: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return ED2_OK;
}

int ed2_peer_open(const void* handle, void** ptr_out) {
  if (handle == nullptr || ptr_out == nullptr) return set_error(ED2_ERR_BAD_ARG, "ed2_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "ed2_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  *ptr_out = p;
  return ED2_OK;
}

int ed2_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? ED2_OK : set_error(ED2_ERR_CUDA, "ed2_peer_close: %s", cudaGetErrorString(e));
}

int ed2_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? ED2_OK : set_error(ED2_ERR_CUDA, "ed2_peer_free: %s", cudaGetErrorString(e));
}

int ed2_peer_flag(const ed2_peer_exchange* e, int which, void* stream) {
  return k_peer_flag(to_internal(e), which, reinterpret_cast<cudaStream_t>(stream));
}

int ed2_peer_wait(const ed2_peer_exchange* e, int which, void* stream) {
  return k_peer_wait(to_internal(e), which, reinterpret_cast<cudaStream_t>(stream));
}

int ed2_dma_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return ED2_OK;
  if (dst == nullptr || src == nullptr) return set_error(ED2_ERR_BAD_ARG, "dma_copy: null pointer");
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "dma_copy: %s", cudaGetErrorString(e));
  return ED2_OK;
}

int ed2_peer_signal(const ed2_peer_exchange* e, const float* small_src, void* stream) {
  return k_peer_signal(to_internal(e), small_src, reinterpret_cast<cudaStream_t>(stream));
}

int ed2_peer_reduce_slice(const ed2_peer_exchange* e, int light, void* stream) {
  return k_peer_reduce_slice(to_internal(e), light, reinterpret_cast<cudaStream_t>(stream));
}

int ed2_peer_gather_f32(const ed2_peer_exchange* e, float* out, void* stream) {
  return k_peer_gather_f32(to_internal(e), out, reinterpret_cast<cudaStream_t>(stream));
}

int ed2_peer_grad_finish(const ed2_peer_exchange* e, const float* mu, const float* rho, ed2_noise eps, float kl_grad_scale,
                         const double* kl_upstream, float prior_mean, float prior_std, float* dmu, float* drho,
                         float* small_out, int fast, int blocks_per_sm, void* stream) {
  return k_peer_grad_finish(to_internal(e), mu, rho, Noise{eps.injected, eps.seed, eps.stream, eps.step}, kl_grad_scale,
                            kl_upstream, Prior{prior_mean, prior_std}, dmu, drho, small_out, fast != 0, blocks_per_sm,
                            reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
