// Peer-memory (NVLink / NVSwitch, CUDA IPC) gradient exchange of the data-parallel step: shared device helpers.
//
// Every rank owns one IPC-mapped arena per exchanged weight gradient:
//   bucket   [n] bf16   this rank's raw dW (written by the dW GEMM; peers PULL their slice from it)
//   reduced  [n] bf16   only the owned slice [rank*chunk, (rank+1)*chunk) is valid: sum over ranks, in rank order
//   small    [m] fp32   this rank's small companion gradient (the layer's bias gradient), summed one-shot by every rank
//   flags    u64 words  [0, W) "rank r's bucket of epoch e is ready"   (written remotely by rank r)
//                       [W, 2W) "owner o's reduced slice of epoch e is ready" (written remotely by owner o)
//                       [2W] epoch (local)  [2W+1] block counter (local)  [2W+2] error word (local)
// Flags carry a monotonically increasing epoch, so nothing is ever reset and a captured CUDA graph can be replayed.
// A rank polls only its OWN memory and writes flags into its peers' memory (release at system scope).
#pragma once
#include <cstdint>

#include "kernels.h"

namespace ed2 {
namespace peer {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// 16 / 8 bytes of (possibly remote) data that a peer re-writes every step: system-scope relaxed loads (never served
// from this SM's L1)
__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_sys_v2(const void* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr unsigned long long kWaitTimeoutNs = 20ull * 1000 * 1000 * 1000;  // a peer that is 20 s late is dead

// Spin until *flag >= epoch.  A missing peer must not hang the GPU: after the timeout the error word is set and the
// kernel traps (the process then fails loudly with a CUDA error).
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch,
                                          unsigned long long* err, unsigned long long code) {
  if (ld_acquire_sys(flag) >= epoch) return;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) < epoch) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > kWaitTimeoutNs) {
      *err = code;
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ unsigned long long* ready_flags(const PeerExchange& x, int on_rank) { return x.flags[on_rank]; }
__device__ __forceinline__ unsigned long long* reduced_flags(const PeerExchange& x, int on_rank) {
  return x.flags[on_rank] + x.world;
}
__device__ __forceinline__ unsigned long long* local_words(const PeerExchange& x) { return x.flags[x.rank] + 2 * x.world; }

}  // namespace peer
}  // namespace ed2
