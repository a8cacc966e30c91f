// Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
// SC'11; Random123 reference constants) and the element-indexed noise streams used by every sampling
// kernel in this library.
//
// Stream contract (identical on host and device, see oracle/philox.py):
//   key      = (seed_lo, seed_hi)
//   counter  = (block_lo, block_hi, stream_lo, stream_hi)     block = element_index / 4
//   outputs  = 4 x uint32 -> the 4 consecutive elements 4*block .. 4*block+3 of the stream
// so the value of element i depends only on (seed, stream, i): forward, backward and any tiling regenerate
// the same epsilon without storing it (SURVEY.md §7 hard part 3).
//   uniform  u = ((x >> 8) + 0.5) * 2^-24            in (0, 1) (fp32 rounds the top half of the range to 2^-24 steps)
//   normal   EIGHT per block: element i lives in block i / 8; word w = (i / 2) % 4 of that block gives the pair
//            (z_2w, z_2w+1) = Box-Muller(u0, u1) with the 16-bit uniforms u0 = (hi16(x_w) + 0.5) 2^-16,
//            u1 = (lo16(x_w) + 0.5) 2^-16  (|z| <= 4.86; the radius takes 65536 values, the angle 65536).  The ten
//            Philox rounds are ~3/4 of the instructions of a draw, so halving the blocks per normal is what the
//            issue-bound consumers (GEMM epilogues that regenerate per-example noise, the sampling pass) pay for.
//   sign     ONE BIT per element: element i is bit (i % 32) of word (i / 32) % 4 of block i / 128; +1 if set else -1
//            (P(+1) = 1/2).  A tile epilogue that regenerates the signs of 64 consecutive columns of a row needs one
//            or two Philox blocks instead of sixteen (gemm_flipout.cuh), which is what keeps the short-K Flipout
//            launches tensor-bound.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ed2 {

struct PhiloxKey {
  uint32_t k0, k1;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
  const uint64_t p0 = static_cast<uint64_t>(M0) * c[0], p1 = static_cast<uint64_t>(M1) * c[2];
  const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
  const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
#endif
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += W0;
    k1 += W1;
  }
}

__host__ __device__ __forceinline__ void philox_block(uint64_t seed, uint64_t stream, uint64_t block,
                                                      uint32_t (&out)[4]) {
  out[0] = static_cast<uint32_t>(block);
  out[1] = static_cast<uint32_t>(block >> 32);
  out[2] = static_cast<uint32_t>(stream);
  out[3] = static_cast<uint32_t>(stream >> 32);
  philox4x32_10(out, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

__host__ __device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  return (static_cast<float>(x >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
}

__host__ __device__ __forceinline__ float u16_to_uniform(uint32_t h) {
  return (static_cast<float>(h) + 0.5f) * 1.52587890625e-05f;  // 2^-16, exact in fp32
}

// one 32-bit word -> two normals (16-bit uniforms from its high and low halves)
__device__ __forceinline__ void box_muller(uint32_t x, float& z0, float& z1) {
  const float u0 = u16_to_uniform(x >> 16), u1 = u16_to_uniform(x & 0xFFFFu);
  float r;  // sqrt.approx (1 ulp): the IEEE sqrtf costs a fix-up sequence and a slow-path call per draw
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * logf(u0)));
  float s, c;
  sincospif(2.0f * u1, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// Same pair from hardware approximations (lg2 / sin / cos): abs error ~1e-6 -- used when the sampled tensor is rounded
// to bf16 anyway.  The angle is shifted to (-pi, pi), where __sincosf is accurate: cos(2 pi u) = -cos(2 pi u - pi).
__device__ __forceinline__ void box_muller_fast(uint32_t x, float& z0, float& z1) {
  const float u0 = u16_to_uniform(x >> 16), u1 = u16_to_uniform(x & 0xFFFFu);
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u0)));
  float s, c;
  __sincosf(6.283185307179586f * u1 - 3.141592653589793f, &s, &c);
  z0 = -r * c;
  z1 = -r * s;
}

// the 8 normals of Philox block `block8` (elements 8*block8 .. 8*block8+7)
template <bool kFast = false>
__device__ __forceinline__ void philox_normal8(uint64_t seed, uint64_t stream, uint64_t block8, float (&z)[8]) {
  uint32_t x[4];
  philox_block(seed, stream, block8, x);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if constexpr (kFast) box_muller_fast(x[w], z[2 * w], z[2 * w + 1]);
    else box_muller(x[w], z[2 * w], z[2 * w + 1]);
  }
}

// the first n (<= 8) normals of Philox block `block8`, the rest zero: the Box-Muller transform is skipped for the words
// whose pair lies entirely beyond n (same values as philox_normal8 where defined)
template <bool kFast = false>
__device__ __forceinline__ void philox_normal8_n(uint64_t seed, uint64_t stream, uint64_t block8, int n, float (&z)[8]) {
  uint32_t x[4];
  philox_block(seed, stream, block8, x);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if (2 * w < n) {
      if constexpr (kFast) box_muller_fast(x[w], z[2 * w], z[2 * w + 1]);
      else box_muller(x[w], z[2 * w], z[2 * w + 1]);
    } else {
      z[2 * w] = 0.f;
      z[2 * w + 1] = 0.f;
    }
  }
}

// the 4 normals 4*block4 .. 4*block4+3: one half of Philox block block4 / 2 (callers that consume 8 consecutive
// elements use philox_normal8 and pay for the block once)
template <bool kFast = false>
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t stream, uint64_t block4, float (&z)[4]) {
  uint32_t x[4];
  philox_block(seed, stream, block4 >> 1, x);
  const bool hi = (block4 & 1) != 0;
  const uint32_t w0 = hi ? x[2] : x[0], w1 = hi ? x[3] : x[1];
  if constexpr (kFast) {
    box_muller_fast(w0, z[0], z[1]);
    box_muller_fast(w1, z[2], z[3]);
  } else {
    box_muller(w0, z[0], z[1]);
    box_muller(w1, z[2], z[3]);
  }
}

__device__ __forceinline__ void philox_uniform4(uint64_t seed, uint64_t stream, uint64_t block, float (&u)[4]) {
  uint32_t x[4];
  philox_block(seed, stream, block, x);
#pragma unroll
  // (x >> 8) + 0.5 needs 25 bits above 2^23: the top value rounds to 2^24 in fp32, i.e. u == 1; the uniform STREAM
  // promises the open interval, so it is clamped to 1 - 2^-24.
  for (int i = 0; i < 4; ++i) u[i] = fminf(u32_to_uniform(x[i]), 0.99999994f);
}

// signs of the 4 consecutive elements idx .. idx+3 (idx % 4 == 0: they share one word of one block)
__device__ __forceinline__ void philox_sign4(uint64_t seed, uint64_t stream, uint64_t idx, float (&s)[4]) {
  uint32_t x[4];
  philox_block(seed, stream, idx >> 7, x);
  const uint32_t wsel = static_cast<uint32_t>(idx >> 5) & 3u;
  const uint32_t w = (wsel == 0 ? x[0] : wsel == 1 ? x[1] : wsel == 2 ? x[2] : x[3]) >> (static_cast<uint32_t>(idx) & 31u);
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = ((w >> i) & 1u) ? 1.0f : -1.0f;
}

// Noise source for one tensor: either an injected array (parity tests inject the reference's draws) or a
// Philox stream.  `kind`: 0 normal, 1 sign.
struct Noise {
  const float* injected;  // nullptr -> Philox
  uint64_t seed;
  uint64_t stream;
  const uint64_t* step;   // optional device counter: key = seed + step[0] * 0x9E3779B97F4A7C15 (CUDA-graph replays)
  // Row mapping of a per-example noise tensor [rows][cols] under data parallelism (ed2_noise in include/ed2b200.h):
  // local row m is row noise_row(m) of the GLOBAL batch, so the draw depends on the global row only.
  long long row0;
  int rpg_local;   // rows per group (ensemble member) held locally; 0 = ungrouped
  int rpg_global;  // rows per group in the global batch
};

__host__ __device__ __forceinline__ long long noise_row(const Noise& nz, long long m) {
  if (nz.rpg_local > 0) {
    const long long g = m / nz.rpg_local;
    return m + g * static_cast<long long>(nz.rpg_global - nz.rpg_local) + nz.row0;
  }
  return m + nz.row0;
}
// flat element index of (local row m, column c) of a [rows][cols] noise tensor: injected tensors are indexed locally
__host__ __device__ __forceinline__ long long noise_idx(const Noise& nz, long long m, long long cols, long long c) {
  return (nz.injected != nullptr ? m : noise_row(nz, m)) * cols + c;
}

__device__ __forceinline__ uint64_t noise_key(const Noise& nz) {
  return nz.step != nullptr ? nz.seed + __ldg(reinterpret_cast<const unsigned long long*>(nz.step)) * 0x9E3779B97F4A7C15ull
                            : nz.seed;
}

// 4 consecutive noise values starting at flat element index idx (idx % 4 == 0).
template <int kKind, bool kFast = false>
__device__ __forceinline__ void noise4(const Noise& nz, long long idx, long long n, float (&e)[4]) {
  if (nz.injected != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = (idx + i < n) ? __ldg(nz.injected + idx + i) : 0.f;
  } else if constexpr (kKind == 0) {
    philox_normal4<kFast>(noise_key(nz), nz.stream, static_cast<uint64_t>(idx >> 2), e);
  } else {
    philox_sign4(noise_key(nz), nz.stream, static_cast<uint64_t>(idx), e);
  }
}

// single element at arbitrary flat index
template <int kKind>
__device__ __forceinline__ float noise1(const Noise& nz, long long idx) {
  if (nz.injected != nullptr) return __ldg(nz.injected + idx);
  float e[4];
  if constexpr (kKind == 0) philox_normal4(noise_key(nz), nz.stream, static_cast<uint64_t>(idx >> 2), e);
  else philox_sign4(noise_key(nz), nz.stream, static_cast<uint64_t>(idx) & ~3ull, e);
  return e[idx & 3];
}

}  // namespace ed2
