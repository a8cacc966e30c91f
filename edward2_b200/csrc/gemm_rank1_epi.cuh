// Fused DenseRank1 epilogues of the tcgen05 GEMM (edward2/tensorflow/layers/dense.py:1096-1144): the per-example fast
// weights are drawn from Philox INSIDE the epilogue, once per output element, so no noise / factor tensor ever exists
// in HBM (gemm_epilogue.h: kEpiRank1Fwd / kEpiRank1Bwd describe the arithmetic).
//
// Why the epilogue and not the operand path: the A tile of a GEMM is re-loaded once per N tile (32 times at
// N = 8192), the accumulator tile is visited exactly once.  Why EIGHT epilogue warps and 16-column chunks: the draw is
// ~25 dependent-latency-bound instructions per normal, and with one warp per scheduler the measured issue rate was
// 0.19 IPC (profiles/r02_rank1_fused_4warp.txt) — the epilogue, not the tensor pipe, set the tile time.  Two warps per
// scheduler, each owning a 128-column half of the tile, with a working set small enough (16 accumulator columns) to
// stay in registers at 168 regs/thread, hide that latency.
//
// Per-(member, column) parameters (mu, sigma, sigmoid(rho)) of the tile's 256 columns are staged once per tile in
// shared memory (all 128 rows of a CTA's tile belong to one member: group_rows % 128 == 0 is required by the host).
// Column reductions of the backward (dmu_r, drho_r, dmu_s, drho_s, dbias) are register butterflies: 16 shuffles per
// 16-column chunk per sum, then one atomic per column per warp.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

#include "gemm_epilogue.h"
#include "philox.cuh"
#include "ptx.cuh"

namespace ed2 {
namespace tc {

__device__ __forceinline__ uint64_t fast_key(const FastSide& f) {
  return f.step != nullptr ? f.seed + __ldg(f.step) * 0x9E3779B97F4A7C15ull : f.seed;
}
__device__ __forceinline__ long long fast_row(const FastSide& f, long long m) {
  if (f.rpg_local > 0) {
    const long long g = m / f.rpg_local;
    return m + g * static_cast<long long>(f.rpg_global - f.rpg_local) + f.row0;
  }
  return m + f.row0;
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// eps of 16 consecutive columns of one row for up to two streams, as ONE straight-line section: the 2 (or 4)
// independent 10-round Philox chains interleave (8 normals per block, philox.cuh).  idx = global_row * N + column with
// idx % 8 == 0 (N % 8 == 0, column % 16 == 0).
template <bool kTwo>
__device__ __forceinline__ void noise16(uint64_t key_a, uint64_t stream_a, long long idx_a, float (&ea)[16],
                                        uint64_t key_b, uint64_t stream_b, long long idx_b, float (&eb)[16]) {
  constexpr int kChains = kTwo ? 4 : 2;
  uint32_t x[kChains][4], k0[kTwo ? 2 : 1], k1[kTwo ? 2 : 1];
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const uint64_t blk = static_cast<uint64_t>(idx_a >> 3) + b;
    x[b][0] = static_cast<uint32_t>(blk); x[b][1] = static_cast<uint32_t>(blk >> 32);
    x[b][2] = static_cast<uint32_t>(stream_a); x[b][3] = static_cast<uint32_t>(stream_a >> 32);
    if constexpr (kTwo) {
      const uint64_t blk2 = static_cast<uint64_t>(idx_b >> 3) + b;
      x[2 + b][0] = static_cast<uint32_t>(blk2); x[2 + b][1] = static_cast<uint32_t>(blk2 >> 32);
      x[2 + b][2] = static_cast<uint32_t>(stream_b); x[2 + b][3] = static_cast<uint32_t>(stream_b >> 32);
    }
  }
  k0[0] = static_cast<uint32_t>(key_a); k1[0] = static_cast<uint32_t>(key_a >> 32);
  if constexpr (kTwo) { k0[1] = static_cast<uint32_t>(key_b); k1[1] = static_cast<uint32_t>(key_b >> 32); }
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      philox_round(x[b], k0[0], k1[0]);
      if constexpr (kTwo) philox_round(x[2 + b], k0[1], k1[1]);
    }
    k0[0] += 0x9E3779B9u; k1[0] += 0xBB67AE85u;
    if constexpr (kTwo) { k0[1] += 0x9E3779B9u; k1[1] += 0xBB67AE85u; }
  }
#pragma unroll
  for (int b = 0; b < 2; ++b) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      box_muller_fast(x[b][w], ea[8 * b + 2 * w], ea[8 * b + 2 * w + 1]);
      if constexpr (kTwo) box_muller_fast(x[2 + b][w], eb[8 * b + 2 * w], eb[8 * b + 2 * w + 1]);
    }
  }
}

__device__ __forceinline__ float fast_apply(float mu, float sg, float eps, float lo, float hi, float& pass) {
  const float raw = fmaf(sg, eps, mu);
  pass = (raw < lo || raw > hi) ? 0.f : 1.f;
  return fminf(fmaxf(raw, lo), hi);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 v = __bfloat1622float2(h[t]);
    f[2 * t] = v.x;
    f[2 * t + 1] = v.y;
  }
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// 16 floats of a staged table starting at column c (c % 16 == 0): four broadcast 16-byte loads
__device__ __forceinline__ void table16(uint32_t base, int c, float (&o)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 q = ld_shared_v4f(base + (c + 4 * j) * 4);
    o[4 * j] = q.x; o[4 * j + 1] = q.y; o[4 * j + 2] = q.z; o[4 * j + 3] = q.w;
  }
}

// Column sums of a 32-row x 16-column register tile (one row per lane): butterfly transpose-reduce.  After it, lane L
// holds in v[0] the sum over the warp's 32 rows of column (L >> 1) & ... — concretely column index col16(L) below.
__device__ __forceinline__ void col_sum16(float (&v)[16], uint32_t lane) {
#pragma unroll
  for (int off = 16; off >= 2; off >>= 1) {
    const int half = off >> 1;  // columns kept per lane after this stage
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// column (0..15) whose total lane L holds after col_sum16: bit (off) of the lane selects the upper half at each stage
__device__ __forceinline__ int col16_of_lane(uint32_t lane) {
  return ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------------------------------
// The epilogue loop of the 8 epilogue warps (warp_idx 4..11).  Group g = (warp_idx - 4) / 4 owns columns
// [g*128, g*128+128) of every tile; quadrant ew = warp_idx % 4 is the TMEM lane quadrant (rows ew*32 .. +31).
template <int kEpi, int BLOCK_N, int kCtaGroup>
__device__ __forceinline__ void rank1_epilogue(const CUtensorMap& tm_out, const EpiParams& ep, const Rank1Epi& r1, int M,
                                               int N, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                               uint64_t* tmem_empty_bar, uint8_t* smem_store, uint8_t* smem_table,
                                               int warp_idx, uint32_t lane, uint32_t cta_rank, int cluster_id,
                                               int num_clusters, int num_m, int num_n) {
  static_assert(BLOCK_N == 256 || BLOCK_N == 128, "fused rank-1 epilogue: 256- or 128-column tiles");
  constexpr int kHalf = BLOCK_N / 2;                 // columns owned by one epilogue group
  constexpr int kChunks = kHalf / kR1Chunk;          // 16-column chunks per group per tile
  constexpr int kTileM = kBlockM * kCtaGroup;
  const bool is_leader = cta_rank == 0;
  const int grp_id = (warp_idx - 4) >> 2;
  const int ew = warp_idx & 3;
  const int row = ew * 32 + lane;
  const int et = (warp_idx - 4) * 32 + lane;  // 0..255
  const bool store_leader = ew == 0 && lane == 0;
  const uint32_t bar_id = 1 + grp_id;
  uint32_t store_buf = 0;
  const int num_tiles = num_m * num_n;
  const bool has_s = r1.s.mu != nullptr, has_r = r1.r.mu != nullptr;
  const bool samp_s = has_s && r1.s.sg != nullptr, samp_r = has_r && r1.r.sg != nullptr;
  const uint64_t key_s = has_s ? fast_key(r1.s) : 0, key_r = has_r ? fast_key(r1.r) : 0;
  const float lo = r1.clip_lo, hi = r1.clip_hi;
  const bool relu_prev = r1.prev_act == kActRelu;
  int iter = 0;
  for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
    const TileCoord tcd = tile_coord(t, num_m, num_n);
    const int m0 = tcd.m_blk * kTileM + cta_rank * kBlockM;
    const int n0 = tcd.n_blk * BLOCK_N;
    const int m = m0 + row;
    const bool m_ok = m < M;
    const int e = ep.group_rows > 0 ? min(m0, M - 1) / ep.group_rows : 0;  // CTA-uniform (group_rows % 128 == 0)
    // ---- stage the tile's per-column parameters (double-buffered by tile parity; one barrier per tile)
    const uint32_t tab = ptx::smem_u32(smem_table) + (iter & 1) * (kR1TableFloats * 4);
    {
      const int n = n0 + et;
      const long long o = static_cast<long long>(e) * N + n;
      const bool ok = n < N && et < BLOCK_N;
      st_shared_f32(tab + (0 * 256 + et) * 4, (ok && has_s) ? __ldg(r1.s.mu + o) : 1.f);
      st_shared_f32(tab + (1 * 256 + et) * 4, (ok && samp_s) ? __ldg(r1.s.sg + o) : 0.f);
      st_shared_f32(tab + (3 * 256 + et) * 4, (ok && has_r) ? __ldg(r1.r.mu + o) : 1.f);
      st_shared_f32(tab + (4 * 256 + et) * 4, (ok && samp_r) ? __ldg(r1.r.sg + o) : 0.f);
      if constexpr (kEpi == kEpiRank1Fwd) {
        st_shared_f32(tab + (2 * 256 + et) * 4, (ok && ep.col_bias != nullptr) ? __ldg(ep.col_bias + static_cast<long long>(e) * ep.col_bias_ld + n) : 0.f);
      } else {
        st_shared_f32(tab + (2 * 256 + et) * 4, (ok && samp_s) ? __ldg(r1.s.sgm + o) : 0.f);
        st_shared_f32(tab + (5 * 256 + et) * 4, (ok && samp_r) ? __ldg(r1.r.sgm + o) : 0.f);
      }
    }
    ptx::named_bar_sync(3, kR1EpiThreads);

    const uint32_t acc = iter & 1;
    const uint32_t acc_phase = (iter >> 1) & 1;
    ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
    ptx::tc_fence_after();
    const long long nrow_s = has_s ? fast_row(r1.s, m_ok ? m : 0) * N : 0;
    const long long nrow_r = has_r ? fast_row(r1.r, m_ok ? m : 0) * N : 0;

    // backward: this layer's input x (= the producer's y) and the producer's z, one chunk ahead
    [[maybe_unused]] uint4 xq[2], zq[2];
    [[maybe_unused]] const __nv_bfloat16* x_row = nullptr;
    [[maybe_unused]] const __nv_bfloat16* z_row = nullptr;
    const int ncol0 = n0 + grp_id * kHalf;
    if constexpr (kEpi == kEpiRank1Bwd) {
      x_row = reinterpret_cast<const __nv_bfloat16*>(r1.x) + static_cast<long long>(m_ok ? m : 0) * r1.x_ld;
      if (has_s) z_row = reinterpret_cast<const __nv_bfloat16*>(r1.z_prev) + static_cast<long long>(m_ok ? m : 0) * r1.z_prev_ld;
      if (m_ok) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (ncol0 + 8 * j < N) {  // N % 8 == 0: an 8-column vector is entirely inside or outside
            xq[j] = __ldg(reinterpret_cast<const uint4*>(x_row + ncol0) + j);
            if (has_s) zq[j] = __ldg(reinterpret_cast<const uint4*>(z_row + ncol0) + j);
          }
        }
      }
    }

#pragma unroll 1
    for (int cc = 0; cc < kChunks; ++cc) {
      const int tcol = grp_id * kHalf + cc * kR1Chunk;  // column inside the tile
      const int nc = n0 + tcol;
      uint32_t vr[16];
      tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N + tcol, vr);
      ptx::tmem_ld_wait();
      if (cc == kChunks - 1) {
        ptx::tc_fence_before();
        if (is_leader) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        else ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
      }
      const int n_valid = max(0, min(kR1Chunk, N - nc));  // 0, 8 or 16 (N % 8 == 0)
      const bool c_ok = n_valid > 0, full = n_valid == kR1Chunk;
      const bool active = m_ok && c_ok;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(vr[j]) * ep.alpha;

      if constexpr (kEpi == kEpiRank1Fwd) {
        // z -> raw;  y = act(z * s + bias[e]) -> out;  out2 = bf16(y) * r_next
        if (ep.raw != nullptr && active) {
          uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.raw) + static_cast<long long>(m) * ep.raw_ld + nc);
          p[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          if (full) p[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        }
        const bool want_r = r1.out2 != nullptr;
        float es[16], er[16];
        if (samp_s && samp_r && want_r) {
          noise16<true>(key_s, r1.s.stream, nrow_s + nc, es, key_r, r1.r.stream, nrow_r + nc, er);
        } else {
          if (samp_s) noise16<false>(key_s, r1.s.stream, nrow_s + nc, es, 0, 0, 0, er);
          if (samp_r && want_r) noise16<false>(key_r, r1.r.stream, nrow_r + nc, er, 0, 0, 0, es);
        }
        {
          float mu[16], sg[16], bs[16];
          table16(tab + 0 * 1024, tcol, mu);
          table16(tab + 1 * 1024, tcol, sg);
          table16(tab + 2 * 1024, tcol, bs);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float ps;
            const float f = samp_s ? fast_apply(mu[j], sg[j], es[j], lo, hi, ps) : mu[j];
            float y = fmaf(v[j], f, bs[j]);
            if (ep.act == kActRelu) y = fmaxf(y, 0.f);
            v[j] = y;
          }
        }
        if (want_r) {
          float mu[16], sg[16];
          table16(tab + 3 * 1024, tcol, mu);
          table16(tab + 4 * 1024, tcol, sg);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float p0, p1;
            const float f0 = samp_r ? fast_apply(mu[2 * j], sg[2 * j], er[2 * j], lo, hi, p0) : mu[2 * j];
            const float f1 = samp_r ? fast_apply(mu[2 * j + 1], sg[2 * j + 1], er[2 * j + 1], lo, hi, p1) : mu[2 * j + 1];
            pk[j] = pack_bf16x2(bf16_round(v[2 * j]) * f0, bf16_round(v[2 * j + 1]) * f1);
          }
          if (active) {
            uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(r1.out2) + static_cast<long long>(m) * r1.out2_ld + nc);
            p[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (full) p[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      } else {
        // dxr = acc;  dx = dxr * r;  dr sums;  producer: gp = bf16(dx) * act'(x), dz = gp * s, ds / dbias sums
        float er[16], es[16];
        if (samp_r && samp_s) noise16<true>(key_r, r1.r.stream, nrow_r + nc, er, key_s, r1.s.stream, nrow_s + nc, es);
        else if (samp_r) noise16<false>(key_r, r1.r.stream, nrow_r + nc, er, 0, 0, 0, es);
        else if (samp_s) noise16<false>(key_s, r1.s.stream, nrow_s + nc, es, 0, 0, 0, er);
        float xs[16], zs[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { xs[j] = 0.f; zs[j] = 0.f; }
        if (active) {
          unpack_bf16x8(xq[0], xs);
          if (has_s) unpack_bf16x8(zq[0], zs);
          if (full) {
            unpack_bf16x8(xq[1], xs + 8);
            if (has_s) unpack_bf16x8(zq[1], zs + 8);
          }
        }
        // next chunk's operands: in flight during this chunk's arithmetic
        if (m_ok && cc + 1 < kChunks) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (nc + kR1Chunk + 8 * j < N) {
              xq[j] = __ldg(reinterpret_cast<const uint4*>(x_row + nc + kR1Chunk) + j);
              if (has_s) zq[j] = __ldg(reinterpret_cast<const uint4*>(z_row + nc + kR1Chunk) + j);
            }
          }
        }
        const int warp_row0 = m0 + ew * 32;
        const int my_col = nc + col16_of_lane(lane);
        const bool warp_ok = warp_row0 < M && my_col < N;
        const long long o = static_cast<long long>(e) * N + my_col;
        uint32_t dxp[8];
        {
          float mu[16], sg[16], q0[16], q1[16];
          table16(tab + 3 * 1024, tcol, mu);
          table16(tab + 4 * 1024, tcol, sg);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p = 1.f;
            const float f = samp_r ? fast_apply(mu[j], sg[j], er[j], lo, hi, p) : mu[j];
            const float dxr = (active && j < n_valid) ? v[j] : 0.f;
            const float dr = dxr * xs[j] * p;
            q0[j] = dr;
            q1[j] = samp_r ? dr * er[j] : 0.f;
            v[j] = dxr * f;  // dx
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) dxp[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          col_sum16(q0, lane);
          if (samp_r) col_sum16(q1, lane);
          if (warp_ok && (lane & 1) == 0) {
            atomicAdd(r1.dmu_r + o, q0[0]);
            if (samp_r && r1.drho_r != nullptr) atomicAdd(r1.drho_r + o, q1[0] * ld_shared_f32(tab + (5 * 256 + tcol + col16_of_lane(lane)) * 4));
          }
        }
        if (has_s) {
          if (r1.out2 != nullptr && active) {
            uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(r1.out2) + static_cast<long long>(m) * r1.out2_ld + nc);
            p[0] = make_uint4(dxp[0], dxp[1], dxp[2], dxp[3]);
            if (full) p[1] = make_uint4(dxp[4], dxp[5], dxp[6], dxp[7]);
          }
          float mu[16], sg[16], q2[16], q3[16], q4[16];
          table16(tab + 0 * 1024, tcol, mu);
          table16(tab + 1 * 1024, tcol, sg);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p = 1.f;
            const float f = samp_s ? fast_apply(mu[j], sg[j], es[j], lo, hi, p) : mu[j];
            const float dxb = bf16_round(v[j]);
            const float gp = relu_prev ? (xs[j] > 0.f ? dxb : 0.f) : dxb;
            const float ds = gp * zs[j] * p;
            q2[j] = ds;
            q3[j] = samp_s ? ds * es[j] : 0.f;
            q4[j] = gp;
            v[j] = gp * f;  // dz of the producer -> out
          }
          col_sum16(q2, lane);
          if (samp_s) col_sum16(q3, lane);
          if (r1.dbias != nullptr) col_sum16(q4, lane);
          if (warp_ok && (lane & 1) == 0) {
            atomicAdd(r1.dmu_s + o, q2[0]);
            if (samp_s && r1.drho_s != nullptr) atomicAdd(r1.drho_s + o, q3[0] * ld_shared_f32(tab + (2 * 256 + tcol + col16_of_lane(lane)) * 4));
            if (r1.dbias != nullptr) atomicAdd(r1.dbias + o, q4[0]);
          }
        }
      }

      // ---- out (bf16): 64-column staging buffers, one TMA store per buffer; 4 chunks per buffer
      if (ep.out == nullptr) continue;
      if (!ep.out_tma) {
        if (active) {
          uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + static_cast<long long>(m) * ep.out_ld + nc);
          p[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          if (full) p[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        }
        continue;
      }
      const int sub = cc & 3;
      uint8_t* buf = smem_store + (grp_id * 2 + store_buf) * kStoreBufBytes;
      if (sub == 0) {
        if (store_leader) ptx::tma_store_wait_read<1>();
        ptx::named_bar_sync(bar_id, 128);
      }
      const uint32_t buf_addr = ptx::smem_u32(buf) + row * 128;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int chunk = (sub * 2 + j) ^ (row & 7);
        ptx::st_shared_v4(buf_addr + chunk * 16, pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      }
      if (sub == 3) {
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(bar_id, 128);
        if (store_leader) {
          const int col0 = nc - 48;
          if (col0 < N && m0 < M) ptx::tma_store_2d(&tm_out, buf, col0, m0);
          ptx::tma_store_commit();
        }
        store_buf ^= 1;
      }
    }
  }
  if (ep.out_tma && store_leader) ptx::tma_store_wait_all<0>();
}

}  // namespace tc
}  // namespace ed2
