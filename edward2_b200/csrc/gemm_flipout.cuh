// DenseFlipout as ONE tcgen05 kernel with two TMEM accumulators (edward2/tensorflow/layers/dense.py:255-285).
//
//   out[m,n]  = post( A0·B0 + (A1·B1) ∘ sa[m,n] )            sa, sb: per-example ±1 Philox streams, regenerated in the
//   out2[m,n] = bf16(out[m,n]) ∘ sb[m,n]      (optional)       epilogue from (seed, stream, global row, column)
//
// forward   A0 = x, A1 = x∘S, B0 = mean, B1 = Delta:  y = act(x·mean + ((x∘S)·Delta)∘T + b);  out2 = y∘S_next is the NEXT
//           Flipout layer's second operand, so only the first layer of a stack needs a sign pass over its input.
// backward  A0 = gp, A1 = gp∘T, B0 = meanᵀ, B1 = Deltaᵀ:  dx = gp·meanᵀ + ((gp∘T)·Deltaᵀ)∘S;  with a producer layer
//           (its ReLU output is this layer's x): out = gp_prev = dx∘1[x > 0], out2 = gp_prev∘T_prev, col_sum = the
//           producer's bias gradient — the producer's backward then starts directly with its weight-gradient GEMMs.
// Both products share the tile schedule and the k-block pipeline: one TMA stage carries the k-block of all four
// operands, the MMA warp issues the two accumulations back to back, and the epilogue reads both accumulators.  Neither
// the fp32 perturbation [B][O] nor the sign tensors ever exist in HBM (the two-GEMM formulation wrote and re-read
// 134 MB per 4096x4096 layer and materialised T and S with separate fill passes).
//
// Tile: 2-CTA pair, 256 rows x 128 columns per accumulator; TMEM = 2 stages x (128 + 128) columns = 512.
// Stage = A0 16 KB + A1 16 KB + B0 8 KB + B1 8 KB (per CTA), 4 stages.  8 epilogue warps: group g (warps 4-7 / 8-11)
// owns columns [64g, 64g+64) of both accumulators = one 64-column bf16 staging buffer and ONE TMA store per tile.
#pragma once
#include "gemm_sm100.cuh"

namespace ed2 {
namespace tc {

constexpr int kFoBlockN = 128;
constexpr int kFoStages = 4;
constexpr int kFoABytes = kBlockM * kBlockK * 2;         // 16 KB
constexpr int kFoBBytes = (kFoBlockN / 2) * kBlockK * 2;  // 8 KB: each CTA stages half of the B tile
constexpr int kFoStageBytes = 2 * kFoABytes + 2 * kFoBBytes;
constexpr int kFoThreads = 128 + 256;
constexpr int kFoEpiThreads = 256;
constexpr int kFoSmemBytes = kFoStages * kFoStageBytes + 2 * kStoreBufBytes + 1024 + 1024;
constexpr int kFoTmemCols = 512;

__device__ __forceinline__ uint64_t sign_key(const SignSide& s) {
  return s.step != nullptr ? s.seed + __ldg(s.step) * 0x9E3779B97F4A7C15ull : s.seed;
}

// Grouped rasterisation with a run-time band height: consecutive tiles walk down `group` M-blocks before moving to
// the next N-block.  The DRAM traffic of one wave of C concurrent tiles is (unique M-blocks x 256 + unique N-blocks x
// 128) rows of K elements per product, minimal for a band ~ sqrt(C / 2) high (measured with the fixed 16-block band of
// the generic kernel: 503 MB read for 134 MB of operands at 4096^3, L2 hit rate 69 %); the host picks the band.
__device__ __forceinline__ TileCoord tile_coord_g(int t, int num_m, int num_n, int group) {
  const int per_group = group * num_n;
  const int g = t / per_group;
  const int first_m = g * group;
  const int gsize = min(group, num_m - first_m);
  const int r = t - g * per_group;
  TileCoord c;
  c.m_blk = first_m + r % gsize;
  c.n_blk = r / gsize;
  return c;
}

// 64 consecutive sign bits of one row, starting at stream element idx (idx % 8 == 0), for one or two streams: bit j of
// the result is 1 when element idx + j is +1 (philox.cuh: one bit per element, 128 elements per Philox block).  The two
// blocks that can hold them are evaluated together (2 or 4 independent 10-round chains interleave).
__device__ __forceinline__ uint64_t extract64(const uint32_t (&w)[8], uint32_t p) {
  const uint32_t sh = p & 31u;
  uint32_t a, b, c;
  switch (p >> 5) {
    case 0: a = w[0]; b = w[1]; c = w[2]; break;
    case 1: a = w[1]; b = w[2]; c = w[3]; break;
    case 2: a = w[2]; b = w[3]; c = w[4]; break;
    default: a = w[3]; b = w[4]; c = w[5]; break;
  }
  const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
template <bool kTwo>
__device__ __forceinline__ void sign_bits64(uint64_t key_a, uint64_t stream_a, long long idx_a, uint64_t& bits_a,
                                            uint64_t key_b, uint64_t stream_b, long long idx_b, uint64_t& bits_b) {
  constexpr int kChains = kTwo ? 4 : 2;
  uint32_t x[kChains][4], k0[kTwo ? 2 : 1], k1[kTwo ? 2 : 1];
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    const uint64_t blk = (static_cast<uint64_t>(idx_a) >> 7) + b;
    x[b][0] = static_cast<uint32_t>(blk); x[b][1] = static_cast<uint32_t>(blk >> 32);
    x[b][2] = static_cast<uint32_t>(stream_a); x[b][3] = static_cast<uint32_t>(stream_a >> 32);
    if constexpr (kTwo) {
      const uint64_t blk2 = (static_cast<uint64_t>(idx_b) >> 7) + b;
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
  {
    const uint32_t w[8] = {x[0][0], x[0][1], x[0][2], x[0][3], x[1][0], x[1][1], x[1][2], x[1][3]};
    bits_a = extract64(w, static_cast<uint32_t>(idx_a) & 127u);
  }
  if constexpr (kTwo) {
    const uint32_t w[8] = {x[2][0], x[2][1], x[2][2], x[2][3], x[3][0], x[3][1], x[3][2], x[3][3]};
    bits_b = extract64(w, static_cast<uint32_t>(idx_b) & 127u);
  }
}

// bf16-path softplus (elementwise.cu: sigma_from_rho_fast — the finishing pass must see the forward's sigma)
__device__ __forceinline__ float fo_sigma_from_rho_fast(float rho, float& e) {
  e = __expf(fminf(rho, 80.f));
  const float sp = e < 2.44140625e-4f ? e : __logf(1.f + e);
  return (rho > 20.f ? rho : sp) + 1e-7f;
}

// kMode 0: the summed epilogue above.
// kMode 1: the two WEIGHT-GRADIENT products of a Flipout layer in one launch, finished in the epilogue:
//   acc0 = xᵀ·gp, acc1 = (x∘S)ᵀ·(gp∘T)   (A MN-major: the activations are read transposed by the TMA descriptors)
//   dmu[i,o]  = acc0 + klg (mu - m) / s²
//   drho[i,o] = (acc1 · eps[i,o] + klg (sigma / s² - 1 / sigma)) · sigmoid(rho)     eps regenerated from Philox
//   — the arithmetic of grad_finish_lean_kernel, so the fp32 raw gradients are never written and re-read (24 B per
//   weight) and the two products share one tile grid (512 tiles = 6.9 waves of 74 pairs instead of 2 x 3.46).
// kMode 2: the same finishing epilogue on ONE product (DenseReparameterization: the single weight gradient xᵀ·gp feeds
//   both dmu and drho); only A0 / B0 are staged, so the ring is 8 stages of 24 KB.
template <int kMajorA, int kMajorB, int kMode>
__global__ void __maxnreg__(168)
gemm_flipout_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                    const __grid_constant__ CUtensorMap tm_b0, const __grid_constant__ CUtensorMap tm_b1,
                    const __grid_constant__ CUtensorMap tm_out, int M, int N, int K, void* out, int out_ld, int out_tma,
                    int raster_group, const __grid_constant__ FlipoutEpi fe) {
  constexpr int kTileM = kBlockM * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr bool kDual = kMode != 2;
  constexpr int kSt = kDual ? kFoStages : 2 * kFoStages;  // same bytes: 4 x 48 KB or 8 x 24 KB
  uint8_t* smem_a0 = smem;
  uint8_t* smem_a1 = smem_a0 + kSt * kFoABytes;                       // dual only
  uint8_t* smem_b0 = smem_a1 + (kDual ? kSt * kFoABytes : 0);
  uint8_t* smem_b1 = smem_b0 + kSt * kFoBBytes;                       // dual only
  uint8_t* smem_store = smem + kFoStages * kFoStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + 2 * kStoreBufBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kSt;
  uint64_t* tmem_full_bar = bars + 2 * kSt;
  uint64_t* tmem_empty_bar = bars + 2 * kSt + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kSt + 4);

  const int warp_idx = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
  const uint32_t lane = ptx::lane_id();
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool is_leader = cta_rank == 0;

  ptx::cluster_sync();
  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm_a0);
    ptx::prefetch_tensormap(&tm_a1);
    ptx::prefetch_tensormap(&tm_b0);
    ptx::prefetch_tensormap(&tm_b1);
    if (out_tma) ptx::prefetch_tensormap(&tm_out);
  }
  if (warp_idx == 1 && ptx::elect_one()) {
    for (int i = 0; i < kSt; ++i) {
      ptx::mbar_init(&full_bar[i], 2);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], 2 * kFoEpiThreads);
    }
    ptx::fence_barrier_init();
  } else if (warp_idx == 2) {
    ptx::tmem_alloc<2>(tmem_ptr_smem, kFoTmemCols);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_m = (M + kTileM - 1) / kTileM;
  const int num_n = (N + kFoBlockN - 1) / kFoBlockN;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int num_tiles = num_m * num_n;
  const int cluster_id = blockIdx.x / 2;
  const int num_clusters = gridDim.x / 2;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const TileCoord tc = tile_coord_g(t, num_m, num_n, raster_group);
        const int m0 = tc.m_blk * kTileM + cta_rank * kBlockM;
        const int n0 = tc.n_blk * kFoBlockN + cta_rank * (kFoBlockN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const int k0 = kb * kBlockK;
          uint64_t* fb = &full_bar[stage];
          if constexpr (kMajorA == kMajorK) {
            ptx::tma_load_2d_pair(smem_a0 + stage * kFoABytes, &tm_a0, fb, k0, m0);
            if constexpr (kDual) ptx::tma_load_2d_pair(smem_a1 + stage * kFoABytes, &tm_a1, fb, k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i) {
              ptx::tma_load_2d_pair(smem_a0 + stage * kFoABytes + i * (kBlockK * 128), &tm_a0, fb, m0 + i * 64, k0);
              if constexpr (kDual)
                ptx::tma_load_2d_pair(smem_a1 + stage * kFoABytes + i * (kBlockK * 128), &tm_a1, fb, m0 + i * 64, k0);
            }
          }
          if constexpr (kMajorB == kMajorK) {
            ptx::tma_load_2d_pair(smem_b0 + stage * kFoBBytes, &tm_b0, fb, k0, n0);
            if constexpr (kDual) ptx::tma_load_2d_pair(smem_b1 + stage * kFoBBytes, &tm_b1, fb, k0, n0);
          } else {
            ptx::tma_load_2d_pair(smem_b0 + stage * kFoBBytes, &tm_b0, fb, n0, k0);
            if constexpr (kDual) ptx::tma_load_2d_pair(smem_b1 + stage * kFoBBytes, &tm_b1, fb, n0, k0);
          }
          if (is_leader) ptx::mbar_arrive_expect_tx(fb, (kDual ? kFoStageBytes : kFoStageBytes / 2) * 2);
          else ptx::mbar_arrive_cluster(fb, 0);
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (is_leader && ptx::elect_one()) {
      constexpr uint32_t idesc = make_idesc<kMajorA, kMajorB>(kTileM, kFoBlockN);
      uint32_t stage = 0, phase = 0;
      int iter = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
        const uint32_t acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d0 = tmem_base + acc * (2 * kFoBlockN);
        const uint32_t tmem_d1 = tmem_d0 + kFoBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da0 = operand_desc<kMajorA>(ptx::smem_u32(smem_a0 + stage * kFoABytes));
          const uint64_t da1 = operand_desc<kMajorA>(ptx::smem_u32(smem_a1 + stage * kFoABytes));
          const uint64_t db0 = operand_desc<kMajorB>(ptx::smem_u32(smem_b0 + stage * kFoBBytes));
          const uint64_t db1 = operand_desc<kMajorB>(ptx::smem_u32(smem_b1 + stage * kFoBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
            ptx::umma<2, false>(tmem_d0, da0 + k * desc_k_step<kMajorA>(), db0 + k * desc_k_step<kMajorB>(), idesc, accum);
            if constexpr (kDual)
              ptx::umma<2, false>(tmem_d1, da1 + k * desc_k_step<kMajorA>(), db1 + k * desc_k_step<kMajorB>(), idesc, accum);
          }
          ptx::umma_commit<2>(&empty_bar[stage]);
          if (kb == num_kb - 1) ptx::umma_commit<2>(&tmem_full_bar[acc]);
          if (++stage == kSt) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx >= 4) {
   if constexpr (kMode >= 1) {
    // ------------------------------------------------------------------ weight-gradient epilogue (8 warps)
    const int grp_id = (warp_idx - 4) >> 2;
    const int ew = warp_idx & 3;
    const int row = ew * 32 + lane;
    const int gtid = ew * 32 + static_cast<int>(lane);  // thread index within the group (== row)
    const uint32_t bar_id = 1 + grp_id;
    const uint32_t sbuf = ptx::smem_u32(smem_store + grp_id * kStoreBufBytes);  // 16 KB: two [128][16] fp32 slabs
    const uint64_t key_e = sign_key(fe.sa);  // sa carries the eps stream in this mode
    const float klg = fe.klg * (fe.kl_upstream != nullptr ? static_cast<float>(*fe.kl_upstream) : 1.f);
    const float inv_s2 = 1.f / (fe.prior_std * fe.prior_std);
    int iter = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
      const TileCoord tcd = tile_coord_g(t, num_m, num_n, raster_group);
      const int m0 = tcd.m_blk * kTileM + cta_rank * kBlockM;
      const int n0 = tcd.n_blk * kFoBlockN;
      const int m = m0 + row;
      const bool m_ok = m < M;
      const uint32_t acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const long long row_off = static_cast<long long>(m_ok ? m : 0) * N;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int tcol = grp_id * 64 + cc * 16;
        const int nc = n0 + tcol;
        // A thread of this epilogue owns a ROW of the tile (tcgen05.ld 32x32b), so touching mu / rho / dmu / drho straight
        // from its registers is 32 lines per warp instruction (measured: ~16 K LSU wavefronts per tile, as long as the
        // tile's MMAs).  Global traffic therefore uses a column-major thread mapping — 4 threads cover the chunk's 64
        // bytes of a row, a warp instruction 8 rows — and is transposed through the group's staging buffer (XOR-swizzled
        // 16-byte pieces: conflict-free for both mappings).
        float4 mq[4], rq[4];
        {
          const int gr = gtid >> 2, gc = gtid & 3;  // row within a 32-row pass, 16-byte piece within the 64-byte row
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int r = p * 32 + gr;
            const bool ok = m0 + r < M && nc + 4 * gc < N;  // N % 4 == 0
            const long long off = static_cast<long long>(m0 + r) * N + nc;
            mq[p] = ok ? __ldg(reinterpret_cast<const float4*>(fe.mu + off) + gc) : make_float4(0.f, 0.f, 0.f, 0.f);
            rq[p] = ok ? __ldg(reinterpret_cast<const float4*>(fe.rho + off) + gc) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t a0[16], a1[16];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * (2 * kFoBlockN) + tcol;
        tmem_ld_32x16(taddr, a0);
        if constexpr (kMode == 1) tmem_ld_32x16(taddr + kFoBlockN, a1);
        ptx::tmem_ld_wait();
        if constexpr (kMode == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) a1[j] = a0[j];  // one product feeds both gradients
        }
        if (cc == 3) {
          ptx::tc_fence_before();
          if (is_leader) ptx::mbar_arrive(&tmem_empty_bar[acc]);
          else ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        }
        float e[16], unused[16];
        noise16<false>(key_e, fe.sa.stream, row_off + nc, e, 0, 0, 0, unused);
        // stage the parameters (column-major mapping) and read them back row-per-thread
        {
          const int gr = gtid >> 2, gc = gtid & 3;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int r = p * 32 + gr;
            const uint32_t a = r * 64 + ((gc ^ ((r >> 1) & 3)) << 4);
            ptx::st_shared_v4(sbuf + a, __float_as_uint(mq[p].x), __float_as_uint(mq[p].y), __float_as_uint(mq[p].z), __float_as_uint(mq[p].w));
            ptx::st_shared_v4(sbuf + 8192 + a, __float_as_uint(rq[p].x), __float_as_uint(rq[p].y), __float_as_uint(rq[p].z), __float_as_uint(rq[p].w));
          }
        }
        ptx::named_bar_sync(bar_id, 128);
        float4 om4[4], or4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = row * 64 + ((j ^ ((row >> 1) & 3)) << 4);
          const float4 m4 = ld_shared_v4f(sbuf + a), r4 = ld_shared_v4f(sbuf + 8192 + a);
          const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
          float om[4], orr[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float ex;
            const float sigma = fo_sigma_from_rho_fast(rr[i], ex);
            const float sgm = __fdividef(ex, 1.f + ex);
            om[i] = __uint_as_float(a0[4 * j + i]) + klg * (mm[i] - fe.prior_mean) * inv_s2;
            orr[i] = (__uint_as_float(a1[4 * j + i]) * e[4 * j + i] + klg * (sigma * inv_s2 - __fdividef(1.f, sigma))) * sgm;
          }
          om4[j] = make_float4(om[0], om[1], om[2], om[3]);
          or4[j] = make_float4(orr[0], orr[1], orr[2], orr[3]);
        }
        ptx::named_bar_sync(bar_id, 128);  // every thread has read its parameters: the buffer takes the results
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = row * 64 + ((j ^ ((row >> 1) & 3)) << 4);
          ptx::st_shared_v4(sbuf + a, __float_as_uint(om4[j].x), __float_as_uint(om4[j].y), __float_as_uint(om4[j].z), __float_as_uint(om4[j].w));
          ptx::st_shared_v4(sbuf + 8192 + a, __float_as_uint(or4[j].x), __float_as_uint(or4[j].y), __float_as_uint(or4[j].z), __float_as_uint(or4[j].w));
        }
        ptx::named_bar_sync(bar_id, 128);
        {
          const int gr = gtid >> 2, gc = gtid & 3;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int r = p * 32 + gr;
            if (m0 + r < M && nc + 4 * gc < N) {
              const uint32_t a = r * 64 + ((gc ^ ((r >> 1) & 3)) << 4);
              const long long off = static_cast<long long>(m0 + r) * N + nc;
              reinterpret_cast<float4*>(fe.dmu + off)[gc] = ld_shared_v4f(sbuf + a);
              reinterpret_cast<float4*>(fe.drho + off)[gc] = ld_shared_v4f(sbuf + 8192 + a);
            }
          }
        }
        ptx::named_bar_sync(bar_id, 128);  // results read out: the next chunk may overwrite the buffer
      }
    }
   } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int grp_id = (warp_idx - 4) >> 2;
    const int ew = warp_idx & 3;
    const int row = ew * 32 + lane;
    const bool store_leader = ew == 0 && lane == 0;
    const uint32_t bar_id = 1 + grp_id;
    uint8_t* buf = smem_store + grp_id * kStoreBufBytes;
    const bool has_sb = fe.sb.on != 0 && fe.out2 != nullptr;
    const uint64_t key_a = sign_key(fe.sa), key_b = has_sb ? sign_key(fe.sb) : 0;
    const bool relu = fe.act == kActRelu;
    const bool has_mask = fe.mask_src != nullptr;
    int iter = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
      const TileCoord tcd = tile_coord_g(t, num_m, num_n, raster_group);
      const int m0 = tcd.m_blk * kTileM + cta_rank * kBlockM;
      const int n0 = tcd.n_blk * kFoBlockN;
      const int m = m0 + row;
      const bool m_ok = m < M;
      const uint32_t acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const long long nrow_a = (static_cast<long long>(m_ok ? m : 0) + fe.sa.row0) * N;
      const long long nrow_b = (static_cast<long long>(m_ok ? m : 0) + fe.sb.row0) * N;
      const int ncol0 = n0 + grp_id * 64;

      // producer-side ReLU mask source (this layer's x), one chunk ahead of its use
      uint4 xq[2];
      const __nv_bfloat16* x_row = nullptr;
      if (has_mask) {
        x_row = reinterpret_cast<const __nv_bfloat16*>(fe.mask_src) + static_cast<long long>(m_ok ? m : 0) * fe.mask_ld;
        if (m_ok) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (ncol0 + 8 * j < N) xq[j] = __ldg(reinterpret_cast<const uint4*>(x_row + ncol0) + j);
        }
      }
      // the staging buffer was handed to the TMA engine one tile ago: make sure it has been read out
      if (out_tma) {
        if (store_leader) ptx::tma_store_wait_read<0>();
        ptx::named_bar_sync(bar_id, 128);
      }
      // this thread's 64 sign bits per stream (its row, the group's 64 columns): before the accumulator wait
      uint64_t bits_a = 0, bits_b = 0;
      if (has_sb) sign_bits64<true>(key_a, fe.sa.stream, nrow_a + ncol0, bits_a, key_b, fe.sb.stream, nrow_b + ncol0, bits_b);
      else sign_bits64<false>(key_a, fe.sa.stream, nrow_a + ncol0, bits_a, 0, 0, 0, bits_b);
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();

#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int tcol = grp_id * 64 + cc * 16;
        const int nc = n0 + tcol;
        const uint32_t na = ~static_cast<uint32_t>(bits_a >> (16 * cc));  // bit j set: element j is -1
        const uint32_t nb = ~static_cast<uint32_t>(bits_b >> (16 * cc));
        uint32_t a0[16], a1[16];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * (2 * kFoBlockN) + tcol;
        tmem_ld_32x16(taddr, a0);
        tmem_ld_32x16(taddr + kFoBlockN, a1);
        ptx::tmem_ld_wait();
        if (cc == 3) {
          ptx::tc_fence_before();
          if (is_leader) ptx::mbar_arrive(&tmem_empty_bar[acc]);
          else ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        }
        const int n_valid = max(0, min(16, N - nc));  // 0, 8 or 16 (N % 8 == 0)
        const bool full = n_valid == 16;
        const bool active = m_ok && n_valid > 0;

        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = __uint_as_float(a0[j]) + __uint_as_float(a1[j] ^ ((na >> j) << 31));
        if (fe.bias != nullptr) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            if (nc + 4 * j4 < N) {  // N % 4 == 0: a 4-column vector is entirely inside or outside
              const float4 b = __ldg(reinterpret_cast<const float4*>(fe.bias + nc) + j4);
              v[4 * j4] += b.x; v[4 * j4 + 1] += b.y; v[4 * j4 + 2] += b.z; v[4 * j4 + 3] += b.w;
            }
          }
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (has_mask) {
          float xs[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) xs[j] = 0.f;
          if (active) {
            unpack_bf16x8(xq[0], xs);
            if (full) unpack_bf16x8(xq[1], xs + 8);
          }
          if (m_ok && cc + 1 < 4) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              if (nc + 16 + 8 * j < N) xq[j] = __ldg(reinterpret_cast<const uint4*>(x_row + nc + 16) + j);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = xs[j] > 0.f ? v[j] : 0.f;
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);

        if (has_sb && active) {
          uint32_t p2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            p2[j] = pk[j] ^ ((((nb >> (2 * j)) & 1u) << 15) | ((nb >> (2 * j + 1)) << 31));
          uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(fe.out2) + static_cast<long long>(m) * fe.out2_ld + nc);
          p[0] = make_uint4(p2[0], p2[1], p2[2], p2[3]);
          if (full) p[1] = make_uint4(p2[4], p2[5], p2[6], p2[7]);
        }
        if (fe.col_sum != nullptr) {
          // column sums of the bf16-rounded values (what the consumer of `out` reads)
          float q[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            q[2 * j] = (active && 2 * j < n_valid) ? __uint_as_float(pk[j] << 16) : 0.f;
            q[2 * j + 1] = (active && 2 * j + 1 < n_valid) ? __uint_as_float(pk[j] & 0xFFFF0000u) : 0.f;
          }
          col_sum16(q, lane);
          const int my_col = nc + col16_of_lane(lane);
          if ((lane & 1) == 0 && m0 + ew * 32 < M && my_col < N) atomicAdd(fe.col_sum + my_col, q[0]);
        }

        if (out == nullptr) continue;
        if (!out_tma) {
          if (active) {
            uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + static_cast<long long>(m) * out_ld + nc);
            p[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            if (full) p[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          continue;
        }
        const uint32_t buf_addr = ptx::smem_u32(buf) + row * 128;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int chunk = (cc * 2 + j) ^ (row & 7);
          ptx::st_shared_v4(buf_addr + chunk * 16, pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        if (cc == 3) {
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (store_leader) {
            if (ncol0 < N && m0 < M) ptx::tma_store_2d(&tm_out, buf, ncol0, m0);
            ptx::tma_store_commit();
          }
        }
      }
    }
    if (out_tma && store_leader) ptx::tma_store_wait_all<0>();
   }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp_idx == 2) ptx::tmem_dealloc<2>(tmem_base, kFoTmemCols);
}

}  // namespace tc
}  // namespace ed2
