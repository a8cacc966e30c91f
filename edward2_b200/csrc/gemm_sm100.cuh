// Persistent, warp-specialised tcgen05 GEMM for sm_100a with a fused, runtime-configurable epilogue.
//
//   D[M,N] = epilogue( A[M,K] * B[K,N] ),   bf16 operands, fp32 accumulation in TMEM.
//
// Every dense contraction of the Edward2 hot path (kernels K1-K8) is an instance of this
// kernel: the forward products X·W, the input gradients G·Wᵀ, the weight gradients XᵀG, the random
// feature projection cos(xW+b), the Laplace precision update ΦᵀΦ and the predictive variance ΦΣ.
// Operand layouts are expressed as "K-major" (contraction index contiguous) or "MN-major" (row/col index
// contiguous), so none of those products needs a transposed copy in HBM.
//
// Roles per CTA (256 threads):  warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4-7 = epilogue (TMEM -> registers -> fused math -> swizzled smem -> TMA store).
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA, double-buffered TMEM accumulator
// (tmem_full/tmem_empty) between MMA and epilogue, so the epilogue of tile i overlaps the MMAs of tile i+1.
// kCtaGroup == 2 pairs two SMs on one 256-row tile (tcgen05 cta_group::2): each CTA stages its own 128
// rows of A and half of the B tile; the leader CTA issues the MMAs for both.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "gemm_epilogue.h"
#include "ptx.cuh"

namespace ed2 {

namespace tc {

constexpr int kBlockM = 128;       // rows per CTA (UMMA M = 128 * kCtaGroup)
constexpr int kBlockK = 64;        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kStoreBufBytes = kBlockM * 128;  // one staging buffer: 128 rows x 128 B (SWIZZLE_128B)
constexpr int kNumStoreBufs = 2;
// Leave ~32 KB of the SM's 227 KB to other kernels: the issue-bound sampling / gradient-finishing kernels are launched
// on side streams and must be able to co-reside with a GEMM CTA (tensor cores + TMA leave the CUDA cores idle).
constexpr int kSmemBudget = 195 * 1024;

constexpr int kR1EpiWarps = 8;
constexpr int kR1EpiThreads = kR1EpiWarps * 32;
constexpr int kR1Threads = 128 + kR1EpiThreads;  // 4 control warps + 8 epilogue warps
constexpr int kR1Chunk = 16;                     // accumulator columns per TMEM load
constexpr int kR1StoreBufs = 4;                  // 2 per epilogue group, 128 rows x 64 bf16 columns each
constexpr int kR1TableFloats = 6 * 256;          // mu_s, sg_s, sgm_s, mu_r, sg_r, sgm_r for the tile's 256 columns
constexpr int kR1TableBytes = 2 * kR1TableFloats * 4;  // double-buffered by tile parity

// The fused rank-1 epilogues (gemm_rank1_epi.cuh) own the SM — there are no side-stream elementwise kernels left to
// co-reside with — and carry four staging buffers (two per epilogue group) plus the per-tile parameter tables.
constexpr int kSmemBudgetFused = 226 * 1024;

template <int BLOCK_N, int kCtaGroup, int kEpi = kEpiGeneric>
struct Config {
  static constexpr bool kFused = kEpi != kEpiGeneric;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kLoadN = BLOCK_N / kCtaGroup;
  static constexpr int kBBytes = kLoadN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStoreBufs = kFused ? kR1StoreBufs : kNumStoreBufs;
  static constexpr int kTableBytes = kFused ? kR1TableBytes : 0;
  static constexpr int kNumThreads = kFused ? kR1Threads : kThreads;
  static constexpr int kEpiThreadsTotal = kFused ? kR1EpiThreads : kEpiThreads;
  static constexpr int kFixedBytes = kStoreBufs * kStoreBufBytes + 1024 /*barriers*/ + 1024 /*align slack*/ + kTableBytes;
  static constexpr int kStagesRaw = ((kFused ? kSmemBudgetFused : kSmemBudget) - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages; 256 or 512 (power of two)
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns must be a power of 2");
};

// 64-bit shared-memory matrix descriptor (tcgen05): start address, leading/stride byte offsets (>>4),
// descriptor version 1, SWIZZLE_128B layout.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// Descriptor for one operand tile.
//  K-major : tile is [rows][64] bf16, 128 B per row, 8-row swizzle atoms 1024 B apart (SBO); LBO unused.
//  MN-major: tile is a sequence of [64 k][64 mn] blocks (one per TMA box), 8 KB apart (LBO);
//            inside a block, 8-k-row swizzle atoms are 1024 B apart (SBO).
template <int kMajor>
__device__ __forceinline__ uint64_t operand_desc(uint32_t smem_addr) {
  if constexpr (kMajor == kMajorK) {
    return make_smem_desc(smem_addr, 0, 8 * 128);
  } else {
    return make_smem_desc(smem_addr, kBlockK * 128, 8 * 128);
  }
}

// Descriptor start-address increment (in 16-byte units) per UMMA_K = 16 step.
template <int kMajor>
__device__ __forceinline__ constexpr uint32_t desc_k_step() {
  return kMajor == kMajorK ? (kUmmaK * 2) >> 4 : (kUmmaK * 128) >> 4;
}

// 32-bit instruction descriptor for kind::f16 with bf16 inputs and fp32 accumulation.
template <int kMajorA, int kMajorB>
__device__ __forceinline__ constexpr uint32_t make_idesc(int umma_m, int umma_n) {
  uint32_t d = 0;
  d |= 1u << 4;   // D format: f32
  d |= 1u << 7;   // A format: bf16
  d |= 1u << 10;  // B format: bf16
  d |= static_cast<uint32_t>(kMajorA) << 15;
  d |= static_cast<uint32_t>(kMajorB) << 16;
  d |= static_cast<uint32_t>(umma_n >> 3) << 17;
  d |= static_cast<uint32_t>(umma_m >> 4) << 24;
  return d;
}

struct TileCoord {
  int m_blk, n_blk;
};

// Grouped rasterisation: consecutive tiles walk down a band of kGroup M-blocks before moving to the next
// N-block, so that the A band and the current B panel stay L2-resident across the concurrently running CTAs.
__device__ __forceinline__ TileCoord tile_coord(int t, int num_m, int num_n) {
  constexpr int kGroup = 16;
  const int per_group = kGroup * num_n;
  const int g = t / per_group;
  const int first_m = g * kGroup;
  const int gsize = min(kGroup, num_m - first_m);
  const int r = t - g * per_group;
  TileCoord c;
  c.m_blk = first_m + r % gsize;
  c.n_blk = r / gsize;
  return c;
}

}  // namespace tc
}  // namespace ed2
#include "gemm_rank1_epi.cuh"
namespace ed2 {
namespace tc {

// t-th tile of the lower triangle (diagonal included) of a square nt x nt tile grid, row by row
__device__ __forceinline__ TileCoord tri_coord(int t) {
  int m = static_cast<int>((sqrtf(8.f * static_cast<float>(t) + 1.f) - 1.f) * 0.5f);
  while ((m + 1) * (m + 2) / 2 <= t) ++m;
  while (m * (m + 1) / 2 > t) --m;
  TileCoord c;
  c.m_blk = m;
  c.n_blk = t - m * (m + 1) / 2;
  return c;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Load 32 consecutive elements of a row (bounds- and alignment-aware) as fp32.
__device__ __forceinline__ void load_row32(const void* base, int dt, long long row_off, int n0, int n_valid,
                                           float (&f)[32]) {
  if (dt == kBF16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + row_off + n0;
    if (n_valid == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 x = __bfloat1622float2(h[t]);
          f[j * 8 + 2 * t] = x.x;
          f[j * 8 + 2 * t + 1] = x.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = j < n_valid ? __bfloat162float(p[j]) : 0.f;
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base) + row_off + n0;
    if (n_valid == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 q = __ldg(reinterpret_cast<const float4*>(p) + j);
        f[4 * j] = q.x;
        f[4 * j + 1] = q.y;
        f[4 * j + 2] = q.z;
        f[4 * j + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = j < n_valid ? p[j] : 0.f;
    }
  }
}

// Store 32 consecutive elements of a row directly to global memory.
__device__ __forceinline__ void store_row32(void* base, int dt, long long row_off, int n0, int n_valid,
                                            const float (&f)[32]) {
  if (dt == kBF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + row_off + n0;
    if (n_valid == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 q;
        q.x = pack_bf16(f[8 * j], f[8 * j + 1]);
        q.y = pack_bf16(f[8 * j + 2], f[8 * j + 3]);
        q.z = pack_bf16(f[8 * j + 4], f[8 * j + 5]);
        q.w = pack_bf16(f[8 * j + 6], f[8 * j + 7]);
        reinterpret_cast<uint4*>(p)[j] = q;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_valid) p[j] = __float2bfloat16_rn(f[j]);
    }
  } else {
    float* p = reinterpret_cast<float*>(base) + row_off + n0;
    if (n_valid == 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(p)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < n_valid) p[j] = f[j];
    }
  }
}


// Column sums of a 32-row x 32-column register tile (one row per lane): butterfly transpose-reduce, 31 shuffles;
// lane j ends with the sum of column j over the warp's 32 rows and issues one atomic.  All 32 rows of a warp must
// belong to the same group (the host only enables col_sum when group_rows is 0 or a multiple of 32).
__device__ __forceinline__ void col_sum_chunk(const EpiParams& ep, float (&v)[32], bool active, int warp_row0, int nc,
                                              int N, uint32_t lane) {
  if (!active) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  const int n = nc + static_cast<int>(lane);
  if (n < N) {
    const int grp = ep.group_rows > 0 ? warp_row0 / ep.group_rows : 0;
    atomicAdd(ep.col_sum + static_cast<long long>(grp) * ep.col_sum_ld + n, v[0]);
  }
}

// sin and cos of a random-feature phase (|v| up to a few hundred radians): two-constant Cody-Waite reduction to
// [-pi, pi] in fp32 (error <= |v| * 2^-24 ~ 1e-5 rad at |v| = 200) followed by the hardware sin / cos (abs error ~1e-6 on
// that interval).  ~10 instructions per element instead of the ~60 of sincosf's general path: with four epilogue warps
// the cosine epilogue of the feature GEMM was ALU-bound (32 K elements per tile).
__device__ __forceinline__ void sincos_phase(float v, float& s, float& c) {
  const float k = rintf(v * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, v);   // 2 pi rounded to fp32
  r = fmaf(-k, -1.7484555314695172e-07f, r);    // 2 pi - fp32(2 pi)
  s = __sinf(r);
  c = __cosf(r);
}

template <int kMajorA, int kMajorB, int BLOCK_N, int kCtaGroup, int kEpi = kEpiGeneric>
__global__ void __maxnreg__(kEpi == kEpiGeneric ? 128 : 168)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, int M, int N, int K, EpiParams ep,
                 const __grid_constant__ Rank1Epi r1) {
  using Cfg = Config<BLOCK_N, kCtaGroup, kEpi>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = kBlockM * kCtaGroup;
  constexpr int kUmmaM = kTileM;
  constexpr int kUmmaN = BLOCK_N;
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256, "BLOCK_N must be 64/128/192/256");

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + kStages * Cfg::kABytes;
  uint8_t* smem_store = smem_b + kStages * Cfg::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + Cfg::kStoreBufs * kStoreBufBytes);
  [[maybe_unused]] uint8_t* smem_table = reinterpret_cast<uint8_t*>(bars) + 1024;
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp_idx = __shfl_sync(0xffffffff, threadIdx.x / 32, 0);
  const uint32_t lane = ptx::lane_id();
  const uint32_t cta_rank = kCtaGroup == 2 ? ptx::cluster_ctarank() : 0;
  const bool is_leader = cta_rank == 0;

  if (kCtaGroup == 2) ptx::cluster_sync();

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
    if (ep.out_tma) ptx::prefetch_tensormap(&tm_out);
  }
  if (warp_idx == 1 && ptx::elect_one()) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], kCtaGroup);  // one producer arrive per CTA (on the leader's barrier)
      ptx::mbar_init(&empty_bar[i], 1);         // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);                         // one tcgen05.commit
      ptx::mbar_init(&tmem_empty_bar[i], kCtaGroup * Cfg::kEpiThreadsTotal);  // every epilogue thread of the pair
    }
    ptx::fence_barrier_init();
  } else if (warp_idx == 2) {
    ptx::tmem_alloc<kCtaGroup>(tmem_ptr_smem, Cfg::kTmemCols);
  }
  ptx::tc_fence_before();
  if (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_m = (M + kTileM - 1) / kTileM;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int num_clusters = gridDim.x / kCtaGroup;
  // work units: (tile, K slice).  lower_only enumerates the lower triangle; split_k cuts the K loop (host guarantees
  // split_k <= num_kb, so no slice is empty)
  const bool tri = ep.lower_only != 0;
  const int tiles_mn = tri ? num_m * (num_m + 1) / 2 : num_m * num_n;
  const int split_k = ep.split_k > 1 ? ep.split_k : 1;
  const int kb_per = (num_kb + split_k - 1) / split_k;
  const int num_tiles = tiles_mn * split_k;
  auto unit_tile = [&](int u) -> TileCoord { const int tt = u % tiles_mn; return tri ? tri_coord(tt) : tile_coord(tt, num_m, num_n); };
  auto unit_kb0 = [&](int u) -> int { return (u / tiles_mn) * kb_per; };
  auto unit_kb1 = [&](int u) -> int { return min(num_kb, (u / tiles_mn) * kb_per + kb_per); };

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const TileCoord tc = unit_tile(t);
        const int m0 = tc.m_blk * kTileM + cta_rank * kBlockM;
        const int n0 = tc.n_blk * BLOCK_N + cta_rank * Cfg::kLoadN;
        for (int kb = unit_kb0(t), kb_end = unit_kb1(t); kb < kb_end; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const int k0 = kb * kBlockK;
          uint8_t* sa = smem_a + stage * Cfg::kABytes;
          uint8_t* sb = smem_b + stage * Cfg::kBBytes;
          uint64_t* fb = &full_bar[stage];
          auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
            if constexpr (kCtaGroup == 1) ptx::tma_load_2d(dst, tm, fb, c0, c1);
            else ptx::tma_load_2d_pair(dst, tm, fb, c0, c1);
          };
          if constexpr (kMajorA == kMajorK) {
            load(sa, &tm_a, k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i) load(sa + i * (kBlockK * 128), &tm_a, m0 + i * 64, k0);
          }
          if constexpr (kMajorB == kMajorK) {
            load(sb, &tm_b, k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::kLoadN / 64; ++i) load(sb + i * (kBlockK * 128), &tm_b, n0 + i * 64, k0);
          }
          if (is_leader) ptx::mbar_arrive_expect_tx(fb, Cfg::kStageBytes * kCtaGroup);
          else ptx::mbar_arrive_cluster(fb, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (is_leader && ptx::elect_one()) {
      constexpr uint32_t idesc = make_idesc<kMajorA, kMajorB>(kUmmaM, kUmmaN);
      uint32_t stage = 0, phase = 0;
      int iter = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
        const uint32_t acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        const int kb_first = unit_kb0(t), kb_end = unit_kb1(t);
        for (int kb = kb_first; kb < kb_end; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = operand_desc<kMajorA>(ptx::smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = operand_desc<kMajorB>(ptx::smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            ptx::umma<kCtaGroup, false>(tmem_d, da + k * desc_k_step<kMajorA>(), db + k * desc_k_step<kMajorB>(),
                                        idesc, (kb > kb_first || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit<kCtaGroup>(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (kb == kb_end - 1) ptx::umma_commit<kCtaGroup>(&tmem_full_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx >= 4 && kEpi != kEpiGeneric) {
    // ------------------------------------------------------------------ fused rank-1 epilogue (256 threads)
    if constexpr (kEpi != kEpiGeneric)
      rank1_epilogue<kEpi, BLOCK_N, kCtaGroup>(tm_out, ep, r1, M, N, tmem_base, tmem_full_bar, tmem_empty_bar, smem_store,
                                               smem_table, warp_idx, lane, cta_rank, cluster_id, num_clusters, num_m, num_n);
  } else if (warp_idx >= 4) {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int ew = warp_idx - 4;  // == warp_idx % 4: the TMEM lane quadrant this warp may read
    const int row = ew * 32 + lane;
    const bool store_leader = (warp_idx == 4) && (lane == 0);
    uint32_t store_buf = 0;
    int iter = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
      const TileCoord tc = unit_tile(t);
      const int m0 = tc.m_blk * kTileM + cta_rank * kBlockM;
      const int n0 = tc.n_blk * BLOCK_N;
      const int m = m0 + row;
      const bool m_ok = m < M;
      const uint32_t acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const int grp = ep.group_rows > 0 ? (m_ok ? m / ep.group_rows : 0) : 0;
      float row_acc = 0.f;

      // ReLU-mask source (the producer layer's bf16 activations): its global loads are issued most of a 32-column chunk
      // AHEAD of their use, otherwise every chunk exposes a full global-load latency and short-K GEMMs (the dX of the
      // 1000-unit head) become epilogue-bound.
      const bool mask_fast = ep.mask_src != nullptr && ep.mask_dt == kBF16 && m_ok && (ep.mask_ld & 7) == 0 &&
                             (reinterpret_cast<uintptr_t>(ep.mask_src) & 15) == 0;
      const uint4* mask_row = reinterpret_cast<const uint4*>(
          reinterpret_cast<const __nv_bfloat16*>(ep.mask_src) + static_cast<long long>(m_ok ? m : 0) * ep.mask_ld + n0);
      uint4 mq[4];
      if (mask_fast && n0 + 32 <= N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) mq[j] = __ldg(mask_row + j);
      }

#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t vr[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N + c * 32, vr);
        ptx::tmem_ld_wait();
        if (c == BLOCK_N / 32 - 1) {
          // accumulator fully drained into registers: hand the TMEM stage back to the MMA issuer
          ptx::tc_fence_before();
          if (is_leader) ptx::mbar_arrive(&tmem_empty_bar[acc]);
          else ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
        }
        const int nc = n0 + c * 32;
        const int n_valid = max(0, min(32, N - nc));
        const bool active = m_ok && n_valid > 0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(vr[j]) * ep.alpha;

        if (split_k > 1) {  // K-slice partial: out += alpha * partial (fp32 reductions in L2), nothing else
          if (active) {
            float* o = reinterpret_cast<float*>(ep.out) + static_cast<long long>(m) * ep.out_ld + nc;
            // 16-byte vector reductions where the row is aligned: a quarter of the L2 atomic operations (the scalar form
            // made the short split-K launches of the SNGP head atomic-bound: 64 K-slices x 65536 elements per tile)
            if (n_valid == 32 && ((reinterpret_cast<uintptr_t>(ep.out) | (static_cast<uintptr_t>(ep.out_ld) << 2)) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < n_valid) atomicAdd(o + j, v[j]);
            }
          }
          continue;
        }
        if (ep.raw != nullptr && active && !ep.raw_neg_sin)
          store_row32(ep.raw, ep.raw_dt, static_cast<long long>(m) * ep.raw_ld, nc, n_valid, v);

        if (active) {
          if (ep.elem_scale != nullptr) {
            float s[32];
            load_row32(ep.elem_scale, ep.elem_scale_dt, static_cast<long long>(m) * ep.elem_scale_ld, nc, n_valid, s);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= s[j];
          }
          if (ep.col_scale != nullptr) {
            const float* p = ep.col_scale + static_cast<long long>(grp) * ep.col_scale_ld + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= (j < n_valid ? __ldg(p + j) : 0.f);
          }
          if (ep.addend != nullptr) {
            float s[32];
            load_row32(ep.addend, ep.addend_dt, static_cast<long long>(m) * ep.addend_ld, nc, n_valid, s);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(ep.beta, s[j], v[j]);
          }
          if (ep.col_bias != nullptr) {
            const float* p = ep.col_bias + static_cast<long long>(grp) * ep.col_bias_ld + nc;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (j < n_valid ? __ldg(p + j) : 0.f);
          }
          if (ep.act == kActRelu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (ep.act == kActCos && ep.raw_neg_sin && ep.raw != nullptr) {
            // phi = cos(v) * scale -> out,  -sin(v) * scale -> raw: one range reduction for both (sincosf), the raw tile
            // is stored in 8-column pieces so that only 8 extra registers are live
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              float sn[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float sv, cv;
                sincos_phase(v[8 * j8 + j], sv, cv);
                v[8 * j8 + j] = cv * ep.act_scale;
                sn[j] = -sv * ep.act_scale;
              }
              const int nv = n_valid - 8 * j8;
              if (nv > 0) {
                if (ep.raw_dt == kBF16) {
                  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(ep.raw) + static_cast<long long>(m) * ep.raw_ld + nc + 8 * j8;
                  if (nv >= 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(sn[0], sn[1]), pack_bf16(sn[2], sn[3]), pack_bf16(sn[4], sn[5]), pack_bf16(sn[6], sn[7]));
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (j < nv) p[j] = __float2bfloat16_rn(sn[j]);
                  }
                } else {
                  float* p = reinterpret_cast<float*>(ep.raw) + static_cast<long long>(m) * ep.raw_ld + nc + 8 * j8;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (j < nv) p[j] = sn[j];
                }
              }
            }
          } else if (ep.act == kActCos) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float sv, cv;
              sincos_phase(v[j], sv, cv);
              v[j] = cv * ep.act_scale;
            }
          }
        }

        if (mask_fast && n_valid == 32) {  // prefetched chunk
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&mq[j]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 x = __bfloat1622float2(h[t]);
              v[j * 8 + 2 * t] = x.x > 0.f ? v[j * 8 + 2 * t] : 0.f;
              v[j * 8 + 2 * t + 1] = x.y > 0.f ? v[j * 8 + 2 * t + 1] : 0.f;
            }
          }
        } else if (ep.mask_src != nullptr && active) {
          float s[32];
          load_row32(ep.mask_src, ep.mask_dt, static_cast<long long>(m) * ep.mask_ld, nc, n_valid, s);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = s[j] > 0.f ? v[j] : 0.f;
        }
        if (mask_fast && c + 1 < BLOCK_N / 32 && n0 + (c + 2) * 32 <= N) {
          // next chunk's mask: in flight during this chunk's column sums / staged store and the next TMEM load
#pragma unroll
          for (int j = 0; j < 4; ++j) mq[j] = __ldg(mask_row + (c + 1) * 4 + j);
        }
        if (ep.row_sum != nullptr && active) {
#pragma unroll
          for (int j = 0; j < 32; ++j) row_acc += (j < n_valid ? v[j] : 0.f);
        }
        if (!ep.out_tma) {
          if (active && ep.out != nullptr)
            store_row32(ep.out, ep.out_dt, static_cast<long long>(m) * ep.out_ld, nc, n_valid, v);
          if (ep.col_sum != nullptr) col_sum_chunk(ep, v, active, m0 + ew * 32, nc, N, lane);
          continue;
        }
        // ---- staged store: swizzled smem (conflict-free 16-byte chunks) + one TMA store per buffer.
        // bf16: one buffer = 64 columns (two 32-column passes); fp32: one buffer = 32 columns.
        const bool is_bf16 = ep.out_dt == kBF16;
        const int sub = is_bf16 ? (c & 1) : 0;
        if (sub == 0) {
          // the buffer we are about to overwrite must have been read out by its previous TMA store
          if (store_leader) ptx::tma_store_wait_read<kNumStoreBufs - 1>();
          ptx::named_bar_sync(1, kEpiThreads);
        }
        const uint32_t buf_addr = ptx::smem_u32(smem_store + store_buf * kStoreBufBytes) + row * 128;
        if (is_bf16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = (sub * 4 + j) ^ (row & 7);
            ptx::st_shared_v4(buf_addr + chunk * 16, pack_bf16(v[8 * j], v[8 * j + 1]),
                              pack_bf16(v[8 * j + 2], v[8 * j + 3]), pack_bf16(v[8 * j + 4], v[8 * j + 5]),
                              pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int chunk = j ^ (row & 7);
            ptx::st_shared_v4(buf_addr + chunk * 16, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                              __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          }
        }
        if (ep.col_sum != nullptr) col_sum_chunk(ep, v, active, m0 + ew * 32, nc, N, lane);
        if (!is_bf16 || sub == 1) {
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1, kEpiThreads);
          if (store_leader) {
            const int col0 = is_bf16 ? nc - 32 : nc;
            if (col0 < N && m0 < M) ptx::tma_store_2d(&tm_out, smem_store + store_buf * kStoreBufBytes, col0, m0);
            ptx::tma_store_commit();
          }
          store_buf = (store_buf + 1) % kNumStoreBufs;
        }
      }
      if (ep.row_sum != nullptr && m_ok) atomicAdd(ep.row_sum + m, row_acc);
    }
    if (ep.out_tma && store_leader) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  if (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp_idx == 2) ptx::tmem_dealloc<kCtaGroup>(tmem_base, Cfg::kTmemCols);
}

}  // namespace tc
}  // namespace ed2
