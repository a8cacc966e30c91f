// Host side of the tcgen05 GEMM: tensor-map construction, configuration choice, launch.
#include "gemm.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gemm_sm100.cuh"
#include "gemm_flipout.cuh"
#include "status.h"

namespace ed2 {

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

// 2-D row-major tensor [rows][cols] with pitch `ld` elements; box = [box_rows][box_cols]; 128-byte swizzle.
bool make_map(CUtensorMap* tm, const void* ptr, int dt, long long rows, long long cols, long long ld, int box_rows,
              int box_cols) {
  EncodeFn enc = get_encode_fn();
  if (!enc) return false;
  const int es = dt == kBF16 ? 2 : 4;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tma_ok(const void* p, long long ld, int es) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ((ld * es) & 15) == 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    // Data-parallel runs leave a few SMs to the NCCL kernels: a persistent GEMM CTA holds ~194 KB of shared memory, so
    // a collective's CTAs cannot co-reside with it and would otherwise only run between GEMMs.
    if (const char* e = std::getenv("ED2_GEMM_MAX_CTAS")) {
      const int cap = std::atoi(e);
      if (cap >= 2 && cap < n) n = cap & ~1;
    }
  }
  return n;
}

template <int MA, int MB, int BN, int CG, int EPI = kEpiGeneric>
int launch_one(const GemmDesc& g, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
               cudaStream_t stream) {
  using Cfg = tc::Config<BN, CG, EPI>;
  auto kern = tc::gemm_bf16_kernel<MA, MB, BN, CG, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes,
                                           cudaGetErrorString(e));
    // Ask for the largest shared-memory carveout: the CTA itself needs ~194 KB, and the remaining ~32 KB are what lets
    // blocks of the side-stream kernels (sampling, KL, gradient finishing) co-reside with it on the same SM.
    static const bool max_carveout = [] { const char* e = std::getenv("ED2_GEMM_CARVEOUT"); return e == nullptr || std::atoi(e) != 0; }();
    if (max_carveout) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  const int tile_m = tc::kBlockM * CG;
  const int nm = (g.M + tile_m - 1) / tile_m;
  const int num_tiles = (g.ep.lower_only ? nm * (nm + 1) / 2 : nm * ((g.N + BN - 1) / BN)) * std::max(1, g.ep.split_k);
  int clusters = num_sms() / CG;
  if (g.max_ctas > 0) clusters = std::min(clusters, std::max(1, g.max_ctas / CG));
  clusters = std::min(clusters, num_tiles);
  // the smallest grid that needs the same number of waves: 256 tiles on 74 SM pairs are 4 waves, and so they are on 64
  // pairs — the 20 SMs left over run the side-stream kernels (weight casts, KL, the gradient exchange) that cannot
  // co-reside with a persistent GEMM CTA
  // ED2_GEMM_MINWAVE=0 keeps the full grid: in data-parallel runs the exchange's SM kernels (persistent grids) would
  // otherwise start on the few free SMs while a GEMM runs and crawl there instead of running at full width between two
  // GEMMs (measured at N = 8: 1.92 ms per step with the reduced grids)
  static const bool min_wave = [] { const char* e = std::getenv("ED2_GEMM_MINWAVE"); return e == nullptr || std::atoi(e) != 0; }();
  if (min_wave) {
    const int waves = (num_tiles + clusters - 1) / clusters;
    clusters = (num_tiles + waves - 1) / waves;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(Cfg::kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const Rank1Epi kNoRank1{};
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, to, g.M, g.N, g.K, g.ep, g.r1 != nullptr ? *g.r1 : kNoRank1);
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return ED2_OK;
}

template <int MA, int MB>
int dispatch_cfg(const GemmDesc& g, int bn, int cg, const CUtensorMap& ta, const CUtensorMap& tb,
                 const CUtensorMap& to, cudaStream_t s) {
  if (bn == 128 && cg == 1) return launch_one<MA, MB, 128, 1>(g, ta, tb, to, s);
  if (bn == 256 && cg == 1) return launch_one<MA, MB, 256, 1>(g, ta, tb, to, s);
  if (bn == 128 && cg == 2) return launch_one<MA, MB, 128, 2>(g, ta, tb, to, s);
  if (bn == 256 && cg == 2) return launch_one<MA, MB, 256, 2>(g, ta, tb, to, s);
  return set_error(ED2_ERR_BAD_ARG, "unsupported gemm config block_n=%d cta_group=%d", bn, cg);
}

int default_cta_group() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ED2_GEMM_CTA_GROUP");
    v = e ? std::atoi(e) : 2;
    if (v != 1 && v != 2) v = 2;
  }
  return v;
}

}  // namespace

bool gemm_rank1_fusable(int M, int N, int group_rows) {
  // one 2-CTA 256x256 configuration; the 128 rows of a CTA's tile must belong to one member (its per-column
  // parameters are staged once per tile); 4-column noise blocks, 8-column vector stores
  return M > 128 && N % 8 == 0 && (group_rows == 0 || group_rows % 128 == 0);
}

int gemm_bf16_accumulate(const GemmDesc& g_in, cudaStream_t stream) {
  GemmDesc g = g_in;
  g.ep.split_k = -1;
  g.accumulate = 1;
  return gemm_bf16(g, stream);
}

int gemm_bf16(const GemmDesc& g_in, cudaStream_t stream) {
  GemmDesc g = g_in;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return set_error(ED2_ERR_BAD_SHAPE, "gemm: empty shape %d %d %d", g.M, g.N, g.K);
  if (!tma_ok(g.A, g.lda, 2) || !tma_ok(g.B, g.ldb, 2))
    return set_error(ED2_ERR_BAD_ARG, "gemm: A/B must be 16-byte aligned with a pitch that is a multiple of 8 bf16 "
                                      "(lda=%d ldb=%d)", g.lda, g.ldb);
  int bn = g.block_n ? g.block_n : (g.N > 128 ? 256 : 128);
  int cg = g.cta_group ? g.cta_group : default_cta_group();
  if (cg == 2 && g.M <= 128) cg = 1;
  if (g.epi_mode != kEpiGeneric) {
    if (g.r1 == nullptr || !gemm_rank1_fusable(g.M, g.N, g.ep.group_rows))
      return set_error(ED2_ERR_BAD_ARG, "gemm: fused rank-1 epilogue not available for M=%d N=%d group_rows=%d", g.M,
                       g.N, g.ep.group_rows);
    bn = 256; cg = 2;
    // few output tiles (the 1000-unit head at 2048 rows: 8 x 4 = 32 tiles of 256 x 256 for 74 SM pairs): 128-column tiles
    // double the number of work units (forward epilogue only)
    const long long tiles256 = static_cast<long long>((g.M + 255) / 256) * ((g.N + 255) / 256);
    if (g.epi_mode == kEpiRank1Fwd && tiles256 * 2 <= num_sms() / 2) bn = 128;
    // (measured, not kept: 128-column tiles for the short-K, epilogue-bound launches — 512 units on 74 pairs instead of
    // 256 on 64 — leave the dX launch of the 1000-unit head at 148 us: the bound is the epilogue's total instruction
    // work, not tile granularity)
  }
  if (g.ep.lower_only) {
    if (g.M != g.N || g.epi_mode != kEpiGeneric) return set_error(ED2_ERR_BAD_ARG, "gemm: lower_only needs a square generic product");
    if (!(cg == 2 && bn == 256) && !(cg == 1 && bn == 128)) { cg = g.M > 128 ? 2 : 1; bn = cg == 2 ? 256 : 128; }
  }
  if (g.ep.split_k < 0) {  // auto: enough K slices to give every SM pair a work unit, each at least 8 k-blocks long
    const int tile_m = tc::kBlockM * cg;
    const int nm = (g.M + tile_m - 1) / tile_m, nn = (g.N + bn - 1) / bn;
    const int tiles = g.ep.lower_only ? nm * (nm + 1) / 2 : nm * nn;
    const int units = num_sms() / cg;
    const int num_kb = (g.K + tc::kBlockK - 1) / tc::kBlockK;
    int sk = 1;
    if (tiles * 2 <= units) sk = std::max(1, std::min(units / tiles, num_kb / 8));
    g.ep.split_k = sk;
  }
  if (g.accumulate && g.ep.split_k <= 1) {  // out += alpha * AB without K slices: read-modify-write in the epilogue
    g.ep.addend = g.ep.out; g.ep.addend_ld = g.ep.out_ld; g.ep.addend_dt = g.ep.out_dt; g.ep.beta = 1.f;
  }
  const int load_n = bn / cg;

  CUtensorMap ta, tb, to;
  std::memset(&to, 0, sizeof(to));
  bool ok = true;
  if (g.major_a == kMajorK) ok &= make_map(&ta, g.A, kBF16, g.M, g.K, g.lda, tc::kBlockM, tc::kBlockK);
  else ok &= make_map(&ta, g.A, kBF16, g.K, g.M, g.lda, tc::kBlockK, 64);
  if (g.major_b == kMajorK) ok &= make_map(&tb, g.B, kBF16, g.N, g.K, g.ldb, load_n, tc::kBlockK);
  else ok &= make_map(&tb, g.B, kBF16, g.K, g.N, g.ldb, tc::kBlockK, 64);
  if (!ok) return set_error(ED2_ERR_CUDA, "cuTensorMapEncodeTiled failed for a GEMM operand");

  const int oes = g.ep.out_dt == kBF16 ? 2 : 4;
  g.ep.out_tma = 0;
  if (g.ep.split_k > 1) {
    const int num_kb = (g.K + tc::kBlockK - 1) / tc::kBlockK;
    g.ep.split_k = std::min(g.ep.split_k, num_kb);
    if (g.ep.out == nullptr || g.ep.out_dt != kF32 || g.epi_mode != kEpiGeneric)
      return set_error(ED2_ERR_BAD_ARG, "gemm: split_k accumulates into an fp32 out");
    g.no_tma_store = 1;
  }
  if (g.ep.out == nullptr && g.ep.row_sum == nullptr && g.epi_mode != kEpiRank1Bwd)
    return set_error(ED2_ERR_BAD_ARG, "gemm: out is null");
  if (g.ep.out != nullptr && !g.no_tma_store && tma_ok(g.ep.out, g.ep.out_ld, oes)) {
    if (make_map(&to, g.ep.out, g.ep.out_dt, g.M, g.N, g.ep.out_ld, tc::kBlockM, 128 / oes)) g.ep.out_tma = 1;
  }

  if (g.epi_mode == kEpiRank1Fwd) {
    if (g.major_a != kMajorK || g.major_b != kMajorMN) return set_error(ED2_ERR_BAD_ARG, "rank-1 fwd epilogue: layout");
    if (bn == 128) return launch_one<kMajorK, kMajorMN, 128, 2, kEpiRank1Fwd>(g, ta, tb, to, stream);
    return launch_one<kMajorK, kMajorMN, 256, 2, kEpiRank1Fwd>(g, ta, tb, to, stream);
  }
  if (g.epi_mode == kEpiRank1Bwd) {
    if (g.major_a != kMajorK || g.major_b != kMajorK) return set_error(ED2_ERR_BAD_ARG, "rank-1 bwd epilogue: layout");
    return launch_one<kMajorK, kMajorK, 256, 2, kEpiRank1Bwd>(g, ta, tb, to, stream);
  }
  if (g.major_a == kMajorK && g.major_b == kMajorK) return dispatch_cfg<kMajorK, kMajorK>(g, bn, cg, ta, tb, to, stream);
  if (g.major_a == kMajorK && g.major_b == kMajorMN) return dispatch_cfg<kMajorK, kMajorMN>(g, bn, cg, ta, tb, to, stream);
  if (g.major_a == kMajorMN && g.major_b == kMajorK) return dispatch_cfg<kMajorMN, kMajorK>(g, bn, cg, ta, tb, to, stream);
  return dispatch_cfg<kMajorMN, kMajorMN>(g, bn, cg, ta, tb, to, stream);
}

bool gemm_flipout_fusable(int M, int N, int K) { return M > 0 && N % 8 == 0 && K % 8 == 0; }

namespace {
template <int MA, int MB, int MODE>
int launch_flipout(const FlipoutGemm& g, const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tb0,
                   const CUtensorMap& tb1, const CUtensorMap& to, int out_tma, cudaStream_t stream) {
  auto kern = tc::gemm_flipout_kernel<MA, MB, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kFoSmemBytes);
    if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "cudaFuncSetAttribute(flipout smem=%d): %s", tc::kFoSmemBytes,
                                           cudaGetErrorString(e));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  const int nm = (g.M + 2 * tc::kBlockM - 1) / (2 * tc::kBlockM);
  const int nn = (g.N + tc::kFoBlockN - 1) / tc::kFoBlockN;
  const int clusters = std::min(num_sms() / 2, nm * nn);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(tc::kFoThreads);
  cfg.dynamicSmemBytes = tc::kFoSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // band height of the rasterisation (gemm_flipout.cuh: tile_coord_g): ~sqrt(clusters / 2) M-blocks, or one wave per
  // band when the output is narrow
  int group = 8;
  if (nn * group < clusters) group = (clusters + nn - 1) / nn;
  group = std::max(1, std::min(group, nm));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta0, ta1, tb0, tb1, to, g.M, g.N, g.K, g.out, g.out_ld, out_tma, group, g.epi);
  if (e != cudaSuccess) return set_error(ED2_ERR_CUDA, "flipout gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return ED2_OK;
}
}  // namespace

int gemm_flipout(const FlipoutGemm& g, cudaStream_t stream) {
  if (!gemm_flipout_fusable(g.M, g.N, g.K))
    return set_error(ED2_ERR_BAD_SHAPE, "flipout gemm: needs N, K multiples of 8 (M=%d N=%d K=%d)", g.M, g.N, g.K);
  if (!tma_ok(g.A0, g.lda0, 2) || !tma_ok(g.A1, g.lda1, 2) || !tma_ok(g.B0, g.ldb0, 2) || !tma_ok(g.B1, g.ldb1, 2))
    return set_error(ED2_ERR_BAD_ARG, "flipout gemm: operands must be 16-byte aligned with a pitch that is a multiple of 8 bf16");
  if (g.mode == 1 || g.mode == 2) {
    if (g.major_a != kMajorMN || g.major_b != kMajorMN) return set_error(ED2_ERR_BAD_ARG, "flipout gemm: weight-gradient mode takes MN-major operands");
    if (g.epi.mu == nullptr || g.epi.rho == nullptr || g.epi.dmu == nullptr || g.epi.drho == nullptr)
      return set_error(ED2_ERR_BAD_ARG, "flipout gemm: weight-gradient mode needs mu / rho / dmu / drho");
    for (const void* p : {static_cast<const void*>(g.epi.mu), static_cast<const void*>(g.epi.rho),
                          static_cast<const void*>(g.epi.dmu), static_cast<const void*>(g.epi.drho)})
      if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) return set_error(ED2_ERR_BAD_ARG, "flipout gemm: mu / rho / dmu / drho must be 16-byte aligned");
    CUtensorMap ta0, ta1, tb0, tb1, to;
    std::memset(&to, 0, sizeof(to));
    const bool ok = make_map(&ta0, g.A0, kBF16, g.K, g.M, g.lda0, tc::kBlockK, 64) &&
                    make_map(&ta1, g.A1, kBF16, g.K, g.M, g.lda1, tc::kBlockK, 64) &&
                    make_map(&tb0, g.B0, kBF16, g.K, g.N, g.ldb0, tc::kBlockK, 64) &&
                    make_map(&tb1, g.B1, kBF16, g.K, g.N, g.ldb1, tc::kBlockK, 64);
    if (!ok) return set_error(ED2_ERR_CUDA, "cuTensorMapEncodeTiled failed for a flipout GEMM operand");
    if (g.mode == 2) return launch_flipout<kMajorMN, kMajorMN, 2>(g, ta0, ta1, tb0, tb1, to, 0, stream);
    return launch_flipout<kMajorMN, kMajorMN, 1>(g, ta0, ta1, tb0, tb1, to, 0, stream);
  }
  if (g.major_a != kMajorK) return set_error(ED2_ERR_BAD_ARG, "flipout gemm: the summed mode takes K-major A operands");
  if (g.out == nullptr && g.epi.out2 == nullptr && g.epi.col_sum == nullptr)
    return set_error(ED2_ERR_BAD_ARG, "flipout gemm: no output");
  if (g.epi.out2 != nullptr && (!tma_ok(g.epi.out2, g.epi.out2_ld, 2)))
    return set_error(ED2_ERR_BAD_ARG, "flipout gemm: out2 must be 16-byte aligned with a pitch that is a multiple of 8 bf16");
  if (g.epi.mask_src != nullptr && !tma_ok(g.epi.mask_src, g.epi.mask_ld, 2))
    return set_error(ED2_ERR_BAD_ARG, "flipout gemm: mask_src must be 16-byte aligned with a pitch that is a multiple of 8 bf16");
  if (g.epi.bias != nullptr && (reinterpret_cast<uintptr_t>(g.epi.bias) & 15) != 0)
    return set_error(ED2_ERR_BAD_ARG, "flipout gemm: bias must be 16-byte aligned");
  CUtensorMap ta0, ta1, tb0, tb1, to;
  std::memset(&to, 0, sizeof(to));
  bool ok = make_map(&ta0, g.A0, kBF16, g.M, g.K, g.lda0, tc::kBlockM, tc::kBlockK) &&
            make_map(&ta1, g.A1, kBF16, g.M, g.K, g.lda1, tc::kBlockM, tc::kBlockK);
  if (g.major_b == kMajorK) {
    ok = ok && make_map(&tb0, g.B0, kBF16, g.N, g.K, g.ldb0, tc::kFoBlockN / 2, tc::kBlockK) &&
         make_map(&tb1, g.B1, kBF16, g.N, g.K, g.ldb1, tc::kFoBlockN / 2, tc::kBlockK);
  } else {
    ok = ok && make_map(&tb0, g.B0, kBF16, g.K, g.N, g.ldb0, tc::kBlockK, 64) &&
         make_map(&tb1, g.B1, kBF16, g.K, g.N, g.ldb1, tc::kBlockK, 64);
  }
  if (!ok) return set_error(ED2_ERR_CUDA, "cuTensorMapEncodeTiled failed for a flipout GEMM operand");
  int out_tma = 0;
  if (g.out != nullptr) {
    if (!tma_ok(g.out, g.out_ld, 2)) return set_error(ED2_ERR_BAD_ARG, "flipout gemm: out must be 16-byte aligned, pitch % 8 == 0");
    if (make_map(&to, g.out, kBF16, g.M, g.N, g.out_ld, tc::kBlockM, 64)) out_tma = 1;
  }
  if (g.major_b == kMajorK) return launch_flipout<kMajorK, kMajorK, 0>(g, ta0, ta1, tb0, tb1, to, out_tma, stream);
  return launch_flipout<kMajorK, kMajorMN, 0>(g, ta0, ta1, tb0, tb1, to, out_tma, stream);
}

}  // namespace ed2
