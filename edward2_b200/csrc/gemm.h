// Internal interface of the tcgen05 GEMM (see gemm_sm100.cuh for the kernel and the epilogue contract).
#pragma once
#include <cuda_runtime.h>
#include "gemm_epilogue.h"

namespace ed2 {

struct GemmDesc {
  const void* A;  // bf16. K-major: [M][lda] (k contiguous); MN-major: [K][lda] (m contiguous)
  int lda;
  int major_a;
  const void* B;  // bf16. K-major: [N][ldb] (k contiguous); MN-major: [K][ldb] (n contiguous)
  int ldb;
  int major_b;
  int M, N, K;
  EpiParams ep;
  int block_n;       // 0 = auto, else 128 | 256
  int cta_group;     // 0 = auto, else 1 | 2
  int max_ctas;      // 0 = all SMs
  int no_tma_store;  // force the direct-store epilogue (testing)
  int accumulate = 0;               // out (fp32, pre-initialised) += alpha * A B   (see gemm_bf16_accumulate)
  int epi_mode = kEpiGeneric;       // kEpiRank1Fwd / kEpiRank1Bwd: fused rank-1 epilogue described by *r1
  const Rank1Epi* r1 = nullptr;
};

// true when gemm_bf16 can run the fused rank-1 epilogue for this shape (else the caller takes the unfused path)
bool gemm_rank1_fusable(int M, int N, int group_rows);

int gemm_bf16(const GemmDesc& g, cudaStream_t stream);  // tcgen05 tensor cores
// out += alpha * A B into a pre-initialised fp32 `out`: K slices with atomics when the product has few output tiles and
// a long K, otherwise one read-modify-write epilogue (ep.lower_only honoured)
int gemm_bf16_accumulate(const GemmDesc& g, cudaStream_t stream);
int gemm_f32(const GemmDesc& g, cudaStream_t stream);   // fp32 FFMA path
int gemm(int dtype, const GemmDesc& g, cudaStream_t stream);

// DenseFlipout as one launch with two TMEM accumulators (gemm_flipout.cuh): out = post(A0 B0 + (A1 B1) ∘ sa), bf16.
// A0 / A1 are K-major [M][K]; B0 / B1 share one layout (K-major [N][K] or MN-major [K][N]).
struct FlipoutGemm {
  const void* A0; int lda0;
  const void* A1; int lda1;
  const void* B0; int ldb0;
  const void* B1; int ldb1;
  int major_b;
  int M, N, K;
  void* out; int out_ld;  // bf16, optional when epi.out2 / epi.col_sum carry the result
  FlipoutEpi epi;
  int major_a = kMajorK;
  int mode = 0;           // 1: weight-gradient pair finished in the epilogue (epi.mu / rho / dmu / drho; A and B MN-major)
                          // 2: the same epilogue on the single product A0·B0 (A1 / B1 = A0 / B0, not staged)
};
bool gemm_flipout_fusable(int M, int N, int K);
int gemm_flipout(const FlipoutGemm& g, cudaStream_t stream);

}  // namespace ed2
