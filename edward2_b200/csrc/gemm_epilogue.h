// Operand-layout / dtype enums and the fused-epilogue parameter block of the tcgen05 GEMM.
#pragma once

namespace ed2 {

enum : int { kMajorK = 0, kMajorMN = 1 };
enum : int { kActNone = 0, kActRelu = 1, kActCos = 2 };
enum : int { kF32 = 0, kBF16 = 1 };

// Fused epilogue, applied per accumulator element (m, n) in this order:
//   v = alpha * acc
//   raw[m,n] = v                                   (optional second output: the un-scaled product)
//   v *= elem_scale[m,n]                           (Flipout sign_output / per-example rank-1 s)
//   v *= col_scale[(m / group_rows), n]            (BatchEnsemble gamma per member; group_rows=0: one row)
//   v += beta * addend[m,n]                        (Flipout mean term, momentum blend of the precision)
//   v += col_bias[(m / group_rows), n]
//   v = act(v)                                     (none | relu | cos(v) * act_scale)
//   v = mask_src[m,n] > 0 ? v : 0                  (optional: ReLU backward of the layer that produced A's consumer input)
//   col_sum[g(m), n] += sum_m v                    (optional; fp32 atomics: bias gradient of that layer)
//   out[m,n] = v                                   (skipped when out == nullptr)
//   row_sum[m] += sum_n v                          (optional; fp32 atomicAdd, one per row per tile)
struct EpiParams {
  float alpha;
  float beta;
  float act_scale;
  int act;
  int group_rows;
  const void* elem_scale;
  int elem_scale_ld;
  int elem_scale_dt;
  const float* col_scale;
  int col_scale_ld;
  const void* addend;
  int addend_ld;
  int addend_dt;
  const float* col_bias;
  int col_bias_ld;
  void* out;
  int out_ld;
  int out_dt;
  int out_tma;  // 1: `out` goes through swizzled smem + TMA store (needs 16-byte aligned pitch)
  void* raw;
  int raw_ld;
  int raw_dt;
  float* row_sum;
  const void* mask_src;
  int mask_ld;
  int mask_dt;
  float* col_sum;
  int col_sum_ld;
};

}  // namespace ed2
