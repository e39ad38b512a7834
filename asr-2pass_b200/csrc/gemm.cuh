// tcgen05 / TMEM / TMA GEMM:  C[M,N] = A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major), fp32 accumulate,
// fused epilogue.  Replaces every MatMul/Gemm node of the reference's ONNX graph
// (Ort::Session::Run, onnxruntime/src/paraformer.cpp:541): QKV / out / FFN projections, decoder
// projections, the predictor's k=3 convolution (as three row-shifted K passes) and the vocabulary
// projection with a fused greedy argmax (FindMax, onnxruntime/src/util.cpp:63-74).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pf {

struct GemmEpilogue {
  const float* bias = nullptr;            // [N]
  const float* res_f32 = nullptr;         // [M, ld_res]  added to the result
  int ld_res = 0;
  const __nv_bfloat16* add_bf16 = nullptr;  // [M, ld_add] added to the result (FSMN memory)
  int ld_add = 0;
  float* out_f32 = nullptr;               // [M, ld_out_f32]
  int ld_out_f32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;      // [M, ld_out_bf16]
  int ld_out_bf16 = 0;
  int relu = 0;                           // 1: after bias, 2: after every addend
  int dbg = 0;                            // micro-benchmark / tests only: 1 = skip global stores, 2 = skip the whole epilogue body,
                                          // 3 = bf16 TMA epilogue without TMEM reads, 4 = force the general epilogue
  unsigned long long* argmax = nullptr;   // [M] packed (ordered value << 32 | ~index); caller zero-fills
  // LayerNorm folded across two GEMMs (decoder feed-forward: h = relu(y W1 + b1); LN(h) W2 without a LayerNorm pass over h):
  //   producer (bf16 TMA epilogue only): row_stats_out[row * stats_slots + col / 128] = (sum, sum of squares) of the row's 128
  //     ROUNDED outputs of that column span -- plain stores, one slot per span, so the result does not depend on scheduling;
  //   consumer (fp32 TMA epilogue only), W pre-multiplied by gamma along K:
  //     out = rstd * (acc - mean * ln_csum[col]) + bias[col], mean / rstd from the row's ln_slots partial sums over ln_dim
  //     columns, ln_csum[col] = sum_k W'[col][k] (of the rounded folded weights), bias = sum_k beta[k] W[col][k].
  float2* row_stats_out = nullptr;
  int stats_slots = 0;
  const float2* ln_stats = nullptr;
  int ln_slots = 0, ln_dim = 0;
  float ln_eps = 0.f;
  const float* ln_csum = nullptr;         // [N]
};

struct GemmProblem {
  const __nv_bfloat16* A = nullptr;  // [rows_a, lda]
  int64_t rows_a = 0;                // rows the A buffer holds (>= M; 0 = M): tensor-map extents are whole tiles inside it
  int64_t rows_c = 0;                // rows the output buffers hold (0 = M: exact extent, the last tile's stores are clipped at M)
  int lda = 0;
  const __nv_bfloat16* W = nullptr;  // [N, ldw]
  int ldw = 0;
  int M = 0, N = 0, K = 0;
  const int* m_dev = nullptr;        // optional: M read on device (decoder token count)
  // convolution-as-GEMM: K is split in passes of a_k_wrap columns; pass p reads A rows shifted by
  // (p + a_row_shift0).  a_k_wrap = 0 -> plain GEMM.
  int a_k_wrap = 0;
  int a_row_shift0 = 0;
  int f16 = 0;                       // 0: A, W, the 16-bit output and addend are bf16; 1: IEEE fp16 (same storage type in the signatures)
};

// Returns cudaError_t as int.  `num_sms` caps the persistent grid.
int gemm_bf16_tcgen05(const GemmProblem& p, const GemmEpilogue& e, int num_sms, cudaStream_t stream);

// 2-D bf16 row-major tensor map with 128-byte swizzle, box = [box_rows x 64 elements].
int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                         uint32_t box_rows, uint32_t box_cols = 64);

// Same without swizzle: the box lands row-major in shared memory ([box_rows][box_cols], box_cols <= 256).
int make_tmap_bf16_plain(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                         uint32_t box_rows, uint32_t box_cols);

__device__ __forceinline__ unsigned long long argmax_pack(float v, int idx) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

}  // namespace pf
