// Variable-length multi-head attention on tcgen05/TMEM (head dim 128): persistent CTAs walking
// (segment, 128-row query tile) work items, all heads of an item pipelined through one CTA.  Serves both the SAN-M encoder self-attention (q,k,v slices of
// the fused QKV buffer) and the decoder cross-attention (q from decoder tokens, k/v from the encoder
// memory).  Replaces the MatMul-Softmax-MatMul subgraphs of the reference's ONNX model
// (Ort::Session::Run, onnxruntime/src/paraformer.cpp:541).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pf {

struct AttnWork {  // one (segment, query tile); built on the host from the known frame counts
  int seg;
  int q0;
};

struct AttnProblem {
  const __nv_bfloat16* q = nullptr;   // [q_rows, ldq]; head h at columns q_col0 + 128 h
  int64_t q_rows = 0;
  int ldq = 0, q_col0 = 0;
  const __nv_bfloat16* kv = nullptr;  // [kv_rows, ldkv]; k at k_col0 + 128 h, v at v_col0 + 128 h
  int64_t kv_rows = 0;
  int ldkv = 0, k_col0 = 0, v_col0 = 0;
  __nv_bfloat16* out = nullptr;       // [q_rows, ldo]; head h at columns 128 h
  int ldo = 0;
  const int* q_row_off = nullptr;     // [n_seg] first query row of each segment   (device)
  const int* q_len = nullptr;         // [n_seg] query rows per segment            (device)
  const int* kv_row_off = nullptr;    // [n_seg] first key row                     (device)
  const int* kv_len = nullptr;        // [n_seg] keys per segment                  (device)
  const AttnWork* work = nullptr;     // [n_work]                                  (device)
  int n_work = 0;
  int n_heads = 4;
  float scale = 0.08838834764831845f;  // 128^-1/2
  int f16 = 0;                         // 0: q, kv, out are bf16; 1: IEEE fp16 (the engine's precision 1)
  int num_sms = 0;                     // SMs of the device (0 -> 148): the persistent grid is 2 CTAs per SM
};

int attention_tcgen05(const AttnProblem& p, cudaStream_t stream);

// Plain CUDA-core implementation of the same contract.  NOT on the product path: it exists so the GPU
// test-suite can cross-check the tcgen05 kernel at sizes the CPU oracle is too slow for.
int attention_check_kernel(const AttnProblem& p, cudaStream_t stream);

}  // namespace pf
