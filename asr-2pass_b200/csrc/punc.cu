// CT-Transformer punctuation network on the GPU: what CTTransformer::Infer runs through its onnxruntime
// session (onnxruntime/src/ct-transformer.cpp:164-203) -- int32 token ids in, one punctuation class per token out -- for MANY
// token sequences per call.  The reference calls its session once per 20-token mini-sentence of one request; every call here can
// carry the current mini-sentence of every request that is being punctuated (funasr_b200::CTTransformerB200::AddPuncBatch), so the
// ~40 small launches of a pass are shared by all of them.
//
//   ids -> Embedding * sqrt(D) + sinusoidal position (from 1, per sequence) -> L x { LN -> QKV GEMM -> [FSMN(K) + id on V,
//   H-head softmax attention] -> out-proj GEMM (+x +memory) -> LN -> FFN GEMM ReLU -> FFN GEMM (+x) } -> LN -> Linear(D, n_punc)
//   -> first maximum over classes [0, n_punc - 1)   (Argmax(first, first + CANDIDATE_NUM - 1), ct-transformer.cpp:191-195)
//
// GEMMs run on the tcgen05 kernel (gemm.cu; bf16 operands, fp32 accumulate, leading dimensions zero-padded to multiples of 8); the
// residual stream, LN statistics, softmax and the classifier output are fp32.  The other kernels here are small and bandwidth /
// latency bound: sequences are tens to hundreds of tokens.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "engine.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "model_dir.h"

using namespace pf;

namespace {

constexpr int P_MAX_POS = 4096;     // longest single sequence (position table rows)
constexpr int P_MAX_SEQ = 16384;    // sequences per call
constexpr int P_QTILE = 16;         // queries per attention block (4 warps x 4)
constexpr int P_MAX_DK = 64;

struct PLinear {
  __nv_bfloat16* w = nullptr;   // [out_p, in_p] row-major, zero padded
  float* b = nullptr;           // [out_p]
  int out = 0, in = 0;          // padded sizes
};

// x[r, :] = embed[id] * sqrt(D) + pe[t]
__global__ void __launch_bounds__(128)
punc_embed_kernel(const int* __restrict__ ids, const int2* __restrict__ row_info, int rows, const float* __restrict__ embed, int vocab,
                  const float* __restrict__ pe, int D, float scale, float* __restrict__ x) {
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x;
  if (r >= rows) return;
  int id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const int t = row_info[r].x;
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    x[(size_t)r * D + c] = __fadd_rn(__fmul_rn(embed[(size_t)id * D + c], scale), pe[(size_t)t * D + c]);
}

// LayerNorm over D fp32 columns -> bf16 [rows, Dp] (pad columns zero).  One warp per row, two-pass statistics in registers.
__global__ void __launch_bounds__(128)
punc_ln_kernel(const float* __restrict__ x, int rows, int D, int Dp, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
               __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + (size_t)r * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / (float)D;
  float v = 0.f;
  for (int c = lane; c < D; c += 32) { const float d = xr[c] - mean; v += d * d; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const float rstd = rsqrtf(v / (float)D + eps);
  for (int c = lane; c < Dp; c += 32)
    out[(size_t)r * Dp + c] = __float2bfloat16(c < D ? (xr[c] - mean) * rstd * gamma[c] + beta[c] : 0.f);
}

// FSMN memory on V (columns [2D, 3D) of qkv): x[t, c] += v[t, c] + sum_k w[k][c] v[t + k - left, c], taps outside the sequence are zero;
// left = (K - 1) / 2 + sanm_shift (0 for the offline model: symmetric; 5 with K = 11 for the realtime model: causal).
// One thread per (row, 4 channels): float4 loads of V (the K neighbouring rows come from L1 / L2), float4 read-modify-write of x.
__global__ void __launch_bounds__(256)
punc_fsmn_kernel(const float* __restrict__ qkv, const int2* __restrict__ row_info, int rows, int D, int K, int left,
                 const float* __restrict__ w_t /*[K][D]*/, float* __restrict__ x) {
  pdl_wait();
  pdl_launch_dependents();
  const int d4 = D >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * d4) return;
  const int r = (int)(idx / d4), c = ((int)(idx - (long long)r * d4)) << 2;
  const int2 inf = row_info[r];    // {t, T}
  const size_t ld = (size_t)3 * D;
  const float* v = qkv + 2 * D + c;
  float4 acc = *reinterpret_cast<const float4*>(v + (size_t)r * ld);
  for (int k = 0; k < K; ++k) {
    const int tt = inf.x + k - left;
    if (tt < 0 || tt >= inf.y) continue;
    const float4 w = *reinterpret_cast<const float4*>(w_t + (size_t)k * D + c);
    const float4 u = *reinterpret_cast<const float4*>(v + (size_t)(r + k - left) * ld);
    acc.x += w.x * u.x; acc.y += w.y * u.y; acc.z += w.z * u.z; acc.w += w.w * u.w;
  }
  float4* xo = reinterpret_cast<float4*>(x + (size_t)r * D + c);
  float4 xv = *xo;
  xv.x += acc.x; xv.y += acc.y; xv.z += acc.z; xv.w += acc.w;
  *xo = xv;
}

// Softmax attention inside each sequence, head dimension dk <= 64.  Block = (tile of 16 queries of one sequence, head); the K / V
// rows of the sequence stream through shared memory 32 keys at a time; a warp owns one query at a time: lane j scores key j of the
// chunk, the running maximum / sum are warp-reduced (online softmax), and lane d accumulates output dimensions d and d + 32.
__global__ void __launch_bounds__(128)
punc_attn_kernel(const float* __restrict__ qkv, const int4* __restrict__ tiles, const int* __restrict__ tile_vad, int D, int dk, float scale,
                 __nv_bfloat16* __restrict__ ctx, int Dp) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float Ks[32][P_MAX_DK + 4];   // row stride 68 floats: float4 reads of 8 consecutive rows hit 32 distinct banks
  __shared__ __align__(16) float Vs[32][P_MAX_DK];
  __shared__ __align__(16) float Qs[P_QTILE][P_MAX_DK];
  const int4 tl = tiles[blockIdx.x];     // {first row of the tile, queries in the tile, first row of the sequence, sequence length}
  // realtime model (CTTransformerOnline::VadMask, ct-transformer-online.cpp:219-233): with 0 < vad_pos < T, queries before
  // vad_pos - 1 do not see keys from vad_pos on; everything else sees the whole sequence
  const int vad_pos = tile_vad[blockIdx.x];
  const int q_first = tl.x - tl.z;       // position of the tile's first query inside its sequence
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = (size_t)3 * D;
  const int col = h * dk;
  const bool vec = (dk & 3) == 0;           // rows of a head are 16-byte aligned (D % 4 == 0 always)
  const int dk4 = dk >> 2;
  if (vec) {
    for (int i = threadIdx.x; i < P_QTILE * dk4; i += blockDim.x) {
      const int q = i / dk4, d = (i - q * dk4) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < tl.y) v = *reinterpret_cast<const float4*>(qkv + (size_t)(tl.x + q) * ld + col + d);
      Qs[q][d] = v.x * scale; Qs[q][d + 1] = v.y * scale; Qs[q][d + 2] = v.z * scale; Qs[q][d + 3] = v.w * scale;
    }
  } else {
    for (int i = threadIdx.x; i < P_QTILE * dk; i += blockDim.x) {
      const int q = i / dk, d = i - q * dk;
      Qs[q][d] = q < tl.y ? qkv[(size_t)(tl.x + q) * ld + col + d] * scale : 0.f;
    }
  }
  float m[4], l[4], a0[4], a1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; a0[i] = 0.f; a1[i] = 0.f; }
  for (int k0 = 0; k0 < tl.w; k0 += 32) {
    __syncthreads();
    if (vec) {
      // 32 keys x dk floats of K and of V: all global loads of the chunk are issued before the first shared-memory store
      float4 kr[4], vr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * 128;
        kr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        vr[u] = kr[u];
        if (i < 32 * dk4) {
          const int j = i / dk4, d = (i - j * dk4) << 2;
          if (k0 + j < tl.w) {
            const float* row = qkv + (size_t)(tl.z + k0 + j) * ld + col + d;
            kr[u] = *reinterpret_cast<const float4*>(row + D);
            vr[u] = *reinterpret_cast<const float4*>(row + 2 * D);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * 128;
        if (i < 32 * dk4) {
          const int j = i / dk4, d = (i - j * dk4) << 2;
          *reinterpret_cast<float4*>(&Ks[j][d]) = kr[u];
          *reinterpret_cast<float4*>(&Vs[j][d]) = vr[u];
        }
      }
    } else {
      for (int i = threadIdx.x; i < 32 * dk; i += blockDim.x) {
        const int j = i / dk, d = i - j * dk;
        const bool ok = k0 + j < tl.w;
        const size_t row = (size_t)(tl.z + k0 + j) * ld;
        Ks[j][d] = ok ? qkv[row + D + col + d] : 0.f;
        Vs[j][d] = ok ? qkv[row + 2 * D + col + d] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = warp * 4 + i;
      if (q >= tl.y) break;          // warp-uniform
      const int k_end = (vad_pos > 0 && vad_pos < tl.w && q_first + q < vad_pos - 1) ? vad_pos : tl.w;
      if (k0 >= k_end) continue;     // warp-uniform: this chunk is entirely masked for the query
      const bool key_ok = k0 + lane < k_end;
      float s = 0.f;
      if (vec) {
        for (int d = 0; d < dk; d += 4) {
          const float4 a = *reinterpret_cast<const float4*>(&Qs[q][d]), b = *reinterpret_cast<const float4*>(&Ks[lane][d]);
          s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
      } else {
        for (int d = 0; d < dk; ++d) s += Qs[q][d] * Ks[lane][d];
      }
      s = key_ok ? s : -INFINITY;
      float mx = s;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m[i], mx);
      const float corr = expf(m[i] - m_new);      // exp(-inf) = 0 on the first chunk
      const float p = key_ok ? expf(s - m_new) : 0.f;
      float ps = p;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l[i] = l[i] * corr + ps;
      m[i] = m_new;
      float acc0 = a0[i] * corr, acc1 = a1[i] * corr;
      if (dk > 32) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const float pj = __shfl_sync(0xffffffffu, p, j);
          acc0 += pj * Vs[j][lane];
          acc1 += pj * Vs[j][lane + 32];
        }
      } else {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) acc0 += __shfl_sync(0xffffffffu, p, j) * Vs[j][lane];
      }
      a0[i] = acc0;
      a1[i] = acc1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = warp * 4 + i;
    if (q >= tl.y) break;
    const float inv = 1.0f / l[i];
    __nv_bfloat16* o = ctx + (size_t)(tl.x + q) * Dp + col;
    if (lane < dk) o[lane] = __float2bfloat16(a0[i] * inv);
    if (lane + 32 < dk) o[lane + 32] = __float2bfloat16(a1[i] * inv);
  }
}

// ---- tensor-core attention (the product path) ----------------------------------------------------------------------------
// mma.sync.m16n8k8 TF32 (fp32 operands, 10-bit mantissa, fp32 accumulate).  Block = (64-query tile of one sequence, head), one
// warp per 16 queries; K / V of the sequence stream through shared memory 64 keys at a time (row stride 68 floats: every fragment
// load below touches 32 distinct banks); Q fragments live in registers.  S = Q K^T for the chunk sits in the accumulator layout
// (row g / g + 8, keys 2t, 2t + 1 of each 8-key block); the same registers are the A operand of P V once the key order of the V
// fragment follows that layout (k index t -> key 2t, t + 4 -> key 2t + 1), so P never leaves the registers.  Online softmax
// across chunks; the realtime model's VadMask is a per-row key limit.
constexpr int PT_Q = 64, PT_K = 64, PT_LD = 68;

__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NB>   // NB = head dim rounded up to a multiple of 8, divided by 8 (4, 6 or 8)
__global__ void __launch_bounds__(128)
punc_attn_mma_kernel(const float* __restrict__ qkv, const int4* __restrict__ tiles, const int* __restrict__ tile_vad, int D, int dk, float scale,
                     __nv_bfloat16* __restrict__ ctx, int Dp) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float Ks[PT_K][PT_LD];
  __shared__ __align__(16) float Vs[PT_K][PT_LD];
  const int4 tl = tiles[blockIdx.x];     // {first row of the tile, queries in the tile, first row of the sequence, sequence length}
  const int vad_pos = tile_vad[blockIdx.x];
  const int h = blockIdx.y, col = h * dk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t ld = (size_t)3 * D;
  const int q_first = tl.x - tl.z;                 // position of the tile's first query inside its sequence
  const int q0 = warp * 16 + g, q1 = q0 + 8;       // this thread's two query rows inside the tile
  const bool warp_on = warp * 16 < tl.y;
  // Q fragments (scaled): a[db] = {Q[q0][8db + t], Q[q1][8db + t], Q[q0][8db + t + 4], Q[q1][8db + t + 4]}
  uint32_t qa[NB][4];
#pragma unroll
  for (int db = 0; db < NB; ++db) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = (e & 1) ? q1 : q0, d = db * 8 + t + ((e & 2) ? 4 : 0);
      const float v = (q < tl.y && d < dk) ? qkv[(size_t)(tl.x + q) * ld + col + d] * scale : 0.f;
      qa[db][e] = __float_as_uint(v);
    }
  }
  auto k_end_of = [&](int q) { return (vad_pos > 0 && vad_pos < tl.w && q_first + q < vad_pos - 1) ? vad_pos : tl.w; };
  const int kend0 = k_end_of(q0), kend1 = k_end_of(q1);
  float o[NB][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) { o[nb][0] = 0.f; o[nb][1] = 0.f; o[nb][2] = 0.f; o[nb][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const bool vec = (dk & 3) == 0;
  for (int k0 = 0; k0 < tl.w; k0 += PT_K) {
    __syncthreads();
    if (vec) {
      const int dk4 = dk >> 2, n4 = PT_K * dk4;
      for (int i0 = 0; i0 < n4; i0 += 128 * 4) {
        float4 kr[4], vr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + threadIdx.x + u * 128;
          kr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          vr[u] = kr[u];
          if (i < n4) {
            const int j = i / dk4, d = (i - j * dk4) << 2;
            if (k0 + j < tl.w) {
              const float* row = qkv + (size_t)(tl.z + k0 + j) * ld + col + d;
              kr[u] = *reinterpret_cast<const float4*>(row + D);
              vr[u] = *reinterpret_cast<const float4*>(row + 2 * D);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + threadIdx.x + u * 128;
          if (i < n4) {
            const int j = i / dk4, d = (i - j * dk4) << 2;
            *reinterpret_cast<float4*>(&Ks[j][d]) = kr[u];
            *reinterpret_cast<float4*>(&Vs[j][d]) = vr[u];
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < PT_K * dk; i += 128) {
        const int j = i / dk, d = i - j * dk;
        const bool ok = k0 + j < tl.w;
        const size_t row = (size_t)(tl.z + k0 + j) * ld;
        Ks[j][d] = ok ? qkv[row + D + col + d] : 0.f;
        Vs[j][d] = ok ? qkv[row + 2 * D + col + d] : 0.f;
      }
    }
    if (dk < NB * 8) {   // zero the padding columns the fragments read
      const int pad = NB * 8 - dk;
      for (int i = threadIdx.x; i < PT_K * pad; i += 128) {
        const int j = i / pad, d = dk + (i - j * pad);
        Ks[j][d] = 0.f;
        Vs[j][d] = 0.f;
      }
    }
    __syncthreads();
    if (!warp_on) continue;
    // S = Q K^T for 64 keys: 8 key blocks x NB dim blocks
    float sc[8][4];
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      sc[kb][0] = 0.f; sc[kb][1] = 0.f; sc[kb][2] = 0.f; sc[kb][3] = 0.f;
#pragma unroll
      for (int db = 0; db < NB; ++db) {
        const uint32_t b0 = __float_as_uint(Ks[kb * 8 + g][db * 8 + t]), b1 = __float_as_uint(Ks[kb * 8 + g][db * 8 + t + 4]);
        mma_tf32_1688(sc[kb], qa[db], b0, b1);
      }
    }
    // mask, running maximum, exponentials (accumulator layout: [0],[1] row q0 keys 2t, 2t+1; [2],[3] row q1)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      const int key = k0 + kb * 8 + 2 * t;
      if (key >= kend0) sc[kb][0] = -INFINITY;
      if (key + 1 >= kend0) sc[kb][1] = -INFINITY;
      if (key >= kend1) sc[kb][2] = -INFINITY;
      if (key + 1 >= kend1) sc[kb][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(sc[kb][0], sc[kb][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[kb][2], sc[kb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    // a row whose every key so far is masked keeps maximum -inf: use 0 as the reference so that exp(-inf - 0) = 0
    const float r0 = mn0 == -INFINITY ? 0.f : mn0, r1 = mn1 == -INFINITY ? 0.f : mn1;
    const float c0 = expf(m0 - r0), c1 = expf(m1 - r1);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      sc[kb][0] = expf(sc[kb][0] - r0); sc[kb][1] = expf(sc[kb][1] - r0);
      sc[kb][2] = expf(sc[kb][2] - r1); sc[kb][3] = expf(sc[kb][3] - r1);
      s0 += sc[kb][0] + sc[kb][1];
      s1 += sc[kb][2] + sc[kb][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    l0 = l0 * c0 + s0; l1 = l1 * c1 + s1;
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) { o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1; }
    // O += P V: A = P in accumulator order (k index t <-> key 2t, t + 4 <-> key 2t + 1), B rows follow the same key order
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      const uint32_t pa[4] = {__float_as_uint(sc[kb][0]), __float_as_uint(sc[kb][2]), __float_as_uint(sc[kb][1]), __float_as_uint(sc[kb][3])};
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const uint32_t b0 = __float_as_uint(Vs[kb * 8 + 2 * t][nb * 8 + g]), b1 = __float_as_uint(Vs[kb * 8 + 2 * t + 1][nb * 8 + g]);
        mma_tf32_1688(o[nb], pa, b0, b1);
      }
    }
  }
  if (!warp_on) return;
  // accumulator layout of O: [0],[1] -> row q0, dims 8nb + 2t, 8nb + 2t + 1; [2],[3] -> row q1
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int d = nb * 8 + 2 * t;
    if (q0 < tl.y) {
      __nv_bfloat16* dst = ctx + (size_t)(tl.x + q0) * Dp + col;
      if (d < dk) dst[d] = __float2bfloat16(o[nb][0] * i0);
      if (d + 1 < dk) dst[d + 1] = __float2bfloat16(o[nb][1] * i0);
    }
    if (q1 < tl.y) {
      __nv_bfloat16* dst = ctx + (size_t)(tl.x + q1) * Dp + col;
      if (d < dk) dst[d] = __float2bfloat16(o[nb][2] * i1);
      if (d + 1 < dk) dst[d + 1] = __float2bfloat16(o[nb][3] * i1);
    }
  }
}

// first maximum over classes [0, n_cand) of logits [rows, ld]
__global__ void __launch_bounds__(256)
punc_argmax_kernel(const float* __restrict__ logits, int rows, int ld, int n_cand, int* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* x = logits + (size_t)r * ld;
  int best = 0;
  for (int j = 1; j < n_cand; ++j)
    if (x[j] > x[best]) best = j;
  out[r] = best;
}

uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

struct PLayer {
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr, *fsmn_w = nullptr;
  PLinear qkv, out, w1, w2;
};

}  // namespace

struct b200pf_punc {
  int device = 0, num_sms = 148, max_tokens = 0;
  int vocab = 0, D = 0, Dp = 0, H = 0, dk = 0, F = 0, Fp = 0, L = 0, K = 0, n_punc = 0, n_out_p = 8, shift = 0;
  float eps = 1e-12f;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::vector<void*> allocs;
  float *embed = nullptr, *pe = nullptr, *an_g = nullptr, *an_b = nullptr;
  std::vector<PLayer> layers;
  PLinear dec;
  // workspace
  int* d_ids = nullptr;
  int2* d_row_info = nullptr;
  int4* d_tiles = nullptr;
  float *x = nullptr, *qkv = nullptr, *logits = nullptr;
  __nv_bfloat16 *h = nullptr, *ctx = nullptr, *f1 = nullptr;
  int* d_punc = nullptr;
  // pinned staging: one upload (ids | row info | tiles) and one download per call, no pageable-memory bounce
  uint8_t* h_in = nullptr;
  uint8_t* d_in = nullptr;
  int* h_punc = nullptr;
  size_t in_bytes = 0;
  int64_t launches = 0;
};

#define PCK(call, what) do { int rc_ = check_cuda((call), what); if (rc_) return rc_; } while (0)

extern "C" {

static int b200pf_punc_create_impl(const char* punc_dir, int device, int max_tokens, b200pf_punc** out) {
  if (!punc_dir || !out) { set_error("null argument"); return B200PF_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device: the B200 path has no CPU fallback"); return B200PF_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) { set_error("bad device index"); return B200PF_ERR_INVALID; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) { set_error("device is not sm_100 (Blackwell B200)"); return B200PF_ERR_NO_DEVICE; }
  PCK(cudaSetDevice(device), "cudaSetDevice");
  std::string err;
  WeightFile wf;
  if (!read_weight_file(std::string(punc_dir) + "/punc.b200pf", &wf, &err)) { set_error(err); return B200PF_ERR_IO; }
  auto cfgi = [&](const char* k, int dflt) { auto it = wf.cfg.find(k); return it == wf.cfg.end() ? dflt : (int)it->second; };
  std::unique_ptr<b200pf_punc> p(new b200pf_punc);
  p->device = device;
  p->vocab = cfgi("vocab", 0); p->D = cfgi("d_model", 0); p->H = cfgi("n_heads", 0); p->F = cfgi("d_ff", 0);
  p->L = cfgi("n_layers", 0); p->K = cfgi("kernel", 11); p->n_punc = cfgi("n_punc", 6); p->shift = cfgi("sanm_shift", 0);
  { auto it = wf.cfg.find("ln_eps"); if (it != wf.cfg.end()) p->eps = (float)it->second; }
  if (p->vocab <= 0 || p->D <= 0 || p->H <= 0 || p->F <= 0 || p->L <= 0 || p->D % p->H || (p->D & 3) || (p->F & 7) || p->D / p->H > P_MAX_DK ||
      p->n_punc < 2 || p->n_punc > 8 || p->K < 1 || p->K > 31 || !(p->K & 1) || p->shift < 0 || (p->K - 1) / 2 + p->shift > p->K - 1) {
    set_error("punc.b200pf: unsupported configuration (need d_model % n_heads == 0, d_model % 4 == 0, d_ff % 8 == 0, head dim <= 64, n_punc <= 8)");
    return B200PF_ERR_IO;
  }
  p->dk = p->D / p->H;
  p->Dp = (p->D + 7) & ~7;
  p->Fp = p->F;
  p->max_tokens = max_tokens > 0 ? max_tokens : 65536;
  cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device);
  PCK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  bool ok = true;
  auto fail_cleanup = [&]() { for (void* a : p->allocs) cudaFree(a); cudaFreeHost(p->h_in); cudaFreeHost(p->h_punc); cudaStreamDestroy(p->stream); };
  auto dalloc = [&](size_t bytes) -> void* {
    void* d = nullptr;
    if (!ok) return nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 256) != cudaSuccess) { ok = false; err = "cudaMalloc failed (punctuation model)"; cudaGetLastError(); return nullptr; }
    p->allocs.push_back(d);
    return d;
  };
  auto up = [&](const void* h, size_t bytes) -> void* {
    void* d = dalloc(bytes);
    if (d && cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { ok = false; err = "cudaMemcpy failed (punctuation model)"; }
    return d;
  };
  auto vec = [&](const std::string& name, int n) -> float* {
    auto it = wf.tensors.find(name);
    if (it == wf.tensors.end() || (int)it->second.numel() != n) { if (ok) err = "missing or misshaped tensor " + name; ok = false; return nullptr; }
    return (float*)up(it->second.data.data(), (size_t)n * 4);
  };
  auto linear = [&](const std::string& name, int out_d, int in_d, int out_p, int in_p) {
    PLinear Lw;
    Lw.out = out_p; Lw.in = in_p;
    auto it = wf.tensors.find(name + ".weight");
    auto ib = wf.tensors.find(name + ".bias");
    if (it == wf.tensors.end() || it->second.numel() != (int64_t)out_d * in_d || ib == wf.tensors.end() || (int)ib->second.numel() != out_d) {
      if (ok) err = "missing or misshaped tensor " + name;
      ok = false;
      return Lw;
    }
    std::vector<uint16_t> w((size_t)out_p * in_p, 0);
    for (int o = 0; o < out_d; ++o)
      for (int i = 0; i < in_d; ++i) w[(size_t)o * in_p + i] = bf16_rne(it->second.data[(size_t)o * in_d + i]);
    Lw.w = (__nv_bfloat16*)up(w.data(), w.size() * 2);
    std::vector<float> b(out_p, 0.f);
    memcpy(b.data(), ib->second.data.data(), (size_t)out_d * 4);
    Lw.b = (float*)up(b.data(), b.size() * 4);
    return Lw;
  };
  const int D = p->D, Dp = p->Dp, F = p->F;
  {
    auto it = wf.tensors.find("embed.weight");
    if (it == wf.tensors.end() || it->second.numel() != (int64_t)p->vocab * D) { ok = false; err = "missing or misshaped tensor embed.weight"; }
    else p->embed = (float*)up(it->second.data.data(), (size_t)p->vocab * D * 4);
  }
  p->layers.resize(p->L);
  for (int l = 0; l < p->L && ok; ++l) {
    const std::string pre = l == 0 ? std::string("encoder.encoders0.0") : "encoder.encoders." + std::to_string(l - 1);
    PLayer& Ly = p->layers[l];
    Ly.ln1_g = vec(pre + ".norm1.weight", D); Ly.ln1_b = vec(pre + ".norm1.bias", D);
    Ly.ln2_g = vec(pre + ".norm2.weight", D); Ly.ln2_b = vec(pre + ".norm2.bias", D);
    Ly.qkv = linear(pre + ".self_attn.linear_q_k_v", 3 * D, D, 3 * D, Dp);
    Ly.out = linear(pre + ".self_attn.linear_out", D, D, D, Dp);
    Ly.w1 = linear(pre + ".feed_forward.w_1", F, D, F, Dp);
    Ly.w2 = linear(pre + ".feed_forward.w_2", D, F, D, p->Fp);
    auto it = wf.tensors.find(pre + ".self_attn.fsmn_block.weight");
    if (it == wf.tensors.end() || it->second.numel() != (int64_t)D * p->K) { if (ok) err = "missing tensor " + pre + ".self_attn.fsmn_block.weight"; ok = false; break; }
    std::vector<float> wt((size_t)p->K * D);
    for (int c = 0; c < D; ++c)
      for (int k = 0; k < p->K; ++k) wt[(size_t)k * D + c] = it->second.data[(size_t)c * p->K + k];
    Ly.fsmn_w = (float*)up(wt.data(), wt.size() * 4);
  }
  p->an_g = vec("encoder.after_norm.weight", D);
  p->an_b = vec("encoder.after_norm.bias", D);
  p->dec = linear("decoder", p->n_punc, D, p->n_out_p, Dp);
  {
    // sinusoidal position table, positions from 1 (SinusoidalPositionEncoder; same arithmetic as the acoustic model's table)
    const int half = D / 2;
    std::vector<float> pe((size_t)P_MAX_POS * D, 0.f);
    const float inc = (float)(-(log(10000.0) / (half - 1)));
    for (int t = 0; t < P_MAX_POS; ++t)
      for (int i = 0; i < half; ++i) {
        const float inv = expf((float)i * inc);
        const float st = (float)(t + 1) * inv;
        pe[(size_t)t * D + i] = sinf(st);
        pe[(size_t)t * D + half + i] = cosf(st);
      }
    p->pe = (float*)up(pe.data(), pe.size() * 4);
  }
  const size_t R = (size_t)p->max_tokens;
  const size_t max_tiles = R / P_QTILE + P_MAX_SEQ + 1;
  p->in_bytes = R * 4 + R * 8 + max_tiles * 20 + 1024;
  p->d_in = (uint8_t*)dalloc(p->in_bytes);
  if (ok && (cudaMallocHost((void**)&p->h_in, p->in_bytes) != cudaSuccess || cudaMallocHost((void**)&p->h_punc, R * 4) != cudaSuccess)) {
    ok = false; err = "cudaMallocHost failed (punctuation staging)"; cudaGetLastError();
  }
  p->x = (float*)dalloc(R * D * 4); p->qkv = (float*)dalloc(R * 3 * D * 4); p->logits = (float*)dalloc(R * p->n_out_p * 4);
  p->h = (__nv_bfloat16*)dalloc(R * Dp * 2); p->ctx = (__nv_bfloat16*)dalloc(R * Dp * 2); p->f1 = (__nv_bfloat16*)dalloc(R * p->Fp * 2);
  p->d_punc = (int*)dalloc(R * 4);
  if (ok && (cudaMemset(p->h, 0, R * Dp * 2) != cudaSuccess || cudaMemset(p->ctx, 0, R * Dp * 2) != cudaSuccess)) { ok = false; err = "cudaMemset failed"; }
  if (!ok) { set_error(err.empty() ? "punctuation model upload failed" : err); fail_cleanup(); return B200PF_ERR_IO; }
  PCK(cudaDeviceSynchronize(), "punc init");
  *out = p.release();
  return 0;
}
// Parsing a hostile or truncated model directory may throw (std::bad_alloc, std::invalid_argument from the text parsers);
// nothing may unwind through the C ABI: it becomes an error code with the text in b200pf_last_error().
int b200pf_punc_create(const char* punc_dir, int device, int max_tokens, b200pf_punc** out) {
  try {
    return b200pf_punc_create_impl(punc_dir, device, max_tokens, out);
  } catch (const std::exception& ex) {
    set_error(std::string("b200pf_punc_create: ") + ex.what());
    return B200PF_ERR_IO;
  } catch (...) {
    set_error("b200pf_punc_create: unknown exception");
    return B200PF_ERR_IO;
  }
}

void b200pf_punc_destroy(b200pf_punc* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->stream);
  for (void* a : p->allocs) cudaFree(a);
  cudaFreeHost(p->h_in);
  cudaFreeHost(p->h_punc);
  cudaStreamDestroy(p->stream);
  delete p;
}

int b200pf_punc_info(const b200pf_punc* p, int* vocab, int* n_punc, int* d_model, int* max_tokens) {
  if (!p) { set_error("null argument"); return B200PF_ERR_INVALID; }
  if (vocab) *vocab = p->vocab;
  if (n_punc) *n_punc = p->n_punc;
  if (d_model) *d_model = p->D;
  if (max_tokens) *max_tokens = p->max_tokens;
  return 0;
}

long long b200pf_punc_launches(const b200pf_punc* p) { return p ? p->launches : 0; }

int b200pf_punc_infer(b200pf_punc* p, const int32_t* ids, const int32_t* offsets, int n_seq, int32_t* punc_out, float* logits_out) {
  return b200pf_punc_infer_vad(p, ids, offsets, nullptr, n_seq, punc_out, logits_out);
}

int b200pf_punc_infer_vad(b200pf_punc* p, const int32_t* ids, const int32_t* offsets, const int32_t* vad_pos, int n_seq, int32_t* punc_out,
                          float* logits_out) {
  if (!p || !offsets || n_seq < 0 || (n_seq > 0 && !ids) || !punc_out) { set_error("bad argument"); return B200PF_ERR_INVALID; }
  if (n_seq > P_MAX_SEQ) { set_error("too many sequences in one call"); return B200PF_ERR_CAPACITY; }
  if (n_seq == 0) return 0;
  const int base = offsets[0], rows = offsets[n_seq] - base;
  if (rows < 0) { set_error("offsets not monotone"); return B200PF_ERR_INVALID; }
  if (rows > p->max_tokens) { set_error("tokens exceed the punctuation engine's capacity"); return B200PF_ERR_CAPACITY; }
  if (rows == 0) return 0;
  PCK(cudaSetDevice(p->device), "cudaSetDevice");
  std::lock_guard<std::mutex> lock(p->mu);
  // staging layout: ids [rows] | row info [rows] | tiles [n_tiles] | per-tile vad_pos [n_tiles], 16-byte aligned sections
  const size_t off_info = ((size_t)rows * 4 + 15) & ~size_t(15), off_tiles = (off_info + (size_t)rows * 8 + 15) & ~size_t(15);
  static const bool scalar_attn = getenv("B200PF_PUNC_SCALAR_ATTN") != nullptr;   // the CUDA-core kernel, kept for comparison
  const int qtile = scalar_attn ? P_QTILE : PT_Q;
  size_t want_tiles = 0;
  for (int i = 0; i < n_seq; ++i) want_tiles += (size_t)((offsets[i + 1] - offsets[i] + qtile - 1) / qtile);
  const size_t off_vad = off_tiles + want_tiles * 16;
  if (off_vad + want_tiles * 4 > p->in_bytes) { set_error("too many attention tiles"); return B200PF_ERR_CAPACITY; }
  int* h_ids = (int*)p->h_in;
  int2* info = (int2*)(p->h_in + off_info);
  int4* tiles = (int4*)(p->h_in + off_tiles);
  int* tile_vad = (int*)(p->h_in + off_vad);
  const size_t max_tiles = want_tiles;
  size_t n_tiles = 0;
  memcpy(h_ids, ids + base, (size_t)rows * 4);
  for (int i = 0; i < n_seq; ++i) {
    const int T = offsets[i + 1] - offsets[i], r0 = offsets[i] - base;
    if (T < 0) { set_error("offsets not monotone"); return B200PF_ERR_INVALID; }
    if (T > P_MAX_POS) { set_error("a sequence is longer than the position table (4096 tokens)"); return B200PF_ERR_CAPACITY; }
    for (int t = 0; t < T; ++t) info[r0 + t] = make_int2(t, T);
    for (int q0 = 0; q0 < T; q0 += qtile) {
      if (n_tiles >= max_tiles) { set_error("too many attention tiles"); return B200PF_ERR_CAPACITY; }
      tile_vad[n_tiles] = vad_pos ? vad_pos[i] : 0;
      tiles[n_tiles++] = make_int4(r0 + q0, T - q0 < qtile ? T - q0 : qtile, r0, T);
    }
  }
  cudaStream_t s = p->stream;
  const int D = p->D, Dp = p->Dp;
  p->d_ids = (int*)p->d_in;
  p->d_row_info = (int2*)(p->d_in + off_info);
  p->d_tiles = (int4*)(p->d_in + off_tiles);
  const int* d_tile_vad = (const int*)(p->d_in + off_vad);
  PCK(cudaMemcpyAsync(p->d_in, p->h_in, off_vad + n_tiles * 4, cudaMemcpyHostToDevice, s), "H2D punctuation input");
  int rc = launch_kernel(punc_embed_kernel, dim3(rows), dim3(128), 0, s, (const int*)p->d_ids, (const int2*)p->d_row_info, rows, (const float*)p->embed,
                         p->vocab, (const float*)p->pe, D, sqrtf((float)D), p->x);
  if (rc) return check_cuda((cudaError_t)rc, "punc embed");
  int n_launch = 1;
  auto ln = [&](const float* g, const float* b) {
    ++n_launch;
    return launch_kernel(punc_ln_kernel, dim3((rows + 3) / 4), dim3(128), 0, s, (const float*)p->x, rows, D, Dp, g, b, p->eps, p->h);
  };
  auto gemm = [&](const __nv_bfloat16* A, int lda, const PLinear& W, int n_real, int relu, __nv_bfloat16* ob, int ldo, float* of, int ldof, const float* res) {
    GemmProblem gp;
    gp.A = A; gp.lda = lda; gp.rows_a = rows; gp.W = W.w; gp.ldw = W.in; gp.M = rows; gp.N = n_real; gp.K = W.in;
    GemmEpilogue e;
    e.bias = W.b; e.relu = relu; e.out_bf16 = ob; e.ld_out_bf16 = ldo; e.out_f32 = of; e.ld_out_f32 = ldof; e.res_f32 = res; e.ld_res = ldof;
    ++n_launch;
    return gemm_bf16_tcgen05(gp, e, p->num_sms, s);
  };
  const float scale = 1.0f / sqrtf((float)p->dk);
  for (int l = 0; l < p->L; ++l) {
    const PLayer& Ly = p->layers[l];
    if ((rc = ln(Ly.ln1_g, Ly.ln1_b))) return check_cuda((cudaError_t)rc, "punc ln1");
    if ((rc = gemm(p->h, Dp, Ly.qkv, 3 * D, 0, nullptr, 0, p->qkv, 3 * D, nullptr))) return check_cuda((cudaError_t)rc, "punc gemm qkv");
    {
      const dim3 grid((unsigned)n_tiles, p->H);
      const float* q = p->qkv;
      const int4* tl = p->d_tiles;
      if (scalar_attn) rc = launch_kernel(punc_attn_kernel, grid, dim3(128), 0, s, q, tl, d_tile_vad, D, p->dk, scale, p->ctx, Dp);
      else if (p->dk <= 32) rc = launch_kernel(punc_attn_mma_kernel<4>, grid, dim3(128), 0, s, q, tl, d_tile_vad, D, p->dk, scale, p->ctx, Dp);
      else if (p->dk <= 48) rc = launch_kernel(punc_attn_mma_kernel<6>, grid, dim3(128), 0, s, q, tl, d_tile_vad, D, p->dk, scale, p->ctx, Dp);
      else rc = launch_kernel(punc_attn_mma_kernel<8>, grid, dim3(128), 0, s, q, tl, d_tile_vad, D, p->dk, scale, p->ctx, Dp);
    }
    if (rc) return check_cuda((cudaError_t)rc, "punc attention");
    // x += v + fsmn(v): after the attention kernel has been queued (it does not read x), before the out-projection accumulates into x
    rc = launch_kernel(punc_fsmn_kernel, dim3((unsigned)(((long long)rows * (D >> 2) + 255) / 256)), dim3(256), 0, s, (const float*)p->qkv, (const int2*)p->d_row_info, rows, D, p->K,
                       (p->K - 1) / 2 + p->shift, (const float*)Ly.fsmn_w, p->x);
    if (rc) return check_cuda((cudaError_t)rc, "punc fsmn");
    n_launch += 2;
    if ((rc = gemm(p->ctx, Dp, Ly.out, D, 0, nullptr, 0, p->x, D, p->x))) return check_cuda((cudaError_t)rc, "punc gemm out");
    if ((rc = ln(Ly.ln2_g, Ly.ln2_b))) return check_cuda((cudaError_t)rc, "punc ln2");
    if ((rc = gemm(p->h, Dp, Ly.w1, p->F, 1, p->f1, p->Fp, nullptr, 0, nullptr))) return check_cuda((cudaError_t)rc, "punc gemm ffn1");
    if ((rc = gemm(p->f1, p->Fp, Ly.w2, D, 0, nullptr, 0, p->x, D, p->x))) return check_cuda((cudaError_t)rc, "punc gemm ffn2");
  }
  if ((rc = ln(p->an_g, p->an_b))) return check_cuda((cudaError_t)rc, "punc after_norm");
  if ((rc = gemm(p->h, Dp, p->dec, p->n_out_p, 0, nullptr, 0, p->logits, p->n_out_p, nullptr))) return check_cuda((cudaError_t)rc, "punc gemm classifier");
  rc = launch_kernel(punc_argmax_kernel, dim3((rows + 255) / 256), dim3(256), 0, s, (const float*)p->logits, rows, p->n_out_p, p->n_punc - 1, p->d_punc);
  if (rc) return check_cuda((cudaError_t)rc, "punc argmax");
  ++n_launch;
  PCK(cudaMemcpyAsync(p->h_punc, p->d_punc, (size_t)rows * 4, cudaMemcpyDeviceToHost, s), "D2H punc");
  std::vector<float> lg;
  if (logits_out) { lg.resize((size_t)rows * p->n_out_p); PCK(cudaMemcpyAsync(lg.data(), p->logits, lg.size() * 4, cudaMemcpyDeviceToHost, s), "D2H logits"); }
  PCK(cudaStreamSynchronize(s), "punc forward");
  memcpy(punc_out + base, p->h_punc, (size_t)rows * 4);
  if (logits_out)
    for (int r = 0; r < rows; ++r) memcpy(logits_out + (size_t)(base + r) * p->n_punc, lg.data() + (size_t)r * p->n_out_p, (size_t)p->n_punc * 4);
  p->launches += n_launch;
  return 0;
}

}  // extern "C"
