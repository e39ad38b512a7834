// EXPERIMENT, not compiled into libb200pf.so (round 1 measured it slower than LayerNorm + GEMM as two kernels, DESIGN.md 3a').
// It needs `gemm_ln_bf16_tcgen05` declared in gemm.cuh and a b200pf_op_gemm_ln entry point to be built again.
// EXPERIMENT - correct (tests/test_gpu_parity.py::test_gemm_with_fused_layernorm) but NOT on the product path: measured
// 234 us (QKV, 61 k rows) / 262 us (FFN1) against 84 + 27 us / 108 + 27 us for the LayerNorm kernel followed by the plain
// GEMM (tools/bench_gemm_ln.py).  Two warps per CTA cannot keep enough loads in flight to pull a 256 KB fp32 tile in the
// few microseconds an m-tile switch may cost (16 batches x one memory latency each ~ 35 us per switch), and with the
// 128 KB operand tile resident there is no shared memory left to stage x asynchronously.  The stand-alone LayerNorm kernel
// (one warp per row over the whole GPU, ~7 TB/s incl. L2 hits) stays.
//
// LayerNorm fused into the GEMM that consumes it (encoder: LN1 -> QKV projection, LN2 -> FFN1):
//     out[M,N] = bf16( act( LN(x)[M,512] * W[N,512]^T + bias ) ),   x fp32 (the residual stream), LN eps / gamma / beta.
// Replaces a LayerNorm launch (read 2 KB + write 1 KB per row) and the A-operand traffic of the plain GEMM (1 KB per row
// and N-tile through TMA) by ONE read of x per (row, m-tile visit): two spare warps of each CTA normalise the CTA's 128
// rows straight into the tensor core's operand layout (eight [128 x 64] SWIZZLE_128B sub-tiles, 128 KB, resident for all
// N-tiles of the m-tile).  W streams through a 3-stage TMA ring as in gemm.cu; the MMA (tcgen05.mma.cta_group::2,
// 256x256x16 per CTA pair), the TMEM double buffering and the bf16 TMA-store epilogue are the ones of gemm.cu (EPI 1).
//
// Scheduling: (m-tile, n-tile) units in m-major order are cut into contiguous ranges, one per CTA pair, so that a pair
// normalises an m-tile once for all the N-tiles it owns (ranges differ by at most one unit; an m-tile that straddles two
// ranges is normalised by both pairs).
//
//   warp 0      : TMA producer for W            warp 1 : MMA issuer (even CTA)
//   warps 2, 3  : LayerNorm -> A sub-tiles (64 rows each; warp 2 also allocates TMEM)
//   warps 4..11 : epilogue (bias, ReLU, bf16, TMA store)
#include <math.h>

#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace pf {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, KD = 512, KB = KD / BK, SB = 3;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int A_SUB = BM * BK * 2;           // 16 KB
constexpr int A_BYTES = KB * A_SUB;          // 128 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;   // 16 KB
constexpr int STG_BYTES = 4096, BIAS_BYTES = 512;
constexpr int kTmemCols = 512;
constexpr int kSmemBytes = A_BYTES + SB * B_BYTES + kEpiWarps * (STG_BYTES + BIAS_BYTES) + 256 /*barriers*/ + 1024 /*align*/;

struct LnArgs {
  const float* x;
  int ldx;
  const float* gamma;
  const float* beta;
  float eps;
  int M, N;
  const float* bias;
  int relu;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, LnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  uint8_t* sStg = sB + SB * B_BYTES;
  uint8_t* sBias = sStg + kEpiWarps * STG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kEpiWarps * BIAS_BYTES);
  uint64_t* b_full = bars;              // [SB]  even CTA
  uint64_t* b_empty = bars + SB;        // [SB]  each CTA (multicast commit)
  uint64_t* a_full = bars + 2 * SB;     // even CTA: 2 LayerNorm warps x 2 CTAs
  uint64_t* a_empty = a_full + 1;       // each CTA (multicast commit)
  uint64_t* tfull = a_empty + 1;        // [2] each CTA
  uint64_t* tempty = tfull + 2;         // [2] even CTA: 2 x kEpiWarps arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles = a.N / BN;
  const int m_tiles = (a.M + 2 * BM - 1) / (2 * BM);
  const long long U = (long long)m_tiles * n_tiles;
  const int u0 = (int)(U * pair / n_pairs), u1 = (int)(U * (pair + 1) / n_pairs);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(a_full, 4);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u0; u < u1; ++u) {
        const int n0 = (u % n_tiles) * BN + (int)cta * (BN / 2);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&b_full[stage], 2 * B_BYTES);
          tma_load_2d_2cta(sB + stage * B_BYTES, &tmB, &b_full[stage], kb * BK, n0);
          if (++stage == SB) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0, gen = 0;
      for (int u = u0; u < u1; ++u) {
        const int n = u % n_tiles;
        const bool first = (u == u0) || (n == 0);
        const bool last = (u == u1 - 1) || (n == n_tiles - 1);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        if (first) mbar_wait(a_full, gen & 1);   // both CTAs' 128 normalised rows are in shared memory
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(sA + kb * A_SUB));
          const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta_mc(&b_empty[stage], 0x3);
          if (++stage == SB) { stage = 0; phase ^= 1; }
        }
        if (last) { umma_commit_2cta_mc(a_empty, 0x3); ++gen; }   // the A tiles may be refilled once these MMAs are done
        umma_commit_2cta_mc(&tfull[acc], 0x3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ---- LayerNorm producer: this warp's 64 rows of every m-tile the pair touches ----
    const int lw = warp - 2;
    // lane owns columns lane*8 + 256*i + e  (i = 0,1; e = 0..7): sub-tile 4i + lane/8, 16-byte chunk lane%8
    float g[16], bt[16];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 g0 = *reinterpret_cast<const float4*>(a.gamma + lane * 8 + 256 * i), g1 = *reinterpret_cast<const float4*>(a.gamma + lane * 8 + 256 * i + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(a.beta + lane * 8 + 256 * i), b1 = *reinterpret_cast<const float4*>(a.beta + lane * 8 + 256 * i + 4);
      g[8 * i] = g0.x; g[8 * i + 1] = g0.y; g[8 * i + 2] = g0.z; g[8 * i + 3] = g0.w; g[8 * i + 4] = g1.x; g[8 * i + 5] = g1.y; g[8 * i + 6] = g1.z; g[8 * i + 7] = g1.w;
      bt[8 * i] = b0.x; bt[8 * i + 1] = b0.y; bt[8 * i + 2] = b0.z; bt[8 * i + 3] = b0.w; bt[8 * i + 4] = b1.x; bt[8 * i + 5] = b1.y; bt[8 * i + 6] = b1.z; bt[8 * i + 7] = b1.w;
    }
    const uint32_t sA_u = smem_u32(sA);
    uint32_t gen = 0;
    if (u1 > u0) {
      const int m_first = u0 / n_tiles, m_last = (u1 - 1) / n_tiles;
      for (int m = m_first; m <= m_last; ++m, ++gen) {
        if (gen > 0) mbar_wait(a_empty, (gen - 1) & 1);
        const int row0 = m * 2 * BM + (int)cta * BM + lw * 64;
#pragma unroll 1
        for (int r4 = 0; r4 < 16; ++r4) {
          float v[4][16];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int row = row0 + r4 * 4 + rr;
            const bool ok = row < a.M;
            const float* xp = a.x + (size_t)(ok ? row : 0) * a.ldx + lane * 8;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float4 f0 = ok ? ldg128f(xp + 256 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 f1 = ok ? ldg128f(xp + 256 * i + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[rr][8 * i] = f0.x; v[rr][8 * i + 1] = f0.y; v[rr][8 * i + 2] = f0.z; v[rr][8 * i + 3] = f0.w;
              v[rr][8 * i + 4] = f1.x; v[rr][8 * i + 5] = f1.y; v[rr][8 * i + 6] = f1.z; v[rr][8 * i + 7] = f1.w;
            }
          }
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int rowl = lw * 64 + r4 * 4 + rr;
            const bool ok = row0 + r4 * 4 + rr < a.M;
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) s += v[rr][e];
            const float mean = warp_sum_f(s) * (1.0f / (float)KD);
            float sq = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) { const float d = v[rr][e] - mean; sq += d * d; }
            const float rstd = 1.0f / sqrtf(warp_sum_f(sq) * (1.0f / (float)KD) + a.eps);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float o0 = (v[rr][8 * i + 2 * e] - mean) * rstd * g[8 * i + 2 * e] + bt[8 * i + 2 * e];
                const float o1 = (v[rr][8 * i + 2 * e + 1] - mean) * rstd * g[8 * i + 2 * e + 1] + bt[8 * i + 2 * e + 1];
                pk[e] = ok ? pack_bf16x2(o0, o1) : 0u;   // rows past M are zero operands
              }
              const int kb = 4 * i + (lane >> 3), j = lane & 7;
              sts128(sA_u + kb * A_SUB + rowl * 128 + ((j ^ (rowl & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive_even_cta_release(a_full);
      }
    }
  } else {
    // ---- epilogue: bias (+ReLU) -> bf16 -> TMA store (same as gemm.cu EPI 1) ----
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const uint32_t sbias = smem_u32(sBias + ew * BIAS_BYTES);
    const uint32_t stage_tile = smem_u32(sStg + ew * STG_BYTES);
    const uint32_t my_stage_row = stage_tile + lane * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = u0; u < u1; ++u) {
      const int m_blk = u / n_tiles, n_blk = u % n_tiles;
      const int row_base = m_blk * 2 * BM + (int)cta * BM + quarter * 32;
      const int col_base = n_blk * BN + half * 128;
      if (a.bias) {
        const uint4 b = ldg128_nc(a.bias + col_base + lane * 4);
        sts128(sbias + lane * 16, b.x, b.y, b.z, b.w);
        warp_sync_smem();
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * 128);
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(tbase, ra);
      tmem_ld_32x32(tbase + 32, rb);
#pragma unroll
      for (int c64 = 0; c64 < 2; ++c64) {
        tmem_ld_wait();
        uint32_t pk[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t (&r)[32] = hh ? rb : ra;
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) {
            float v0 = __uint_as_float(r[4 * gq]), v1 = __uint_as_float(r[4 * gq + 1]);
            float v2 = __uint_as_float(r[4 * gq + 2]), v3 = __uint_as_float(r[4 * gq + 3]);
            if (a.bias) {
              const float4 bb = lds128f(sbias + (c64 * 64 + hh * 32 + 4 * gq) * 4);
              v0 += bb.x; v1 += bb.y; v2 += bb.z; v3 += bb.w;
            }
            if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
            pk[hh * 16 + 2 * gq] = pack_bf16x2(v0, v1);
            pk[hh * 16 + 2 * gq + 1] = pack_bf16x2(v2, v3);
          }
        }
        if (c64 == 0) {
          tmem_ld_32x32(tbase + 64, ra);
          tmem_ld_32x32(tbase + 96, rb);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_even_cta(&tempty[acc]);
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(my_stage_row + ((j ^ (lane & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, stage_tile, col_base + c64 * 64, row_base);
          tma_store_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

}  // namespace

int gemm_ln_bf16_tcgen05(const float* x, int ldx, int M, const float* gamma, const float* beta, float eps, const __nv_bfloat16* W, int N,
                         const float* bias, int relu, __nv_bfloat16* out, int ld_out, int num_sms, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if ((N % BN) || (ldx & 3) || (ld_out & 7) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(x) & 15))
    return (int)cudaErrorInvalidValue;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t err = cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return (int)err;
  }
  CUtensorMap tmB, tmC;
  int rc = make_tmap_bf16_sw128(&tmB, W, (uint64_t)N, (uint64_t)KD, (uint64_t)KD, BN / 2);
  if (rc) return rc;
  rc = make_tmap_bf16_sw128(&tmC, out, (uint64_t)M, (uint64_t)N, (uint64_t)ld_out, 32, 64);
  if (rc) return rc;
  LnArgs a;
  a.x = x; a.ldx = ldx; a.gamma = gamma; a.beta = beta; a.eps = eps; a.M = M; a.N = N; a.bias = bias; a.relu = relu;
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM), n_tiles = N / BN;
  int grid = 2 * m_tiles * n_tiles;
  if (grid > (num_sms & ~1)) grid = num_sms & ~1;
  return launch_kernel(gemm_ln_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, tmB, tmC, a);
}

}  // namespace pf
