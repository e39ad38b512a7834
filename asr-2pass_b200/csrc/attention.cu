// tcgen05 flash attention, head dim 128, variable length, non-causal with key-length masking.
//
// One CTA = one (segment, 128-query tile, head).  Two passes over the keys in blocks of 64:
//   pass 1:  S = Q K^T            -> exact row maxima m            (no exponentials, no P V)
//   pass 2:  S = Q K^T again,  P = exp2((S - m) * scale*log2e)  (bf16, swizzled into smem),  O += P V
// so O never needs the running-max rescale of single-pass flash attention and every probability is in
// [0,1] exactly as in the reference softmax.  O and the row sums l stay fp32; O/l is applied once.
//
//   warp 0      : TMA producer (Q once; K for pass 1, K+V for pass 2; 3-stage ring) + TMEM alloc
//   warp 1      : MMA issuer   (S: 128x64x128 from smem Q,K; PV: 128x128x64, P K-major, V MN-major)
//   warps 2..5  : softmax / epilogue, one query row per thread (tcgen05.ld 32x32b)
// TMEM: S double buffered (2 x 64 columns) + O (128 columns) = 256 columns.
#include "attention.cuh"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace pf {
namespace {

constexpr int BQ = 128, BKV = 64, HD = 128, KV_STAGES = 2, P_BUFS = 1;
constexpr int Q_BYTES = BQ * HD * 2;       // 32768: two 128x64 swizzled boxes
constexpr int K_BYTES = BKV * HD * 2;      // 16384: two 64x64 boxes
constexpr int STAGE_BYTES = 2 * K_BYTES;   // K + V
constexpr int P_BYTES = BQ * BKV * 2;      // 16384
constexpr int kThreads = 192;
constexpr int kTmemCols = 256;
// 112 KB + barriers: two CTAs fit one SM (2 x 256 TMEM columns, 2 x (112.25 + 1) KB shared memory), so one CTA's
// TMA / softmax latency is covered by the other's MMAs.
constexpr int kSmemBytes = Q_BYTES + KV_STAGES * STAGE_BYTES + P_BUFS * P_BYTES + 256;

struct AArgs {
  const int* q_row_off;
  const int* q_len;
  const int* kv_row_off;
  const int* kv_len;
  const AttnWork* work;
  __nv_bfloat16* out;
  int ldo;
  int q_col0, k_col0, v_col0;
  float scale_log2e;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ONLINE = single pass over the keys with a running row maximum (flash-attention style): S_j -> P_j -> O += P_j V_j per key
// block, no separate maximum pass, so a CTA makes half as many dependent TMA -> MMA -> softmax round trips (the kernel is
// latency bound at this path's 30-340-key segments).  The maximum the exponentials are taken against ("m_used") is only
// advanced when the true running maximum has moved by more than 8 in the log2 domain; then, and only then, O (TMEM) and
// l are rescaled by 2^(m_used_old - m_used_new).  Probabilities therefore stay <= 2^8 and the normalised result O / l
// is the same softmax(S) V up to rounding.  ONLINE = false is the exact two-pass variant described at the top.
template <bool ONLINE>
__global__ void __launch_bounds__(kThreads, 2)
attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AArgs a) {
  pdl_wait();   // q_len of the decoder is produced by the CIF kernels
  pdl_launch_dependents();
  const AttnWork w = a.work[blockIdx.x];
  const int h = blockIdx.y;
  const int Tq = a.q_len[w.seg];
  const int Tk = a.kv_len[w.seg];
  if (w.q0 >= Tq || Tk <= 0) return;  // whole CTA, before any barrier
  const int nb = (Tk + BKV - 1) / BKV;
  const int q_row = a.q_row_off[w.seg] + w.q0;
  const int kv_row = a.kv_row_off[w.seg];

  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B tiles need 1024-byte alignment (checked below)
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + Q_BYTES;
  uint8_t* sP = sKV + KV_STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BUFS * P_BYTES);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // 3
  uint64_t* kv_empty = bars + 4;            // 3
  uint64_t* s_full = bars + 7;              // 2
  uint64_t* s_empty = bars + 9;             // 2
  uint64_t* p_full = bars + 11;             // 2
  uint64_t* p_empty = bars + 13;            // 2
  uint64_t* o_full = bars + 15;             // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, a.q_col0 + h * HD, q_row);
      tma_load_2d(sQ + Q_BYTES / 2, &tmQ, q_full, a.q_col0 + h * HD + 64, q_row);
      const int n_it = ONLINE ? nb : 2 * nb;
      for (int it = 0; it < n_it; ++it) {
        const int stage = it % KV_STAGES;
        const uint32_t phase = (it / KV_STAGES) & 1;
        const bool pass2 = ONLINE || it >= nb;
        const int j = (!ONLINE && pass2) ? it - nb : it;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], pass2 ? STAGE_BYTES : K_BYTES);
        uint8_t* dst = sKV + stage * STAGE_BYTES;
        const int r = kv_row + j * BKV;
        tma_load_2d(dst, &tmKV, &kv_full[stage], a.k_col0 + h * HD, r);
        tma_load_2d(dst + K_BYTES / 2, &tmKV, &kv_full[stage], a.k_col0 + h * HD + 64, r);
        if (pass2) {
          tma_load_2d(dst + K_BYTES, &tmKV, &kv_full[stage], a.v_col0 + h * HD, r);
          tma_load_2d(dst + K_BYTES + K_BYTES / 2, &tmKV, &kv_full[stage], a.v_col0 + h * HD + 64, r);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (=V) is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_s = [&](int it) {
        const int stage = it % KV_STAGES;
        const int sb = it & 1;
        mbar_wait(&kv_full[stage], (it / KV_STAGES) & 1);
        mbar_wait(&s_empty[sb], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + stage * STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint64_t da = umma_desc_sw128(q_addr + (ks >> 2) * (Q_BYTES / 2)) + 2 * (ks & 3);
          const uint64_t db = umma_desc_sw128(k_addr + (ks >> 2) * (K_BYTES / 2)) + 2 * (ks & 3);
          umma_bf16(tmem_base + sb * BKV, da, db, idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
      };
      const int it0 = ONLINE ? 0 : nb;
      if (!ONLINE) {
        // pass 1: row maxima only
        for (int it = 0; it < nb; ++it) {
          issue_s(it);
          umma_commit(&kv_empty[it % KV_STAGES]);
        }
      }
      // (pass 2) S of block j+1 is issued before P V of block j so the tensor pipe overlaps the softmax
      issue_s(it0);
      for (int jj = 0; jj < nb; ++jj) {
        const int it = it0 + jj;
        if (jj + 1 < nb) issue_s(it + 1);
        const int pb = jj % P_BUFS;
        mbar_wait(&p_full[pb], (jj / P_BUFS) & 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + pb * P_BYTES);
        const uint32_t v_addr = smem_u32(sKV + (it % KV_STAGES) * STAGE_BYTES + K_BYTES);
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const uint64_t da = umma_desc_sw128(p_addr) + 2 * ks;
          // V tile: 64 keys x 128 d as two [64 x 64] boxes 8 KB apart; 16 keys = 2048 B per K step
          const uint64_t db = umma_desc_sw128_mn(v_addr + ks * 2048, K_BYTES / 2);
          umma_bf16(tmem_o, da, db, idesc_pv, (jj | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[it % KV_STAGES]);
        umma_commit(&p_empty[pb]);
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float m = -INFINITY;
    uint32_t r[32];
    float l = 0.f;
    if constexpr (ONLINE) {
      float m_used = 0.f;   // the maximum the exponentials are taken against (set by the first block)
      uint32_t r2[32];
      for (int jj = 0; jj < nb; ++jj) {
        const int sb = jj & 1;
        mbar_wait(&s_full[sb], (jj >> 1) & 1);
        tc_fence_after();
        const int nvalid = Tk - jj * BKV;
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV, r);
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV + 32, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);   // the scores are in registers
        float mb = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < nvalid) mb = fmaxf(mb, __uint_as_float(r[c]));
          if (32 + c < nvalid) mb = fmaxf(mb, __uint_as_float(r2[c]));
        }
        m = fmaxf(m, mb);
        float factor = 1.f;
        bool need = false;
        if (jj == 0) {
          m_used = m;
        } else if ((m - m_used) * a.scale_log2e > 8.f) {
          need = true;
          factor = ex2((m_used - m) * a.scale_log2e);
          m_used = m;
          l *= factor;
        }
        const float mc = m_used * a.scale_log2e;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float p0 = (c < nvalid) ? ex2(fmaf(__uint_as_float(r[c]), a.scale_log2e, -mc)) : 0.f;
          const float p1 = (c + 1 < nvalid) ? ex2(fmaf(__uint_as_float(r[c + 1]), a.scale_log2e, -mc)) : 0.f;
          const float p2 = (32 + c < nvalid) ? ex2(fmaf(__uint_as_float(r2[c]), a.scale_log2e, -mc)) : 0.f;
          const float p3 = (33 + c < nvalid) ? ex2(fmaf(__uint_as_float(r2[c + 1]), a.scale_log2e, -mc)) : 0.f;
          l += (p0 + p1) + (p2 + p3);
          pk[c >> 1] = pack_bf16x2(p0, p1);
          pk[16 + (c >> 1)] = pack_bf16x2(p2, p3);
        }
        // P V of the previous block has completed once p_empty flips: P's buffer is free and O is quiescent
        mbar_wait(&p_empty[0], (jj & 1) ^ 1);
        if (__any_sync(0xffffffffu, need)) {
          // rare: this warp's 32 rows of O are rescaled in TMEM (rows that did not move use factor 1)
          tc_fence_after();
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            tmem_ld_32x32(tmem_o + lane_addr + c4 * 32, r);
            tmem_ld_wait();
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              lo[c] = __float_as_uint(__uint_as_float(r[c]) * factor);
              hi[c] = __float_as_uint(__uint_as_float(r[16 + c]) * factor);
            }
            tmem_st_32x16(tmem_o + lane_addr + c4 * 32, lo);
            tmem_st_32x16(tmem_o + lane_addr + c4 * 32 + 16, hi);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        const uint32_t prow = smem_u32(sP + row * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(prow + ((j ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[0]);
      }
    } else {
    // ---- pass 1 ----
    for (int it = 0; it < nb; ++it) {
      const int sb = it & 1;
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      tc_fence_after();
      const int nvalid = Tk - it * BKV;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (half * 32 + c < nvalid) m = fmaxf(m, __uint_as_float(r[c]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
    }
    // ---- pass 2 ----
    const float mc = m * a.scale_log2e;
    for (int jj = 0; jj < nb; ++jj) {
      const int it = nb + jj;
      const int sb = it & 1;
      const int pb = jj % P_BUFS;
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      tc_fence_after();
      const int nvalid = Tk - jj * BKV;
      uint32_t pk[32];  // 64 probabilities packed as bf16x2
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const int k0 = half * 32 + c;
          float p0 = (k0 < nvalid) ? ex2(fmaf(__uint_as_float(r[c]), a.scale_log2e, -mc)) : 0.f;
          float p1 = (k0 + 1 < nvalid) ? ex2(fmaf(__uint_as_float(r[c + 1]), a.scale_log2e, -mc)) : 0.f;
          l += p0 + p1;
          pk[half * 16 + (c >> 1)] = pack_bf16x2(p0, p1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      // P row -> smem, K-major SWIZZLE_128B: 16-byte chunk j of row r lives at chunk (j ^ (r & 7))
      mbar_wait(&p_empty[pb], ((jj / P_BUFS) & 1) ^ 1);
      const uint32_t prow = smem_u32(sP + pb * P_BYTES + row * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(prow + ((j ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
    }
    }
    // ---- epilogue ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const bool row_ok = (w.q0 + row) < Tq;
    __nv_bfloat16* orow = a.out + (size_t)(q_row + row) * a.ldo + h * HD;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      tmem_ld_32x32(tmem_o + lane_addr + c4 * 32, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]) * inv, __uint_as_float(r[8 * g + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv);
          stg128(orow + c4 * 32 + g * 8, o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ---------------------------------------------------------------------------------------------
// Product kernel: single pass (running maximum, lazy rescale — see above) AND all heads of one (segment, query tile)
// pipelined through ONE CTA.  At this path's lengths a (tile, head) work item is only 1-6 key blocks, so a CTA per item
// spends most of its life in set-up, first-load latency and drain.  Here the TMA producer, the MMA issuer and the softmax
// warps run one continuous stream of n_heads x nb key blocks: barriers, TMEM and the K/V ring are set up once, the K/V
// loads of head h+1 are in flight while head h finishes, and head h's output is written while S of head h+1 is computed.
// Q is single buffered: its reload waits for the last S MMA of the previous head (q_empty), O is handed over through
// o_full / o_empty.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
attn_heads_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AArgs a, int n_heads) {
  pdl_wait();   // q_len of the decoder is produced by the CIF kernels
  pdl_launch_dependents();
  const AttnWork w = a.work[blockIdx.x];
  const int Tq = a.q_len[w.seg];
  const int Tk = a.kv_len[w.seg];
  if (w.q0 >= Tq || Tk <= 0) return;  // whole CTA, before any barrier
  const int nb = (Tk + BKV - 1) / BKV;
  const int q_row = a.q_row_off[w.seg] + w.q0;
  const int kv_row = a.kv_row_off[w.seg];

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + Q_BYTES;
  uint8_t* sP = sKV + KV_STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BUFS * P_BYTES);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* q_full = bars;                  // 1
  uint64_t* q_empty = bars + 1;             // 1
  uint64_t* kv_full = bars + 2;             // 2
  uint64_t* kv_empty = bars + 4;            // 2
  uint64_t* s_full = bars + 6;              // 2
  uint64_t* s_empty = bars + 8;             // 2
  uint64_t* p_full = bars + 10;             // 1
  uint64_t* p_empty = bars + 11;            // 1
  uint64_t* o_full = bars + 12;             // 1
  uint64_t* o_empty = bars + 13;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    mbar_init(p_full, 4); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 4);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;   // key blocks streamed so far, over all heads
      for (int h = 0; h < n_heads; ++h) {
        mbar_wait(q_empty, (uint32_t)((h & 1) ^ 1));
        mbar_arrive_expect_tx(q_full, Q_BYTES);
        tma_load_2d(sQ, &tmQ, q_full, a.q_col0 + h * HD, q_row);
        tma_load_2d(sQ + Q_BYTES / 2, &tmQ, q_full, a.q_col0 + h * HD + 64, q_row);
        for (int j = 0; j < nb; ++j, ++g) {
          const int stage = g % KV_STAGES;
          mbar_wait(&kv_empty[stage], (uint32_t)(((g / KV_STAGES) & 1) ^ 1));
          mbar_arrive_expect_tx(&kv_full[stage], STAGE_BYTES);
          uint8_t* dst = sKV + stage * STAGE_BYTES;
          const int r = kv_row + j * BKV;
          tma_load_2d(dst, &tmKV, &kv_full[stage], a.k_col0 + h * HD, r);
          tma_load_2d(dst + K_BYTES / 2, &tmKV, &kv_full[stage], a.k_col0 + h * HD + 64, r);
          tma_load_2d(dst + K_BYTES, &tmKV, &kv_full[stage], a.v_col0 + h * HD, r);
          tma_load_2d(dst + K_BYTES + K_BYTES / 2, &tmKV, &kv_full[stage], a.v_col0 + h * HD + 64, r);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (=V) is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int g, bool last_of_head) {
        const int stage = g % KV_STAGES;
        const int sb = g & 1;
        mbar_wait(&kv_full[stage], (uint32_t)((g / KV_STAGES) & 1));
        mbar_wait(&s_empty[sb], (uint32_t)(((g >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + stage * STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint64_t da = umma_desc_sw128(q_addr + (ks >> 2) * (Q_BYTES / 2)) + 2 * (ks & 3);
          const uint64_t db = umma_desc_sw128(k_addr + (ks >> 2) * (K_BYTES / 2)) + 2 * (ks & 3);
          umma_bf16(tmem_base + sb * BKV, da, db, idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
        if (last_of_head) umma_commit(q_empty);   // Q may be replaced once every S MMA of this head has read it
      };
      int g0 = 0;
      for (int h = 0; h < n_heads; ++h, g0 += nb) {
        mbar_wait(q_full, (uint32_t)(h & 1));
        tc_fence_after();
        issue_s(g0, nb == 1);
        for (int jj = 0; jj < nb; ++jj) {
          const int g = g0 + jj;
          if (jj + 1 < nb) issue_s(g + 1, jj + 2 == nb);
          mbar_wait(p_full, (uint32_t)(g & 1));
          if (jj == 0) mbar_wait(o_empty, (uint32_t)((h & 1) ^ 1));   // the previous head's O has been read out
          tc_fence_after();
          const uint32_t p_addr = smem_u32(sP);
          const uint32_t v_addr = smem_u32(sKV + (g % KV_STAGES) * STAGE_BYTES + K_BYTES);
#pragma unroll
          for (int ks = 0; ks < BKV / 16; ++ks) {
            const uint64_t da = umma_desc_sw128(p_addr) + 2 * ks;
            const uint64_t db = umma_desc_sw128_mn(v_addr + ks * 2048, K_BYTES / 2);
            umma_bf16(tmem_o, da, db, idesc_pv, (jj | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&kv_empty[g % KV_STAGES]);
          umma_commit(p_empty);
        }
        umma_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const bool row_ok = (w.q0 + row) < Tq;
    uint32_t r[32], r2[32];
    int g0 = 0;
    for (int h = 0; h < n_heads; ++h, g0 += nb) {
      float m = -INFINITY, l = 0.f, m_used = 0.f;
      for (int jj = 0; jj < nb; ++jj) {
        const int g = g0 + jj;
        const int sb = g & 1;
        mbar_wait(&s_full[sb], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        const int nvalid = Tk - jj * BKV;
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV, r);
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV + 32, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);   // the scores are in registers
        float mb = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < nvalid) mb = fmaxf(mb, __uint_as_float(r[c]));
          if (32 + c < nvalid) mb = fmaxf(mb, __uint_as_float(r2[c]));
        }
        m = fmaxf(m, mb);
        float factor = 1.f;
        bool need = false;
        if (jj == 0) {
          m_used = m;
        } else if ((m - m_used) * a.scale_log2e > 8.f) {
          need = true;
          factor = ex2((m_used - m) * a.scale_log2e);
          m_used = m;
          l *= factor;
        }
        const float mc = m_used * a.scale_log2e;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float p0 = (c < nvalid) ? ex2(fmaf(__uint_as_float(r[c]), a.scale_log2e, -mc)) : 0.f;
          const float p1 = (c + 1 < nvalid) ? ex2(fmaf(__uint_as_float(r[c + 1]), a.scale_log2e, -mc)) : 0.f;
          const float p2 = (32 + c < nvalid) ? ex2(fmaf(__uint_as_float(r2[c]), a.scale_log2e, -mc)) : 0.f;
          const float p3 = (33 + c < nvalid) ? ex2(fmaf(__uint_as_float(r2[c + 1]), a.scale_log2e, -mc)) : 0.f;
          l += (p0 + p1) + (p2 + p3);
          pk[c >> 1] = pack_bf16x2(p0, p1);
          pk[16 + (c >> 1)] = pack_bf16x2(p2, p3);
        }
        // P V of the previous block has completed once p_empty flips: P's buffer is free and O is quiescent
        mbar_wait(p_empty, (uint32_t)((g & 1) ^ 1));
        if (__any_sync(0xffffffffu, need)) {
          tc_fence_after();
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            tmem_ld_32x32(tmem_o + lane_addr + c4 * 32, r);
            tmem_ld_wait();
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              lo[c] = __float_as_uint(__uint_as_float(r[c]) * factor);
              hi[c] = __float_as_uint(__uint_as_float(r[16 + c]) * factor);
            }
            tmem_st_32x16(tmem_o + lane_addr + c4 * 32, lo);
            tmem_st_32x16(tmem_o + lane_addr + c4 * 32 + 16, hi);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        const uint32_t prow = smem_u32(sP + row * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(prow + ((j ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // ---- this head's output ----
      mbar_wait(o_full, (uint32_t)(h & 1));
      tc_fence_after();
      const float inv = 1.f / l;
      __nv_bfloat16* orow = a.out + (size_t)(q_row + row) * a.ldo + h * HD;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        tmem_ld_32x32(tmem_o + lane_addr + c4 * 32, r);
        tmem_ld_wait();
        if (c4 == 3) {   // O is in registers: the next head's first P V may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
        if (row_ok) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[8 * gq + 0]) * inv, __uint_as_float(r[8 * gq + 1]) * inv);
            o.y = pack_bf16x2(__uint_as_float(r[8 * gq + 2]) * inv, __uint_as_float(r[8 * gq + 3]) * inv);
            o.z = pack_bf16x2(__uint_as_float(r[8 * gq + 4]) * inv, __uint_as_float(r[8 * gq + 5]) * inv);
            o.w = pack_bf16x2(__uint_as_float(r[8 * gq + 6]) * inv, __uint_as_float(r[8 * gq + 7]) * inv);
            stg128(orow + c4 * 32 + gq * 8, o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// CUDA-core cross-check (test-only): one warp per query row, online softmax in fp32.
// ---------------------------------------------------------------------------------------------
__global__ void attn_check_kernel(AArgs a, const __nv_bfloat16* q, int ldq, const __nv_bfloat16* kv, int ldkv,
                                  float scale) {
  pdl_wait();
  pdl_launch_dependents();
  const AttnWork w = a.work[blockIdx.x];
  const int h = blockIdx.y;
  const int Tq = a.q_len[w.seg], Tk = a.kv_len[w.seg];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = warp; rr < BQ; rr += blockDim.x >> 5) {
    const int qi = w.q0 + rr;
    if (qi >= Tq) break;
    const __nv_bfloat16* qp = q + (size_t)(a.q_row_off[w.seg] + qi) * ldq + a.q_col0 + h * HD + lane * 4;
    float qv[4], o[4] = {0, 0, 0, 0};
    for (int d = 0; d < 4; ++d) qv[d] = __bfloat162float(qp[d]);
    float m = -INFINITY, l = 0.f;
    for (int k = 0; k < Tk; ++k) {
      const __nv_bfloat16* kp = kv + (size_t)(a.kv_row_off[w.seg] + k) * ldkv + a.k_col0 + h * HD + lane * 4;
      const __nv_bfloat16* vp = kv + (size_t)(a.kv_row_off[w.seg] + k) * ldkv + a.v_col0 + h * HD + lane * 4;
      float s = 0.f;
      for (int d = 0; d < 4; ++d) s += qv[d] * __bfloat162float(kp[d]);
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      s *= scale;
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      l = l * corr + p;
      for (int d = 0; d < 4; ++d) o[d] = o[d] * corr + p * __bfloat162float(vp[d]);
      m = mn;
    }
    __nv_bfloat16* op = a.out + (size_t)(a.q_row_off[w.seg] + qi) * a.ldo + h * HD + lane * 4;
    for (int d = 0; d < 4; ++d) op[d] = __float2bfloat16(o[d] / l);
  }
}

AArgs make_args(const AttnProblem& p) {
  AArgs a;
  a.q_row_off = p.q_row_off; a.q_len = p.q_len; a.kv_row_off = p.kv_row_off; a.kv_len = p.kv_len;
  a.work = p.work; a.out = p.out; a.ldo = p.ldo;
  a.q_col0 = p.q_col0; a.k_col0 = p.k_col0; a.v_col0 = p.v_col0;
  a.scale_log2e = p.scale * 1.4426950408889634f;
  return a;
}

}  // namespace

int attention_tcgen05(const AttnProblem& p, cudaStream_t stream) {
  if (p.n_work <= 0) return 0;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t err = cudaFuncSetAttribute(attn_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return (int)err;
    err = cudaFuncSetAttribute(attn_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return (int)err;
    err = cudaFuncSetAttribute(attn_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return (int)err;
  }
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_bf16_sw128(&tmQ, p.q, (uint64_t)p.q_rows, (uint64_t)p.ldq, (uint64_t)p.ldq, BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_sw128(&tmKV, p.kv, (uint64_t)p.kv_rows, (uint64_t)p.ldkv, (uint64_t)p.ldkv, BKV);
  if (rc) return rc;
  AArgs a = make_args(p);
  if (p.online == 2) return launch_kernel(attn_heads_kernel, dim3(p.n_work), dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a, p.n_heads);
  dim3 grid(p.n_work, p.n_heads);
  if (p.online) return launch_kernel(attn_tcgen05_kernel<true>, grid, dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a);
  return launch_kernel(attn_tcgen05_kernel<false>, grid, dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a);
}

int attention_check_kernel(const AttnProblem& p, cudaStream_t stream) {
  if (p.n_work <= 0) return 0;
  AArgs a = make_args(p);
  dim3 grid(p.n_work, p.n_heads);
  return launch_kernel(attn_check_kernel, grid, dim3(256), 0, stream, a, p.q, p.ldq, p.kv, p.ldkv, p.scale);
}

}  // namespace pf
