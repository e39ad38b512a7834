// Persistent warp-specialised tcgen05 GEMM for sm_100a, CTA-pair edition (see gemm.cuh for what it replaces).
//
// Two CTAs of a cluster (one TPC) work on one 256x256 output tile with tcgen05.mma.cta_group::2: each CTA
// stages its own 128 rows of A and HALF of the 256 W rows per 64-wide K block (32 KB per stage instead of
// 48 KB), which halves the L2->SMEM traffic per FLOP and the shared-memory read bandwidth the MMA needs.
//
//   warp 0      : TMA producer (both CTAs; completion bytes land on the even CTA's "full" barrier)
//   warp 1      : MMA issuer   (even CTA only, one lane): 256x256x16 per instruction, accumulators in the TMEM
//                 of both CTAs (128 lanes x 256 columns each, double buffered = 512 columns)
//   warp 2      : TMEM allocator (cta_group::2)
//   warps 4..11 : epilogue in each CTA: tcgen05.ld 32x32b.x32 -> bias / ReLU / FSMN-memory add / residual ->
//                 bf16 or fp32.  bf16-only outputs (EPI 1) leave through TMA stores: each warp packs 32 rows x 64
//                 columns into a swizzled 4 KB smem tile and one lane issues cp.async.bulk.tensor (full 128-byte
//                 lines, no LSU / register traffic, asynchronous).  The general path stages through a per-warp smem
//                 tile so every global load and store is a run of full 32-byte sectors; optional fused argmax.
#include <mutex>
#include <unordered_map>

#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace pf {

namespace {

constexpr int BM = 128;        // rows per CTA (256 per pair)
constexpr int BN = 256;        // columns per pair tile; each CTA loads BN/2 rows of W
constexpr int BK = 64, STAGES = 5;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = (BN / 2) * BK * 2;
constexpr int kTmemCols = 512;
constexpr int SCR_STRIDE = 144;                  // bytes per scratch row: 128 + 16 (bank-conflict-free 16 B accesses)
constexpr int SCR_BYTES = 32 * SCR_STRIDE;       // per epilogue warp
constexpr int BIAS_BYTES = 256 * 4;                  // per epilogue warp (128 floats used)
constexpr int kSmemBytes = STAGES * (A_BYTES + B_BYTES) + kEpiWarps * (SCR_BYTES + BIAS_BYTES) + 1024 /*align*/ + 512 /*barriers*/;

struct KArgs {
  int M, N, K;
  const int* m_dev;
  int a_k_wrap, a_row_shift0;
  GemmEpilogue e;
};

// EPI selects the epilogue:
//   0  general: any combination of GemmEpilogue's fields (argmax, separate residual source, fp32 + bf16 outputs ...)
//   1  bias (+ReLU) -> bf16 through TMA stores (QKV, FFN1, decoder q / kv projections, LSTM input projections)
//   2  bias (+ReLU) (+bf16 addend) -> fp32 through TMA; when the residual is the output buffer itself (x += ..., every
//      out-projection and FFN2) the tile leaves as a TMA REDUCE-ADD, so the SMs never load the residual stream: the
//      read-modify-write of x happens in L2.
// F16: the 16-bit operands and outputs are IEEE fp16 instead of bf16 (engine precision 1, see ptx.cuh pack_h2).
//   3  = 1 + per-row partial sums of the rounded outputs (LayerNorm fold, producer side; GemmEpilogue::row_stats_out)
//   4  = 2 + LayerNorm applied from those sums (consumer side; GemmEpilogue::ln_stats) -- separate instantiations, so that the
//        hot epilogues 1 and 2 carry none of the fold's branches or registers
// BNT: tile width.  256 everywhere except the small-batch path: when a GEMM has fewer 256-wide tiles than half the CTA pairs
// (a batch-1 forward: M <= 256, N = 512 -> TWO pairs walk K = 2048 while 72 idle), the TMA epilogues 1 and 2 also exist with
// 128-wide tiles -- twice the pairs, half the MMA time per K block; the accumulation order over K is unchanged, so results are
// bit-identical to the wide tiles (batch invariance).
template <int EPI_, bool F16, int BNT = BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, KArgs a) {
  constexpr int EPI = EPI_ == 3 ? 1 : (EPI_ == 4 ? 2 : EPI_);
  constexpr bool kRowSums = EPI_ == 3, kLnApply = EPI_ == 4;
  constexpr int WSPAN = BNT / 2;                     // columns an epilogue warp owns (two warps per row quarter)
  constexpr int B_TILE_BYTES = (BNT / 2) * BK * 2;   // what one CTA loads of W per K block (the stage stride stays B_BYTES)
  static_assert(BNT == BN || (BNT == 128 && (EPI_ == 1 || EPI_ == 2)), "narrow tiles: TMA epilogues only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sScr = smem + STAGES * (A_BYTES + B_BYTES);
  uint8_t* sBias = sScr + kEpiWarps * SCR_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kEpiWarps * BIAS_BYTES);
  uint64_t* full = bars;                 // used on the even CTA
  uint64_t* empty = bars + STAGES;       // one set per CTA, signalled by the multicast commit
  uint64_t* tfull = bars + 2 * STAGES;   // one set per CTA
  uint64_t* tempty = tfull + 2;          // used on the even CTA: 2 x kEpiWarps arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int N = a.N;
  const int n_tiles = (N + BNT - 1) / BNT;
  const int k_blocks = (a.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI != 0) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlaps the previous kernel's tail (PDL); activations are only touched from here on.
  pdl_wait();
  pdl_launch_dependents();
  const int M = a.m_dev ? *a.m_dev : a.M;
  const int m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int total = m_tiles * n_tiles;

  // Warps 0 and 1 run converged and put only the TMA / tcgen05 instructions under elect_one(): inside an `if (lane == 0)` region
  // the compiler wraps each of them in a per-thread "waterfall" loop (ELECT + branch), ~75 cycles per instruction (attention.cu).
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total; tile += n_pairs) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int m0 = m_blk * 2 * BM + (int)cta * BM;
      const int n0 = n_blk * BNT + (int)cta * (BNT / 2);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        const int k0 = kb * BK;
        int ac0 = k0, ac1 = m0;
        if (a.a_k_wrap > 0) {
          const int pass = k0 / a.a_k_wrap;
          ac0 = k0 - pass * a.a_k_wrap;
          ac1 += pass + a.a_row_shift0;
        }
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + B_TILE_BYTES));
          tma_load_2d_2cta(sA + stage * A_BYTES, &tmA, &full[stage], ac0, ac1);
          tma_load_2d_2cta(sB + stage * B_BYTES, &tmB, &full[stage], k0, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_h16(F16, 2 * BM, BNT);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total; tile += n_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BNT;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * A_BYTES));
          const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * B_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // +32 B per 16-element K step inside the 128 B swizzle atom (address field is >>4)
              umma_bf16_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2cta_mc(&empty[stage], 0x3);
            if (kb + 1 == k_blocks) umma_commit_2cta_mc(&tfull[acc], 0x3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may address
    const int half = ew >> 2;      // which 128-column half of the 256-wide tile
    const GemmEpilogue& e = a.e;
    const uint32_t scr = smem_u32(sScr + ew * SCR_BYTES);          // shared-space addresses (STS/LDS, no generic ops)
    const uint32_t my_row = scr + lane * SCR_STRIDE;
    const uint32_t sbias = smem_u32(sBias + ew * BIAS_BYTES);
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (EPI == 1) {
      // bias (+ReLU) -> bf16 -> TMA store.  Warp (quarter, half) owns rows 32 quarter.. of the CTA's 128 and columns
      // 128 half.. of the 256-wide tile, as two 64-column boxes.  Its 4 KB staging tile is written in the tensor map's
      // SWIZZLE_128B layout (16-byte chunk j of row r at chunk j ^ (r & 7)), which also makes the 32 row-wise 16-byte
      // stores of a warp conflict free.  One lane issues the store and is the only one to wait on it
      // (bulk async-groups are per thread); the next TMEM load is already in flight by then.
      const uint32_t stage_tile = smem_u32(sScr + ew * 4096);
      const uint32_t my_stage_row = stage_tile + lane * 128;
      for (int tile = pair; tile < total; tile += n_pairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int row_base = m_blk * 2 * BM + (int)cta * BM + quarter * 32;
        const int col_base = n_blk * BNT + half * WSPAN;
        if (e.bias) {  // this warp's bias values: one coalesced load per tile, read back as smem broadcasts
          if (lane * 4 < WSPAN) {
            const uint4 b = ldg128_nc(e.bias + col_base + lane * 4);
            sts128(sbias + lane * 16, b.x, b.y, b.z, b.w);
          }
          warp_sync_smem();
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BNT + half * WSPAN);
        uint32_t ra[32], rb[32];
        float st1 = 0.f, st2 = 0.f;   // row_stats_out: this row's sum / sum of squares over the warp's 128 columns
        if (e.dbg == 3) {  // micro-benchmark only: no TMEM reads
#pragma unroll
          for (int i = 0; i < 32; ++i) { ra[i] = 0; rb[i] = 0; }
        }
        if (e.dbg != 3) {
          tmem_ld_32x32(tbase, ra);
          tmem_ld_32x32(tbase + 32, rb);
        }
#pragma unroll
        for (int c64 = 0; c64 < WSPAN / 64; ++c64) {
          tmem_ld_wait();
          uint32_t pk[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t (&r)[32] = hh ? rb : ra;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float v0 = __uint_as_float(r[4 * g]), v1 = __uint_as_float(r[4 * g + 1]);
              float v2 = __uint_as_float(r[4 * g + 2]), v3 = __uint_as_float(r[4 * g + 3]);
              if (e.bias) {
                const float4 bb = lds128f(sbias + (c64 * 64 + hh * 32 + 4 * g) * 4);
                v0 += bb.x; v1 += bb.y; v2 += bb.z; v3 += bb.w;
              }
              if (e.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
              pk[hh * 16 + 2 * g] = pack_h2<F16>(v0, v1);
              pk[hh * 16 + 2 * g + 1] = pack_h2<F16>(v2, v3);
              if constexpr (kRowSums) {   // of the values the consumer GEMM will read
                const float2 q0 = unpack_h2<F16>(pk[hh * 16 + 2 * g]), q1 = unpack_h2<F16>(pk[hh * 16 + 2 * g + 1]);
                st1 += (q0.x + q0.y) + (q1.x + q1.y);
                st2 = fmaf(q0.x, q0.x, st2); st2 = fmaf(q0.y, q0.y, st2); st2 = fmaf(q1.x, q1.x, st2); st2 = fmaf(q1.y, q1.y, st2);
              }
            }
          }
          if (c64 + 1 < WSPAN / 64) {  // the next half's TMEM loads fly while this half is stored
            if (e.dbg != 3) {
              tmem_ld_32x32(tbase + (c64 + 1) * 64, ra);
              tmem_ld_32x32(tbase + (c64 + 1) * 64 + 32, rb);
            }
          } else {         // the accumulator is in registers: hand the TMEM buffer back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_even_cta(&tempty[acc]);
          }
          if (elect_one()) tma_store_wait_read<0>();  // the previous box has left the staging tile (bulk groups are per thread: elect_one picks the same lane every time)
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(my_stage_row + ((j ^ (lane & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (e.dbg != 1 && elect_one()) {
            tma_store_2d(&tmC, stage_tile, col_base + c64 * 64, row_base);  // rows / columns past the tensor are clipped
            tma_store_commit();
          }
        }
        if (kRowSums && row_base + lane < M)
          e.row_stats_out[(size_t)(row_base + lane) * e.stats_slots + (col_base >> 7)] = make_float2(st1, st2);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (elect_one()) tma_store_wait_all();  // the staging tile must outlive the last store
      __syncwarp();
    } else if constexpr (EPI == 2) {
      // bias (+ReLU) (+bf16 addend) -> fp32 -> TMA store / reduce-add.  Warp (quarter, half): rows 32 quarter.., columns
      // 128 half.. of the tile as four 32-column boxes (32 fp32 = one 128-byte row of the SWIZZLE_128B staging tile).
      const uint32_t stage_tile = smem_u32(sScr + ew * 4096);
      const uint32_t my_stage_row = stage_tile + lane * 128;
      const bool reduce = e.res_f32 != nullptr;   // the launcher guarantees res == out here
      for (int tile = pair; tile < total; tile += n_pairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int row_base = m_blk * 2 * BM + (int)cta * BM + quarter * 32;
        const int row = row_base + lane;
        const bool row_ok = row < M;
        const int col_base = n_blk * BNT + half * WSPAN;
        if (e.bias) {
          if (lane * 4 < WSPAN) {
            const uint4 b = ldg128_nc(e.bias + col_base + lane * 4);
            sts128(sbias + lane * 16, b.x, b.y, b.z, b.w);
          }
          if constexpr (kLnApply) {
            const uint4 cs = ldg128_nc(e.ln_csum + col_base + lane * 4);
            sts128(sbias + 512 + lane * 16, cs.x, cs.y, cs.z, cs.w);
          }
          warp_sync_smem();
        }
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (kLnApply && row_ok) {   // the row's partial sums, combined in slot order
          const float2* sp = e.ln_stats + (size_t)row * e.ln_slots;
          float s1 = 0.f, s2 = 0.f;
          for (int i = 0; i < e.ln_slots; ++i) { const float2 t = sp[i]; s1 += t.x; s2 += t.y; }
          const float inv = 1.f / (float)e.ln_dim;
          ln_mean = s1 * inv;
          ln_rstd = rsqrtf(fmaxf(s2 * inv - ln_mean * ln_mean, 0.f) + e.ln_eps);
        }
        // this thread's row of the addend for the whole 128-column span: issued before the accumulator is even ready
        uint4 add[WSPAN / 8];
        if (e.add_bf16) {
          const __nv_bfloat16* ap = e.add_bf16 + (size_t)row * e.ld_add + col_base;
#pragma unroll
          for (int g = 0; g < WSPAN / 8; ++g) add[g] = row_ok ? ldg128_nc(ap + 8 * g) : make_uint4(0, 0, 0, 0);
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BNT + half * WSPAN);
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tbase, ra);
#pragma unroll
        for (int c = 0; c < WSPAN / 32; ++c) {
          tmem_ld_wait();
          if (c + 1 < WSPAN / 32) {
            if (c & 1) tmem_ld_32x32(tbase + (c + 1) * 32, ra); else tmem_ld_32x32(tbase + (c + 1) * 32, rb);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_even_cta(&tempty[acc]);
          }
          const uint32_t (&r)[32] = (c & 1) ? rb : ra;
          float v[32];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float v0 = __uint_as_float(r[4 * g]), v1 = __uint_as_float(r[4 * g + 1]);
            float v2 = __uint_as_float(r[4 * g + 2]), v3 = __uint_as_float(r[4 * g + 3]);
            if constexpr (kLnApply) {
              const float4 cs = lds128f(sbias + 512 + (c * 32 + 4 * g) * 4);
              v0 = ln_rstd * (v0 - ln_mean * cs.x); v1 = ln_rstd * (v1 - ln_mean * cs.y);
              v2 = ln_rstd * (v2 - ln_mean * cs.z); v3 = ln_rstd * (v3 - ln_mean * cs.w);
            }
            if (e.bias) {
              const float4 bb = lds128f(sbias + (c * 32 + 4 * g) * 4);
              v0 += bb.x; v1 += bb.y; v2 += bb.z; v3 += bb.w;
            }
            if (e.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
            if (e.add_bf16) {
              const uint4 aa = add[c * 4 + (g >> 1)];
              const uint32_t w0 = (g & 1) ? aa.z : aa.x, w1 = (g & 1) ? aa.w : aa.y;
              const float2 a0 = unpack_h2<F16>(w0), a1 = unpack_h2<F16>(w1);
              v0 += a0.x; v1 += a0.y; v2 += a1.x; v3 += a1.y;
            }
            v[4 * g] = v0; v[4 * g + 1] = v1; v[4 * g + 2] = v2; v[4 * g + 3] = v3;
          }
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128f(my_stage_row + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            if (reduce) tma_reduce_add_2d(&tmC, stage_tile, col_base + c * 32, row_base);
            else tma_store_2d(&tmC, stage_tile, col_base + c * 32, row_base);
            tma_store_commit();
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    } else
    for (int tile = pair; tile < total; tile += n_pairs) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row_base = m_blk * 2 * BM + (int)cta * BM + quarter * 32;  // first of this warp's 32 rows
      const int row = row_base + lane;
      const bool row_ok = row < M;
      float best_v = -INFINITY;
      int best_i = -1;
      {  // pull the NEXT tile's addends towards L2 while this tile is processed
        const int nt = tile + n_pairs;
        if (nt < total && (e.res_f32 || e.add_bf16)) {
          const int nrow = (nt / n_tiles) * 2 * BM + (int)cta * BM + quarter * 32 + lane;
          const int ncol = (nt % n_tiles) * BN + half * 128;
          if (nrow < M && ncol < N) {
            if (e.res_f32) {
              const char* q = reinterpret_cast<const char*>(e.res_f32 + (size_t)nrow * e.ld_res + ncol);
#pragma unroll
              for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 128 * i));
            }
            if (e.add_bf16) {
              const char* q = reinterpret_cast<const char*>(e.add_bf16 + (size_t)nrow * e.ld_add + ncol);
#pragma unroll
              for (int i = 0; i < 2; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 128 * i));
            }
          }
        }
      }
      if (e.bias) {  // this warp's 128 bias values: one coalesced load per tile, read back as smem broadcasts
        const int bc = n_blk * BN + half * 128 + lane * 4;
        uint4 b = make_uint4(0, 0, 0, 0);
        if (bc < N) b = ldg128_nc(e.bias + bc);
        sts128(sbias + lane * 16, b.x, b.y, b.z, b.w);
        warp_sync_smem();
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = n_blk * BN + half * 128 + c * 32;
        if (col0 >= N || e.dbg == 2) break;  // warp-uniform
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * 128 + c * 32);
        tmem_ld_32x32(taddr, r);
        // Addends: coalesced global -> smem (8 or 4 lanes cover one row's 128 / 64 bytes), then each thread
        // reads its own row back.  The TMEM load is in flight meanwhile.
        float4 res[8];
        uint4 add[4];
        if (e.add_bf16) {
          // each thread reads its own row's 64 bytes (two full sectors); issued first so that it is in flight
          // together with the residual loads below (one memory latency per chunk instead of two)
          const __nv_bfloat16* ap = e.add_bf16 + (size_t)row * e.ld_add + col0;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            add[g] = (row_ok && col0 + 8 * g < N) ? ldg128_nc(ap + 8 * g) : make_uint4(0, 0, 0, 0);
        }
        if (e.res_f32) {
          float4 t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
            t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_base + rr < M && col0 + 4 * ch < N)
              t[i] = ldg128f(e.res_f32 + (size_t)(row_base + rr) * e.ld_res + col0 + 4 * ch);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            sts128f(scr + ((lane >> 3) + 4 * i) * SCR_STRIDE + (lane & 7) * 16, t[i].x, t[i].y, t[i].z, t[i].w);
          warp_sync_smem();
#pragma unroll
          for (int g = 0; g < 8; ++g) res[g] = lds128f(my_row + g * 16);
          warp_sync_smem();
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int col = col0 + 4 * g;
          float v0 = __uint_as_float(r[4 * g + 0]), v1 = __uint_as_float(r[4 * g + 1]);
          float v2 = __uint_as_float(r[4 * g + 2]), v3 = __uint_as_float(r[4 * g + 3]);
          if (e.bias) {
            const float4 b = lds128f(sbias + (c * 32 + 4 * g) * 4);
            v0 += b.x; v1 += b.y; v2 += b.z; v3 += b.w;
          }
          if (e.relu == 1) {
            v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
          }
          if (e.add_bf16) {
            const uint32_t w0 = (g & 1) ? add[g >> 1].z : add[g >> 1].x;
            const uint32_t w1 = (g & 1) ? add[g >> 1].w : add[g >> 1].y;
            const float2 a0 = unpack_h2<F16>(w0), a1 = unpack_h2<F16>(w1);
            v0 += a0.x; v1 += a0.y; v2 += a1.x; v3 += a1.y;
          }
          if (e.res_f32) {
            v0 += res[g].x; v1 += res[g].y; v2 += res[g].z; v3 += res[g].w;
          }
          if (e.relu == 2) {
            v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
          }
          v[4 * g] = v0; v[4 * g + 1] = v1; v[4 * g + 2] = v2; v[4 * g + 3] = v3;
          if (e.argmax && row_ok && col < N) {
            if (v0 > best_v) { best_v = v0; best_i = col; }
            if (v1 > best_v) { best_v = v1; best_i = col + 1; }
            if (v2 > best_v) { best_v = v2; best_i = col + 2; }
            if (v3 > best_v) { best_v = v3; best_i = col + 3; }
          }
        }
        if (e.out_f32) {
#pragma unroll
          for (int g = 0; g < 8; ++g) sts128f(my_row + g * 16, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          warp_sync_smem();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
            const float4 o = lds128f(scr + rr * SCR_STRIDE + ch * 16);
            if (row_base + rr < M && col0 + 4 * ch < N && e.dbg != 1)
              stg128f(e.out_f32 + (size_t)(row_base + rr) * e.ld_out_f32 + col0 + 4 * ch, o);
          }
          warp_sync_smem();
        }
        if (e.out_bf16) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            sts128(my_row + g * 16, pack_h2<F16>(v[8 * g], v[8 * g + 1]), pack_h2<F16>(v[8 * g + 2], v[8 * g + 3]),
                   pack_h2<F16>(v[8 * g + 4], v[8 * g + 5]), pack_h2<F16>(v[8 * g + 6], v[8 * g + 7]));
          warp_sync_smem();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (lane >> 2) + 8 * i, ch = lane & 3;
            const uint4 o = lds128(scr + rr * SCR_STRIDE + ch * 16);
            if (row_base + rr < M && col0 + 8 * ch < N && e.dbg != 1)
              stg128(e.out_bf16 + (size_t)(row_base + rr) * e.ld_out_bf16 + col0 + 8 * ch, o);
          }
          warp_sync_smem();
        }
      }
      if (e.argmax && row_ok && best_i >= 0) atomicMax(e.argmax + row, argmax_pack(best_v, best_i));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_even_cta(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs are done with each other's shared memory, barriers and TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// Tensor maps are pure functions of (address, extents, pitch, box, type) and encoding one costs the host 1-2 us; a forward makes
// ~900 of them (three per GEMM, two per attention launch), which is most of the host time of a latency-bound batch-1 call.
// They are cached by exactly those fields; callers keep the set of distinct extents small by rounding row extents up to whole
// tiles inside their buffers' capacity (see gemm_bf16_tcgen05).  An entry never goes stale: it describes addresses, not contents.
namespace {
struct TmapKey {
  const void* base; uint64_t rows, cols, ld; uint32_t box_rows, box_cols; int dt;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols && dt == o.dt;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 1315423911ull + k.ld * 2654435761ull + ((uint64_t)k.box_rows << 32) + ((uint64_t)k.box_cols << 8) + (uint64_t)k.dt);
    return (size_t)(h ^ (h >> 29));
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
}  // namespace

static int make_tmap_sw128(CUtensorMap* out, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols, bool swizzle = true) {
  const TmapKey key{base, rows, cols, ld_elems, box_rows, box_cols, (int)dt | (swizzle ? 0 : 0x100)};
  {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  if (g_tmap_cache.size() > 16384) g_tmap_cache.clear();   // a bound, not a policy: steady-state forwards use a few hundred entries
  g_tmap_cache.emplace(key, *out);
  return 0;
}

int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                         uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_sw128(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld_elems, box_rows, box_cols);
}

int make_tmap_bf16_plain(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                         uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_sw128(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld_elems, box_rows, box_cols, false);
}

// Row extent of a tensor map over a buffer that holds `cap` rows of which `used` matter: whole 256-row tiles, so that a handful
// of extents covers every batch size (cache hits), never more than the buffer holds.  cap <= used (unknown capacity): exact.
static uint64_t tile_extent(int64_t used, int64_t cap) {
  if (cap <= used) return (uint64_t)used;
  const int64_t r = (used + 255) & ~int64_t(255);
  return (uint64_t)(r < cap ? r : cap);
}

int gemm_bf16_tcgen05(const GemmProblem& p, const GemmEpilogue& e, int num_sms, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
  if ((p.N & 3) || (p.lda & 7) || (p.ldw & 7)) return (int)cudaErrorInvalidValue;
  if (e.out_bf16 && ((p.N & 7) || (e.ld_out_bf16 & 7))) return (int)cudaErrorInvalidValue;
  if (e.add_bf16 && ((p.N & 7) || (e.ld_add & 7))) return (int)cudaErrorInvalidValue;
  static PerDeviceOnce once;
  int rc = once_per_device(once, [] {
    cudaError_t err = cudaSuccess;
    auto set = [&](auto kernel) { if (err == cudaSuccess) err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); };
    set(gemm_tcgen05_kernel<0, false>); set(gemm_tcgen05_kernel<1, false>); set(gemm_tcgen05_kernel<2, false>);
    set(gemm_tcgen05_kernel<3, false>); set(gemm_tcgen05_kernel<4, false>);
    set(gemm_tcgen05_kernel<0, true>); set(gemm_tcgen05_kernel<1, true>); set(gemm_tcgen05_kernel<2, true>);
    set(gemm_tcgen05_kernel<3, true>); set(gemm_tcgen05_kernel<4, true>);
    set(gemm_tcgen05_kernel<1, false, 128>); set(gemm_tcgen05_kernel<2, false, 128>);
    set(gemm_tcgen05_kernel<1, true, 128>); set(gemm_tcgen05_kernel<2, true, 128>);
    return (int)err;
  });
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  const int a_cols = p.a_k_wrap > 0 ? p.a_k_wrap : p.K;
  rc = make_tmap_bf16_sw128(&tmA, p.A, tile_extent(p.M, p.rows_a > 0 ? p.rows_a : p.M), (uint64_t)a_cols, (uint64_t)p.lda, BM);
  if (rc) return rc;
  const bool fast = e.out_bf16 && !e.out_f32 && !e.res_f32 && !e.add_bf16 && !e.argmax && e.relu != 2 && (p.N % BN) == 0 && e.dbg != 2 && e.dbg != 4 &&
                    (reinterpret_cast<uintptr_t>(e.out_bf16) & 15) == 0;
  // fp32 output through TMA: residual (if any) must be the output buffer itself, so that "x += tile" is one reduce-add
  static const bool no_f32_tma = getenv("B200PF_NO_F32_TMA") != nullptr;
  const bool f32tma = !fast && !no_f32_tma && e.out_f32 && !e.out_bf16 && !e.argmax && e.relu != 2 && (p.N % BN) == 0 && e.dbg == 0 &&
                      (e.res_f32 == nullptr || (e.res_f32 == e.out_f32 && e.ld_res == e.ld_out_f32)) &&
                      (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0 && (e.ld_out_f32 & 3) == 0;
  if (e.row_stats_out && !fast) return (int)cudaErrorInvalidValue;                       // only the bf16 TMA epilogue writes row sums
  if (e.ln_stats && (!f32tma || !e.bias || !e.ln_csum)) return (int)cudaErrorInvalidValue;  // only the fp32 TMA epilogue applies them
  // Narrow tiles for launch-bound problems: fewer 256-wide tiles than half the CTA pairs, a TMA epilogue without the LayerNorm
  // fold (p.M is the host-side upper bound when the row count lives on the device).  (B200PF_GEMM_NARROW=0 turns it off.)
  static const bool narrow_ok = !(getenv("B200PF_GEMM_NARROW") && atoi(getenv("B200PF_GEMM_NARROW")) == 0);
  const int m_tiles256 = (p.M + 2 * BM - 1) / (2 * BM);
  const bool narrow = narrow_ok && (fast || f32tma) && !e.row_stats_out && !e.ln_stats && (p.N % 128) == 0 &&
                      4 * m_tiles256 * ((p.N + BN - 1) / BN) <= (num_sms / 2);
  const int bn = narrow ? 128 : BN;
  rc = make_tmap_bf16_sw128(&tmB, p.W, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.ldw, bn / 2);
  if (rc) return rc;
  CUtensorMap tmC = tmA;  // placeholder for the general path
  if (fast) {
    rc = make_tmap_bf16_sw128(&tmC, e.out_bf16, tile_extent(p.M, p.rows_c), (uint64_t)p.N, (uint64_t)e.ld_out_bf16, 32, 64);
    if (rc) return rc;
  } else if (f32tma) {
    rc = make_tmap_sw128(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, e.out_f32, tile_extent(p.M, p.rows_c), (uint64_t)p.N, (uint64_t)e.ld_out_f32, 32, 32);
    if (rc) return rc;
  }
  KArgs a;
  a.M = p.M; a.N = p.N; a.K = p.K; a.m_dev = p.m_dev;
  a.a_k_wrap = p.a_k_wrap; a.a_row_shift0 = p.a_row_shift0;
  a.e = e;
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + bn - 1) / bn;
  int grid = 2 * m_tiles * n_tiles;  // CTA pairs
  if (grid > (num_sms & ~1)) grid = num_sms & ~1;
  const dim3 g(grid), b(kThreads);
  if (narrow) {
    if (p.f16) {
      if (fast) return launch_kernel(gemm_tcgen05_kernel<1, true, 128>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
      return launch_kernel(gemm_tcgen05_kernel<2, true, 128>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
    }
    if (fast) return launch_kernel(gemm_tcgen05_kernel<1, false, 128>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
    return launch_kernel(gemm_tcgen05_kernel<2, false, 128>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
  }
  if (p.f16) {
    if (fast && e.row_stats_out) return launch_kernel(gemm_tcgen05_kernel<3, true>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
    if (fast) return launch_kernel(gemm_tcgen05_kernel<1, true>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
    if (f32tma && e.ln_stats) return launch_kernel(gemm_tcgen05_kernel<4, true>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
    if (f32tma) return launch_kernel(gemm_tcgen05_kernel<2, true>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
    return launch_kernel(gemm_tcgen05_kernel<0, true>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
  }
  if (fast && e.row_stats_out) return launch_kernel(gemm_tcgen05_kernel<3, false>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
  if (fast) return launch_kernel(gemm_tcgen05_kernel<1, false>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
  if (f32tma && e.ln_stats) return launch_kernel(gemm_tcgen05_kernel<4, false>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
  if (f32tma) return launch_kernel(gemm_tcgen05_kernel<2, false>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
  return launch_kernel(gemm_tcgen05_kernel<0, false>, g, b, kSmemBytes, stream, tmA, tmB, tmC, a);
}

}  // namespace pf
