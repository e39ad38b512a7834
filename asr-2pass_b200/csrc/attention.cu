// tcgen05 flash attention, head dim 128, variable length, non-causal with key-length masking (see attention.cuh).
//
// PERSISTENT kernel: 2 CTAs per SM, each walking a static share of the work units -- one head of one (segment, 128-query tile)
// item, costliest first, dealt out as a snake.  Inside a CTA four roles run ONE continuous stream of (unit, 64-key block)
// steps, so barrier set-up, the TMEM allocation and the first-load latency are paid once per CTA and the loads of the next
// unit are in flight while the current one finishes.  Warps 0, 1 and 6 run converged, with their TMA / tcgen05 instructions
// under elect_one() (see the note above the role code):
//
//   warp 0      : TMA producer.  Q per head (single buffer); K and V in SEPARATE 2-stage rings: a K stage is free as soon
//                 as S = Q K^T of its block has completed (early), a V stage only after P V -- with one combined ring the
//                 load of block g+1 could not start before P V of block g-1 had finished and the whole pipeline ran serially
//                 (ncu, round 2: softmax warps 29 % of their time waiting for S).  K runs one block ahead of V.
//   warp 1      : S issuer.  S (128 x n x 128, n = keys of the block rounded up to 16) of block g+1 is issued while block g is in
//                 the softmax, across head and item boundaries, so the tensor pipe works under the softmax.  S is double
//                 buffered in TMEM (2 x 64 columns); S of block g+2 waits for the commit of P V of block g, whose P it overwrites.
//   warp 6      : P V issuer (128 x 128 x n into O, 128 TMEM columns).  Its commit frees the V stage (producer), the S / P
//                 buffer (S issuer) and tells the lazy rescale that O is quiescent.
//   warps 2..5  : softmax, one query row per thread: single pass with a running maximum that is only advanced (and O, l
//                 rescaled in TMEM) when the true maximum moved by more than 8 in the log2 domain.  Full 64-key blocks take
//                 a mask-free path (packed FFMA2 / FADD2, 3-input max); only the last block of a segment is masked.
//                 Warps whose 32 query rows lie beyond the segment only keep the barrier protocol going.
//   P (probabilities, 16-bit) never touches shared memory: each softmax thread writes its row back into the TMEM columns its
//   scores came from (tcgen05.st, two values per 32-bit column) and P V reads its A operand from there (tcgen05.mma with A in
//   TMEM).  That takes 32 KB per key block off the shared-memory pipe -- which, not the tensor pipe, bounds this kernel (Q is
//   re-read for every 64-key S tile) -- and replaces the generic->async proxy fence by a tcgen05 fence.  O / l is applied once
//   per head.
#include "attention.cuh"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace pf {
namespace {

constexpr int BQ = 128, BKV = 64, HD = 128, KV_STAGES = 2;
constexpr int Q_BYTES = BQ * HD * 2;       // 32768: two 128x64 swizzled boxes
constexpr int K_BYTES = BKV * HD * 2;      // 16384: two 64x64 boxes
constexpr int kThreads = 224;             // warp 0 TMA, warp 1 S issuer, warps 2-5 softmax, warp 6 P V issuer
constexpr int kTmemCols = 256;
// 96 KB + barriers: two CTAs fit one SM (2 x 256 TMEM columns), so one CTA's softmax is covered by the other's MMAs.
// + this CTA's unit table (below): every role reads its units from shared memory instead of chasing work[] -> q_len[] ->
// row_off[] through global memory at each unit boundary (seven role streams x ~1.5 us of dependent loads per unit, for
// units that last 3-4 key blocks on the benchmark's segment lengths).
constexpr int kMaxUnits = 224;             // one table entry per thread of the CTA; later rounds fall back to global loads
constexpr int kSmemBytes = Q_BYTES + 2 * KV_STAGES * K_BYTES + 256 + kMaxUnits * 32;

struct AArgs {
  const int* q_row_off;
  const int* q_len;
  const int* kv_row_off;
  const int* kv_len;
  const AttnWork* work;
  __nv_bfloat16* out;
  int ldo;
  int q_col0, k_col0, v_col0;
  int n_work, n_heads;
  float scale_log2e;
  int dbg;   // micro-benchmark ablations only (B200PF_ATTN_DBG): 1 = no exponentials; 0 in the product
  long long* trace;   // ablation 64: CTA 0 writes clock64() stamps of its first 64 key blocks here (4 per block)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)),
        "l"(reinterpret_cast<const unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}

// The (work unit, key block) stream of one CTA; a unit is one head of one (segment, query tile) item.  All three roles walk it
// with their own copy; `g` counts key blocks and `hc` units since the CTA started -- every ring stage and barrier phase is
// derived from them.
struct UnitRec { int q0, Tq, Tk, q_row, kv_row, h, valid, pad; };   // 32 bytes

struct Stream {
  const AArgs& a;
  const UnitRec* table;
  int cta, n_cta, round, n_units;
  int q0, Tq, Tk, nb, q_row, kv_row;
  int h, j, g, hc;
  bool valid;
  __device__ Stream(const AArgs& args, const UnitRec* table_, int cta_, int n_cta_)
      : a(args), table(table_), cta(cta_), n_cta(n_cta_), round(0), j(0), g(0), hc(0) {
    n_units = a.n_work * a.n_heads;
    valid = load_unit();
  }
  // Round r of CTA `cta`: the unit's record, straight from global memory (table fill and rounds beyond the table).
  static __device__ UnitRec fetch(const AArgs& a, int n_units, int n_cta, int cta, int round) {
    UnitRec u;
    u.valid = 0;
    const int unit = round * n_cta + ((round & 1) ? n_cta - 1 - cta : cta);
    if (round * n_cta >= n_units || unit >= n_units) return u;
    const AttnWork w = a.work[unit / a.n_heads];
    u.Tq = a.q_len[w.seg];
    u.Tk = a.kv_len[w.seg];
    if (w.q0 < u.Tq && u.Tk > 0) {
      u.valid = 1;
      u.h = unit % a.n_heads;
      u.q0 = w.q0;
      u.q_row = a.q_row_off[w.seg] + w.q0;
      u.kv_row = a.kv_row_off[w.seg];
    }
    return u;
  }
  // Items are sorted by cost (longest segment first) and the heads of an item are consecutive units; round r hands unit
  // r * n_cta + (cta or n_cta - 1 - cta) to this CTA -- a snake, so no CTA gets the longer unit of every round.
  __device__ bool load_unit() {
    for (;; ++round) {
      if (round * n_cta >= n_units) return false;
      const UnitRec u = round < kMaxUnits ? table[round] : fetch(a, n_units, n_cta, cta, round);
      if (u.valid) {
        h = u.h; q0 = u.q0; Tq = u.Tq; Tk = u.Tk; q_row = u.q_row; kv_row = u.kv_row;
        nb = (Tk + BKV - 1) / BKV;
        return true;
      }
    }
  }
  __device__ void advance() {
    ++g;
    if (++j == nb) {
      j = 0;
      ++hc;
      ++round;
      valid = load_unit();
    }
  }
  __device__ int keys() const { const int n = Tk - j * BKV; return n < BKV ? n : BKV; }    // valid keys of this block
  __device__ int keys16() const { return (keys() + 15) & ~15; }                            // MMA extent of this block
};

// DBG: the ablation / trace switches (B200PF_ATTN_DBG) exist only in a second instantiation; the product kernel carries none of
// their branches.
template <bool F16, bool DBG>
__global__ void __launch_bounds__(kThreads, 2)
attn_heads_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AArgs a) {
  const int dbg = DBG ? a.dbg : 0;
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B tiles need 1024-byte alignment (checked below)
  uint8_t* sQ = smem;
  uint8_t* sK = smem + Q_BYTES;
  uint8_t* sV = sK + KV_STAGES * K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * K_BYTES);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* q_full = bars;                  // 1
  uint64_t* q_empty = bars + 1;             // 1
  uint64_t* k_full = bars + 2;              // 2
  uint64_t* v_full = bars + 4;              // 2
  uint64_t* s_full = bars + 6;              // 2: S of block g has completed -- for the softmax warps AND for the producer (the K
                                            // stage of block g, same parity, may be reloaded)
  uint64_t* p_full = bars + 8;              // 2: one per S / P buffer.  A softmax warp may run one block ahead of a slower one (S of
                                            // block g + 1 is ready early); it then arrives on the OTHER barrier.  Two blocks ahead is
                                            // impossible: S of block g + 2 needs P V of block g, which needs all four arrivals.
  uint64_t* pv_done = bars + 10;            // 2: P V of block g has completed: its V stage may be reloaded (producer), its S / P
                                            // buffer may take S of block g + 2 (S issuer), O is quiescent (lazy rescale)
  uint64_t* o_full = bars + 12;             // 1
  uint64_t* o_empty = bars + 13;            // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  UnitRec* units = reinterpret_cast<UnitRec*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&pv_done[i], 1); }
    mbar_init(&p_full[0], 4); mbar_init(&p_full[1], 4);
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
  pdl_wait();   // q_len of the decoder is produced by the CIF kernels; nothing above touches activations
  pdl_launch_dependents();
  {
    static_assert(kMaxUnits == kThreads, "one table entry per thread");
    const int n_units = a.n_work * a.n_heads;
    const UnitRec u = Stream::fetch(a, n_units, gridDim.x, blockIdx.x, threadIdx.x);
    units[threadIdx.x] = u;
  }
  __syncthreads();

  // Warps 0 and 1 run CONVERGED: all 32 lanes walk the stream and wait on the barriers; only the TMA / tcgen05 instructions
  // sit under elect_one().  Inside an `if (lane == 0)` region the compiler has to wrap every UTCHMMA / UTCBAR / UTMALDG in a
  // "waterfall" loop (ELECT + branch, operands through per-thread registers): ~75 cycles per instruction, and with 12 MMAs and 4
  // commits per 64-key block the issuing thread alone took ~2300 cycles per block (clock64 trace, tools/trace_attn.py) -- more
  // than the tensor pipe (512) or the softmax (~1000) need.
  if (warp == 0) {
    Stream st(a, units, blockIdx.x, gridDim.x);
    int pend_r = 0, pend_c = 0, pend_g = -1;   // V load of the previous block: K runs one block ahead of V
    auto load_v = [&]() {
      const int stage = pend_g % KV_STAGES;
      mbar_wait(&pv_done[stage], (uint32_t)(((pend_g / KV_STAGES) & 1) ^ 1));   // P V of block pend_g - 2 has read this stage
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[stage], K_BYTES);
        uint8_t* dst = sV + stage * K_BYTES;
        tma_load_2d(dst, &tmKV, &v_full[stage], pend_c, pend_r);
        tma_load_2d(dst + K_BYTES / 2, &tmKV, &v_full[stage], pend_c + 64, pend_r);
      }
      __syncwarp();
    };
    while (st.valid) {
      if (st.j == 0) {
        mbar_wait(q_empty, (uint32_t)((st.hc & 1) ^ 1));
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, Q_BYTES);
          tma_load_2d(sQ, &tmQ, q_full, a.q_col0 + st.h * HD, st.q_row);
          tma_load_2d(sQ + Q_BYTES / 2, &tmQ, q_full, a.q_col0 + st.h * HD + 64, st.q_row);
        }
        __syncwarp();
      }
      const int stage = st.g % KV_STAGES;
      const int r = st.kv_row + st.j * BKV;
      mbar_wait(&s_full[stage], (uint32_t)(((st.g / KV_STAGES) & 1) ^ 1));      // S of block g - 2 has read this stage
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[stage], K_BYTES);
        uint8_t* dst = sK + stage * K_BYTES;
        tma_load_2d(dst, &tmKV, &k_full[stage], a.k_col0 + st.h * HD, r);
        tma_load_2d(dst + K_BYTES / 2, &tmKV, &k_full[stage], a.k_col0 + st.h * HD + 64, r);
      }
      __syncwarp();
      if (pend_g >= 0) load_v();
      pend_r = r; pend_c = a.v_col0 + st.h * HD; pend_g = st.g;
      st.advance();
    }
    if (pend_g >= 0) load_v();
  } else if (warp == 1) {
    // ---- S issuer.  The MMA role is split over two warps: one warp doing both spent ~1000 cycles per key block on its own
    // barrier waits and commits (ablation: with almost no MMA work left the block period only fell from 1723 to ~1000 cycles),
    // more than the tensor pipe or the softmax need.  Each warp now has two waits, its MMAs and one commit per block.
    constexpr uint32_t idesc_s0 = umma_idesc_h16(F16, BQ, 0);           // N filled in per block
    const uint32_t q_addr = smem_u32(sQ);
    const bool tracing = DBG && (dbg & 64) && blockIdx.x == 0 && lane == 0;
    for (Stream s_it(a, units, blockIdx.x, gridDim.x); s_it.valid; s_it.advance()) {
      const int g = s_it.g, stage = g % KV_STAGES, sb = g & 1;
      if (tracing && g < 64) a.trace[8 * g + 4] = clock64();   // S issuer reaches block g
      if (s_it.j == 0) mbar_wait(q_full, (uint32_t)(s_it.hc & 1));
      mbar_wait(&k_full[stage], (uint32_t)((g / KV_STAGES) & 1));
      // S buffer sb still holds P of block g - 2, the A operand of that block's P V, which another warp issues: wait for its commit
      mbar_wait(&pv_done[sb], (uint32_t)(((g >> 1) & 1) ^ 1));
      if (tracing && g < 64) a.trace[8 * g + 5] = clock64();   // K of block g is there and the S buffer is free
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + stage * K_BYTES);
      const uint32_t idesc = idesc_s0 | ((uint32_t)(s_it.keys16() >> 3) << 17);
      const bool last = s_it.j + 1 == s_it.nb;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          if ((dbg & 4) && ks > 0) break;
          const uint64_t da = umma_desc_sw128(q_addr + (ks >> 2) * (Q_BYTES / 2)) + 2 * (ks & 3);
          const uint64_t db = umma_desc_sw128(k_addr + (ks >> 2) * (K_BYTES / 2)) + 2 * (ks & 3);
          umma_bf16(tmem_base + sb * BKV, da, db, idesc, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
        if (last) umma_commit(q_empty);   // Q may be replaced once every S MMA of this head has read it
      }
      __syncwarp();
      if (tracing && g < 64) a.trace[8 * g + 0] = clock64();   // S of block g committed
    }
  } else if (warp == 6) {
    // ---- P V issuer
    constexpr uint32_t idesc_pv = umma_idesc_h16(F16, BQ, HD, 0, 1);    // B (= V) is MN-major
    const bool tracing = DBG && (dbg & 64) && blockIdx.x == 0 && lane == 0;
    for (Stream pv_it(a, units, blockIdx.x, gridDim.x); pv_it.valid; pv_it.advance()) {
      const int g = pv_it.g, stage = g % KV_STAGES;
      mbar_wait(&p_full[g & 1], (uint32_t)((g >> 1) & 1));
      if (tracing && g < 64) a.trace[8 * g + 3] = clock64();         // P V issuer saw P of block g
      if (pv_it.j == 0) mbar_wait(o_empty, (uint32_t)((pv_it.hc & 1) ^ 1));   // the previous head's O has been read out
      mbar_wait(&v_full[stage], (uint32_t)((g / KV_STAGES) & 1));
      if (tracing && g < 64) a.trace[8 * g + 6] = clock64();   // V of block g is there
      tc_fence_after();
      const uint32_t p_tmem = tmem_base + (uint32_t)((g & 1) * BKV);   // P sits where S of this block was: 8 columns per 16 keys
      const uint32_t v_addr = smem_u32(sV + stage * K_BYTES);
      const int ksteps = (dbg & 8) ? 0 : pv_it.keys16() >> 4;
      const bool first = pv_it.j == 0, last = pv_it.j + 1 == pv_it.nb;
      if (elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) {
          // V tile: 64 keys x 128 d as two [64 x 64] boxes 8 KB apart; 16 keys = 2048 B per K step
          const uint64_t db = umma_desc_sw128_mn(v_addr + ks * 2048, K_BYTES / 2);
          umma_bf16_ts(tmem_o, p_tmem + 8 * ks, db, idesc_pv, (!first || ks != 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[g & 1]);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
      if (tracing && g < 64) a.trace[8 * g + 7] = clock64();   // P V of block g issued and committed
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float2 sc2 = make_float2(a.scale_log2e, a.scale_log2e);
    uint32_t r[32], r2[32];
    float m = -INFINITY, l = 0.f, m_used = 0.f;
    for (Stream st(a, units, blockIdx.x, gridDim.x); st.valid; st.advance()) {
      const int g = st.g, sb = g & 1;
      const bool warp_active = st.q0 + quarter * 32 < st.Tq && !(dbg & 16);   // warp-uniform: does this warp own any real query row?
      if (st.j == 0) { m = -INFINITY; l = 0.f; m_used = 0.f; }
      mbar_wait(&s_full[sb], (uint32_t)((g >> 1) & 1));
      if (DBG && (dbg & 64) && blockIdx.x == 0 && g < 64 && threadIdx.x == 64) a.trace[8 * g + 1] = clock64();   // softmax warp saw S of block g
      tc_fence_after();
      const int nvalid = st.keys();
      bool need = false;
      float factor = 1.f;
      uint32_t pk[32];
      if (warp_active) {
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV, r);
        tmem_ld_32x32(tmem_base + lane_addr + sb * BKV + 32, r2);
        tmem_ld_wait();   // the scores are in registers; their TMEM columns will take P below
        if (nvalid == BKV) {
          // ---- full block: no masks ----
          float mb0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r2[0]));
          float mb1 = fmaxf(__uint_as_float(r[1]), __uint_as_float(r2[1]));
#pragma unroll
          for (int c = 2; c < 32; c += 2) {
            mb0 = max3(mb0, __uint_as_float(r[c]), __uint_as_float(r2[c]));
            mb1 = max3(mb1, __uint_as_float(r[c + 1]), __uint_as_float(r2[c + 1]));
          }
          m = max3(m, mb0, mb1);
          if (st.j == 0) {
            m_used = m;
          } else if ((m - m_used) * a.scale_log2e > 8.f) {
            need = true;
            factor = ex2((m_used - m) * a.scale_log2e);
            m_used = m;
            l *= factor;
          }
          const float mc = -m_used * a.scale_log2e;
          const float2 mc2 = make_float2(mc, mc);
          float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float2 x = ffma2(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), sc2, mc2);
            const float2 y = ffma2(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), sc2, mc2);
            const float2 p = dbg & 1 ? x : make_float2(ex2(x.x), ex2(x.y));
            const float2 q = dbg & 1 ? y : make_float2(ex2(y.x), ex2(y.y));
            la = fadd2(la, p);
            lb = fadd2(lb, q);
            pk[c >> 1] = pack_h2<F16>(p.x, p.y);
            pk[16 + (c >> 1)] = pack_h2<F16>(q.x, q.y);
          }
          l += (la.x + la.y) + (lb.x + lb.y);
        } else {
          // ---- last block of the segment: keys >= nvalid are masked ----
          float mb = -INFINITY;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c < nvalid) mb = fmaxf(mb, __uint_as_float(r[c]));
            if (32 + c < nvalid) mb = fmaxf(mb, __uint_as_float(r2[c]));
          }
          m = fmaxf(m, mb);
          if (st.j == 0) {
            m_used = m;
          } else if ((m - m_used) * a.scale_log2e > 8.f) {
            need = true;
            factor = ex2((m_used - m) * a.scale_log2e);
            m_used = m;
            l *= factor;
          }
          const float mc = m_used * a.scale_log2e;
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float p0 = (c < nvalid) ? ex2(fmaf(__uint_as_float(r[c]), a.scale_log2e, -mc)) : 0.f;
            const float p1 = (c + 1 < nvalid) ? ex2(fmaf(__uint_as_float(r[c + 1]), a.scale_log2e, -mc)) : 0.f;
            const float p2 = (32 + c < nvalid) ? ex2(fmaf(__uint_as_float(r2[c]), a.scale_log2e, -mc)) : 0.f;
            const float p3 = (33 + c < nvalid) ? ex2(fmaf(__uint_as_float(r2[c + 1]), a.scale_log2e, -mc)) : 0.f;
            l += (p0 + p1) + (p2 + p3);
            pk[c >> 1] = pack_h2<F16>(p0, p1);
            pk[16 + (c >> 1)] = pack_h2<F16>(p2, p3);
          }
        }
      }
      if (warp_active) {
        if (__any_sync(0xffffffffu, need)) {
          // rare: this warp's 32 rows of O are rescaled in TMEM (rows that did not move use factor 1).  O is quiescent once
          // P V of the previous block has completed (pv_done of block g - 1); `need` is never set on a head's first block.
          mbar_wait(&pv_done[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1));
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
        // P row -> the TMEM columns S came from: column c of the block's buffer holds keys 2c, 2c + 1 (columns past the block's
        // MMA extent are written too -- zeros -- and never read)
        {
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) { lo[c] = pk[c]; hi[c] = pk[16 + c]; }
          tmem_st_32x16(tmem_base + lane_addr + sb * BKV, lo);
          tmem_st_32x16(tmem_base + lane_addr + sb * BKV + 16, hi);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (DBG && (dbg & 64) && blockIdx.x == 0 && g < 64 && threadIdx.x == 64) a.trace[8 * g + 2] = clock64();   // softmax warp arrives for block g
      if (lane == 0) mbar_arrive(&p_full[sb]);
      if (st.j + 1 == st.nb) {
        // ---- this head's output ----
        mbar_wait(o_full, (uint32_t)(st.hc & 1));
        tc_fence_after();
        const bool row_ok = (st.q0 + row) < st.Tq;
        if (warp_active) {
          const float inv = 1.f / l;
          __nv_bfloat16* orow = a.out + (size_t)(st.q_row + row) * a.ldo + st.h * HD;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            tmem_ld_32x32(tmem_o + lane_addr + c2 * 64, r);
            tmem_ld_32x32(tmem_o + lane_addr + c2 * 64 + 32, r2);
            tmem_ld_wait();
            if (c2 == 1) {   // O is in registers: the next head's first P V may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(o_empty);
            }
            if (row_ok) {
#pragma unroll
              for (int gq = 0; gq < 8; ++gq) {
                const uint32_t (&rr)[32] = gq < 4 ? r : r2;
                const int b = 8 * (gq & 3);
                uint4 o;
                o.x = pack_h2<F16>(__uint_as_float(rr[b + 0]) * inv, __uint_as_float(rr[b + 1]) * inv);
                o.y = pack_h2<F16>(__uint_as_float(rr[b + 2]) * inv, __uint_as_float(rr[b + 3]) * inv);
                o.z = pack_h2<F16>(__uint_as_float(rr[b + 4]) * inv, __uint_as_float(rr[b + 5]) * inv);
                o.w = pack_h2<F16>(__uint_as_float(rr[b + 6]) * inv, __uint_as_float(rr[b + 7]) * inv);
                stg128(orow + c2 * 64 + gq * 8, o);
              }
            }
          }
        } else {
          if (lane == 0) mbar_arrive(o_empty);
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
template <bool F16>
__global__ void attn_check_kernel(AArgs a, const __nv_bfloat16* q, int ldq, const __nv_bfloat16* kv, int ldkv, float scale) {
  pdl_wait();
  pdl_launch_dependents();
  const AttnWork w = a.work[blockIdx.x];
  const int h = blockIdx.y;
  const int Tq = a.q_len[w.seg], Tk = a.kv_len[w.seg];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto ld4 = [](const __nv_bfloat16* p, float (&o)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a0 = unpack_h2<F16>(u.x), a1 = unpack_h2<F16>(u.y);
    o[0] = a0.x; o[1] = a0.y; o[2] = a1.x; o[3] = a1.y;
  };
  for (int rr = warp; rr < BQ; rr += blockDim.x >> 5) {
    const int qi = w.q0 + rr;
    if (qi >= Tq) break;
    float qv[4], o[4] = {0, 0, 0, 0};
    ld4(q + (size_t)(a.q_row_off[w.seg] + qi) * ldq + a.q_col0 + h * HD + lane * 4, qv);
    float m = -INFINITY, l = 0.f;
    for (int k = 0; k < Tk; ++k) {
      float kf[4], vf[4];
      ld4(kv + (size_t)(a.kv_row_off[w.seg] + k) * ldkv + a.k_col0 + h * HD + lane * 4, kf);
      ld4(kv + (size_t)(a.kv_row_off[w.seg] + k) * ldkv + a.v_col0 + h * HD + lane * 4, vf);
      float s = 0.f;
      for (int d = 0; d < 4; ++d) s += qv[d] * kf[d];
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      s *= scale;
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      l = l * corr + p;
      for (int d = 0; d < 4; ++d) o[d] = o[d] * corr + p * vf[d];
      m = mn;
    }
    uint2 pk;
    pk.x = pack_h2<F16>(o[0] / l, o[1] / l);
    pk.y = pack_h2<F16>(o[2] / l, o[3] / l);
    *reinterpret_cast<uint2*>(a.out + (size_t)(a.q_row_off[w.seg] + qi) * a.ldo + h * HD + lane * 4) = pk;
  }
}

AArgs make_args(const AttnProblem& p) {
  AArgs a;
  a.q_row_off = p.q_row_off; a.q_len = p.q_len; a.kv_row_off = p.kv_row_off; a.kv_len = p.kv_len;
  a.work = p.work; a.out = p.out; a.ldo = p.ldo;
  a.q_col0 = p.q_col0; a.k_col0 = p.k_col0; a.v_col0 = p.v_col0;
  a.n_work = p.n_work; a.n_heads = p.n_heads;
  a.scale_log2e = p.scale * 1.4426950408889634f;
  static const int dbg = getenv("B200PF_ATTN_DBG") ? atoi(getenv("B200PF_ATTN_DBG")) : 0;
  a.dbg = dbg;
  a.trace = nullptr;
  if (dbg & 64) {
    static long long* buf = nullptr;
    if (!buf) { cudaMalloc((void**)&buf, 64 * 8 * sizeof(long long)); cudaMemset(buf, 0, 64 * 8 * sizeof(long long)); }
    a.trace = buf;
  }
  return a;
}

}  // namespace

int attention_tcgen05(const AttnProblem& p, cudaStream_t stream) {
  if (p.n_work <= 0) return 0;
  static PerDeviceOnce once;
  int rc = once_per_device(once, [] {
    cudaError_t err = cudaSuccess;
    auto set = [&](auto kernel) { if (err == cudaSuccess) err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes); };
    set(attn_heads_kernel<false, false>); set(attn_heads_kernel<true, false>);
    set(attn_heads_kernel<false, true>); set(attn_heads_kernel<true, true>);
    return (int)err;
  });
  if (rc) return rc;
  CUtensorMap tmQ, tmKV;
  rc = make_tmap_bf16_sw128(&tmQ, p.q, (uint64_t)p.q_rows, (uint64_t)p.ldq, (uint64_t)p.ldq, BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_sw128(&tmKV, p.kv, (uint64_t)p.kv_rows, (uint64_t)p.ldkv, (uint64_t)p.ldkv, BKV);
  if (rc) return rc;
  const AArgs a = make_args(p);
  static const int ctas_per_sm = getenv("B200PF_ATTN_CTAS") ? atoi(getenv("B200PF_ATTN_CTAS")) : 2;
  const int resident = ctas_per_sm * (p.num_sms > 0 ? p.num_sms : 148), units = p.n_work * p.n_heads;
  const dim3 grid(units < resident ? units : resident);
  if (a.dbg) {   // micro-benchmark ablations / trace only
    if (p.f16) return launch_kernel(attn_heads_kernel<true, true>, grid, dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a);
    return launch_kernel(attn_heads_kernel<false, true>, grid, dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a);
  }
  if (p.f16) return launch_kernel(attn_heads_kernel<true, false>, grid, dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a);
  return launch_kernel(attn_heads_kernel<false, false>, grid, dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, a);
}

// ablation 64: the stamps of the last launch (host copy), for tools/bench_attn.py
extern "C" int b200pf_attn_trace_read(long long* out) {
  AttnProblem dummy;
  const AArgs a = make_args(dummy);
  if (!a.trace) return -1;
  cudaDeviceSynchronize();
  return (int)cudaMemcpy(out, a.trace, 64 * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
}

int attention_check_kernel(const AttnProblem& p, cudaStream_t stream) {
  if (p.n_work <= 0) return 0;
  const AArgs a = make_args(p);
  dim3 grid(p.n_work, p.n_heads);
  if (p.f16) return launch_kernel(attn_check_kernel<true>, grid, dim3(256), 0, stream, a, p.q, p.ldq, p.kv, p.ldkv, p.scale);
  return launch_kernel(attn_check_kernel<false>, grid, dim3(256), 0, stream, a, p.q, p.ldq, p.kv, p.ldkv, p.scale);
}

}  // namespace pf
