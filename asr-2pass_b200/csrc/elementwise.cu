// K3 LayerNorm, K6 FSMN memory block, K7/K8 CIF predictor tail, K11 argmax decode.  See kernels.cuh.
#include "gemm.cuh"
#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace pf {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (two-pass mean / variance in fp32).
// ------------------------------------------------------------------------------------------------
template <int D, bool IN_BF16, bool F16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ in, int rows, const int* __restrict__ rows_dev, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out_bf16,
                 float* __restrict__ out_f32, const int2* __restrict__ row_info, int zero_gap) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int VW = IN_BF16 ? 8 : 4;            // elements per vector load
  constexpr int NVEC = D / VW;                   // vectors per row
  constexpr int PER = (NVEC + 31) / 32;          // vectors per lane
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nrows = rows_dev ? *rows_dev : rows;
  if (row >= nrows) return;
  const bool gap = zero_gap && row_info && row_info[row].x < 0;

  float v[PER * VW];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int vi = lane + 32 * i;
    if (vi < NVEC) {
      if (IN_BF16) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(in) + (size_t)row * D + vi * 8);
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = unpack_h2<F16>(uw[k]); v[i * 8 + 2 * k] = f.x; v[i * 8 + 2 * k + 1] = f.y; }
      } else {
        const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + (size_t)row * D + vi * 4);
        v[i * 4] = f.x; v[i * 4 + 1] = f.y; v[i * 4 + 2] = f.z; v[i * 4 + 3] = f.w;
      }
#pragma unroll
      for (int k = 0; k < VW; ++k) sum += v[i * VW + k];
    } else {
#pragma unroll
      for (int k = 0; k < VW; ++k) v[i * VW + k] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    if (lane + 32 * i < NVEC) {
#pragma unroll
      for (int k = 0; k < VW; ++k) { const float d = v[i * VW + k] - mean; sq += d * d; }
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int vi = lane + 32 * i;
    if (vi < NVEC) {
      const int c = vi * VW;
      float o[VW];
#pragma unroll
      for (int k = 0; k < VW; k += 4) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + c + k);
        const float4 b = *reinterpret_cast<const float4*>(beta + c + k);
        o[k] = (v[i * VW + k] - mean) * rstd * g.x + b.x;
        o[k + 1] = (v[i * VW + k + 1] - mean) * rstd * g.y + b.y;
        o[k + 2] = (v[i * VW + k + 2] - mean) * rstd * g.z + b.z;
        o[k + 3] = (v[i * VW + k + 3] - mean) * rstd * g.w + b.w;
      }
      if (gap) {
#pragma unroll
        for (int k = 0; k < VW; ++k) o[k] = 0.f;
      }
      if (out_f32) {
#pragma unroll
        for (int k = 0; k < VW; k += 4)
          *reinterpret_cast<float4*>(out_f32 + (size_t)row * D + c + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
      }
      if (out_bf16) {
        uint32_t pk[VW / 2];
#pragma unroll
        for (int k = 0; k < VW / 2; ++k) pk[k] = pack_h2<F16>(o[2 * k], o[2 * k + 1]);
        if (VW == 8)
          *reinterpret_cast<uint4*>(out_bf16 + (size_t)row * D + c) = make_uint4(pk[0], pk[1], pk[VW / 2 - 2], pk[VW / 2 - 1]);
        else
          *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * D + c) = make_uint2(pk[0], pk[1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// FSMN: depthwise conv k=11 along time inside a segment + identity.
// One thread owns 2 channels and walks a run of FSMN_RUN consecutive rows: every input row is read ONCE
// and scattered into a ring of 11 open output accumulators (registers), so the kernel streams
// (RUN+10)/RUN rows per output row instead of 11.
// ------------------------------------------------------------------------------------------------
constexpr int FSMN_RUN = 22;  // multiple of 11 keeps the ring indices static
constexpr int FSMN_THREADS = 256;

// packed fp32x2 FMA (sm_100): two FMAs per issued instruction
__device__ __forceinline__ void ffma2(float2& acc, const float2& a, const float2& b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(acc))
      : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}
// y[0..1] += v, fire and forget (the add happens in L2; every element is touched by exactly one thread, once per launch)
__device__ __forceinline__ void red_add_f32x2(float* p, float2 v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// FAST: the whole window of the run (rows r0 - 5 .. r0 + RUN + 4) lies inside ONE segment, so every tap of every output exists and
// no per-row bookkeeping is needed -- loads, conversions, 11 packed FMAs per row and the store.  About four out of five runs of the
// benchmark workload take this path (segments are 33-333 rows, a window is 32); the general path handles segment edges, gap rows
// and the batch tail.  Both paths apply the same FMAs in the same order to an output, so results do not depend on which one a row
// falls into (batch invariance).
// `tile` points at this thread's 2 channels of window row 0; consecutive rows are FSMN_TILE_LD words apart.
constexpr int FSMN_TILE_LD = 128;
template <bool FAST, bool F16>
__device__ __forceinline__ void fsmn_run(const uint32_t* __restrict__ tile, const float2 (&w)[11],
                                         const int2* __restrict__ row_info, int nrows, int r0, int mode,
                                         __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ y_f32) {
  const int c = threadIdx.x * 2;
  float2 acc[11];
#pragma unroll
  for (int s = 0; s < 11; ++s) acc[s] = make_float2(0.f, 0.f);

  // input i (row r0 - 5 + i) feeds outputs o = i - 5 - d, d = -5..5, with tap j = d + 5; output o is
  // complete after input i = o + 10.  Everything below is fully unrolled, so o and the ring slot are static.
  constexpr int NSTEP = (FSMN_RUN + 10 + 10) / 11;
  int2 info[2][11];
  auto load_step = [&](int step, int2 (&inf)[11]) {   // general path only: the rows' (frame, length) pairs, one step ahead
    if constexpr (!FAST) {
#pragma unroll
      for (int ii = 0; ii < 11; ++ii) {
        const int i = step * 11 + ii;
        const int rin = r0 - 5 + i;
        const bool in_range = rin >= 0 && rin < nrows && i < FSMN_RUN + 10;
        inf[ii] = in_range ? row_info[rin] : make_int2(-1, 0);
      }
    }
  };
  load_step(0, info[0]);
  unsigned prev_valid = 0;  // bit ii: input ii of the previous step was a real frame row
#pragma unroll
  for (int step = 0; step < NSTEP; ++step) {
    const int cur = step & 1;
    if (step + 1 < NSTEP) load_step(step + 1, info[cur ^ 1]);  // in flight while this step computes
    unsigned cur_valid = 0;
#pragma unroll
    for (int ii = 0; ii < 11; ++ii) {
      const int i = step * 11 + ii;
      if (FAST) {
        if (i < FSMN_RUN + 10) {
          const float2 x = unpack_h2<F16>(tile[i * FSMN_TILE_LD]);
#pragma unroll
          for (int d = -5; d <= 5; ++d) {
            const int o = i - 5 - d;
            if (o >= 0 && o < FSMN_RUN) {
              const int slot = ((ii - 5 - d) % 11 + 11) % 11;
              ffma2(acc[slot], w[d + 5], x);
            }
          }
        }
      } else {
        const int2 inf = info[cur][ii];
        if (inf.x >= 0) {  // gap rows and rows outside the batch contribute nothing
          const float2 x = unpack_h2<F16>(tile[(i < FSMN_RUN + 10 ? i : 0) * FSMN_TILE_LD]);
          cur_valid |= 1u << ii;
          if (inf.x >= 5 && inf.x + 5 < inf.y) {
            // interior frame: all 11 neighbours are in the same segment
#pragma unroll
            for (int d = -5; d <= 5; ++d) {
              const int o = i - 5 - d;
              if (o >= 0 && o < FSMN_RUN) {  // static: outputs of other runs are not accumulated here
                const int slot = ((ii - 5 - d) % 11 + 11) % 11;
                ffma2(acc[slot], w[d + 5], x);
              }
            }
          } else {
#pragma unroll
            for (int d = -5; d <= 5; ++d) {
              const int o = i - 5 - d;
              if (o >= 0 && o < FSMN_RUN) {
                // output row rin - d lies in the same segment iff its frame index t - d is inside [0, T)
                const bool ok = (inf.x - d) >= 0 && (inf.x - d) < inf.y;
                const int slot = ((ii - 5 - d) % 11 + 11) % 11;
                if (ok) ffma2(acc[slot], w[d + 5], x);
              }
            }
          }
        }
      }
      // output o = i - 10 (row rin - 5) has now seen all of its inputs; it was input i - 5
      const int o = i - 10;
      const int slot_done = (ii + 1) % 11;
      if (o >= 0 && o < FSMN_RUN) {
        if (FAST || r0 + o < nrows) {
          const int rout = r0 + o;
          const bool out_valid = FAST ? true : ((ii >= 5) ? ((cur_valid >> (ii - 5)) & 1u) : ((prev_valid >> (ii + 6)) & 1u));
          if (mode == 0) {
            const uint32_t pk = out_valid ? pack_h2<F16>(acc[slot_done].x, acc[slot_done].y) : 0u;
            *reinterpret_cast<uint32_t*>(out_bf16 + (size_t)rout * 512 + c) = pk;
          } else if (out_valid) {
            red_add_f32x2(y_f32 + (size_t)rout * 512 + c, acc[slot_done]);
          }
        }
        acc[slot_done] = make_float2(0.f, 0.f);
      }
    }
    prev_valid = cur_valid;
  }
}

// PERSISTENT: 2 CTAs per SM, each walking windows blockIdx.x, blockIdx.x + gridDim.x, ... (neighbouring CTAs work on
// neighbouring rows at the same time, so the 10 halo rows two windows share are L2 hits).  A window (RUN + 10 rows x 512
// channels, 32 KB) is fetched by ONE thread with two TMA box loads (rows before the batch / beyond the tensor map arrive as
// zeros) into a 3-stage ring: two windows are always in flight per CTA (128 KB per SM) while the third is being computed, so
// the loads no longer stop while a CTA computes and stores (round 1: load everything, wait, compute, exit -- 0.45 of HBM).
// The "is this window inside one segment" test of window k + 1 is loaded while window k is computed, and the decoder's
// `y += ...` leaves as a vector RED, so nothing in the loop waits on global memory.
constexpr int FSMN_WIN = FSMN_RUN + 10;
constexpr int FSMN_STAGES = 3;
constexpr int FSMN_STAGE_BYTES = FSMN_WIN * 512 * 2;           // two [FSMN_WIN x 256] boxes
constexpr int FSMN_SMEM = FSMN_STAGES * FSMN_STAGE_BYTES + 64;

template <bool F16>
__global__ void __launch_bounds__(FSMN_THREADS, 2)
fsmn_kernel(const __grid_constant__ CUtensorMap tm_in, int col0, const float* __restrict__ w_t,
            const int2* __restrict__ row_info, int rows, const int* __restrict__ rows_dev, int mode,
            __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ y_f32) {
  extern __shared__ __align__(128) uint8_t fsmn_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(fsmn_smem + FSMN_STAGES * FSMN_STAGE_BYTES);
  if (threadIdx.x == 0) {
    for (int i = 0; i < FSMN_STAGES; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_in);
  }
  const int c = threadIdx.x * 2;
  float2 w[11];   // constants: loaded before the dependency wait
#pragma unroll
  for (int j = 0; j < 11; ++j) w[j] = *reinterpret_cast<const float2*>(w_t + j * 512 + c);
  w[5].x += 1.0f; w[5].y += 1.0f;  // identity branch
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  const int nrows = rows_dev ? *rows_dev : rows;
  const int n_win = (nrows + FSMN_RUN - 1) / FSMN_RUN;
  auto issue = [&](int k) {   // this CTA's k-th window
    const int win = blockIdx.x + k * gridDim.x;
    if (win >= n_win) return;
    const int stage = k % FSMN_STAGES;
    uint8_t* dst = fsmn_smem + stage * FSMN_STAGE_BYTES;
    mbar_arrive_expect_tx(&full[stage], FSMN_STAGE_BYTES);
    tma_load_2d(dst, &tm_in, &full[stage], col0, win * FSMN_RUN - 5);
    tma_load_2d(dst + FSMN_STAGE_BYTES / 2, &tm_in, &full[stage], col0 + 256, win * FSMN_RUN - 5);
  };
  if (threadIdx.x == 0) {
    for (int k = 0; k < FSMN_STAGES - 1; ++k) issue(k);
  }
  // block-uniform: first and last row of a window are frames of the same segment (rows of a segment are consecutive)
  auto edge_rows = [&](int win, int2& a, int2& b) {
    const int r0 = win * FSMN_RUN;
    a = make_int2(-1, 0); b = make_int2(-1, 0);
    if (win < n_win && r0 >= 5 && r0 + FSMN_RUN + 4 < nrows) { a = row_info[r0 - 5]; b = row_info[r0 + FSMN_RUN + 4]; }
  };
  int2 ea, eb;
  edge_rows(blockIdx.x, ea, eb);
  // this thread's channels inside a stage: box (threadIdx.x >> 7), 4 bytes at column 2 (threadIdx.x & 127)
  const int t_off = (threadIdx.x >> 7) * (FSMN_STAGE_BYTES / 2) + (threadIdx.x & 127) * 4;
  for (int k = 0;; ++k) {
    const int win = blockIdx.x + k * gridDim.x;
    if (win >= n_win) break;
    if (threadIdx.x == 0) issue(k + FSMN_STAGES - 1);   // its stage was released by the barrier that ended iteration k - 1
    const bool fast = ea.x >= 0 && eb.x == ea.x + FSMN_RUN + 9 && eb.y == ea.y;
    edge_rows(win + gridDim.x, ea, eb);                 // next window's test: in flight during this window's FMAs
    const int stage = k % FSMN_STAGES;
    mbar_wait(&full[stage], (uint32_t)((k / FSMN_STAGES) & 1));
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(fsmn_smem + stage * FSMN_STAGE_BYTES + t_off);
    const int r0 = win * FSMN_RUN;
    if (fast) fsmn_run<true, F16>(tile, w, row_info, nrows, r0, mode, out_bf16, y_f32);
    else fsmn_run<false, F16>(tile, w, row_info, nrows, r0, mode, out_bf16, y_f32);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// CIF
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cif_alpha_kernel(const float* __restrict__ h, int M, const float* __restrict__ w, const float* __restrict__ b,
                 const int2* __restrict__ row_info, float tail, float* __restrict__ alpha) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  if (row_info[row].x < 0) {
    if (lane == 0) alpha[row] = tail;  // tail_process_fn: extra frame with alpha = tail_threshold
    return;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(h + (size_t)row * 512 + (lane + 32 * i) * 4);
    const float4 ww = *reinterpret_cast<const float4*>(w + (lane + 32 * i) * 4);
    s += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float x = s + b[0];
    float al = 1.0f / (1.0f + expf(-x));
    al = fmaxf(al * 1.0f - 0.0f, 0.f);  // relu(alpha * smooth_factor - noise_threshold)
    alpha[row] = al;
  }
}

__global__ void __launch_bounds__(32)
cif_fire_kernel(const float* __restrict__ alpha, const int* __restrict__ row_off, const int* __restrict__ seg_T,
                float threshold, float* __restrict__ cur_o, float* __restrict__ rem_o, float* __restrict__ fire_val,
                int* __restrict__ n_tok, int* __restrict__ fire_row) {
  pdl_wait();
  pdl_launch_dependents();
  const int seg = blockIdx.x;
  const int lane = threadIdx.x;
  const int base = row_off[seg];
  const int n = seg_T[seg] + 1;  // frames + tail frame
  float integrate = 0.f;
  int ntok = 0;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int i = c0 + lane;
    const float a = (i < n) ? alpha[base + i] : 0.f;
    float my_cur = 0.f, my_rem = 0.f, my_fv = 0.f;
    bool my_fire = false;
    const int lim = min(32, n - c0);
    for (int k = 0; k < lim; ++k) {
      const float ak = __shfl_sync(0xffffffffu, a, k);
      const float dc = __fsub_rn(1.0f, integrate);        // distribution_completion
      integrate = __fadd_rn(integrate, ak);
      const float fv = integrate;
      const bool fire = integrate >= threshold;
      const float cur = fire ? dc : ak;
      const float rem = __fsub_rn(ak, cur);
      if (fire) integrate = __fsub_rn(integrate, 1.0f);
      if (lane == k) { my_cur = cur; my_rem = rem; my_fv = fv; my_fire = fire; }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, my_fire);
    if (i < n) {
      cur_o[base + i] = my_cur;
      rem_o[base + i] = my_rem;
      fire_val[base + i] = my_fv;
    }
    if (my_fire) fire_row[base + ntok + __popc(mask & ((1u << lane) - 1u))] = base + i;
    ntok += __popc(mask);
  }
  if (lane == 0) n_tok[seg] = ntok;
}

__global__ void __launch_bounds__(1024)
cif_scan_kernel(const int* __restrict__ n_tok, int n_seg, int* __restrict__ tok_off, int* __restrict__ total) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n_seg + 1023) / 1024;
  const int b = tid * per;
  int local = 0;
  for (int i = 0; i < per; ++i) if (b + i < n_seg) local += n_tok[b + i];
  int incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = wsum[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += t;
    }
    wsum[lane] = v;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? wsum[warp - 1] : 0);
  for (int i = 0; i < per; ++i) {
    if (b + i < n_seg) { tok_off[b + i] = run; run += n_tok[b + i]; }
  }
  if (tid == 1023) { tok_off[n_seg] = wsum[31]; *total = wsum[31]; }
}

__global__ void __launch_bounds__(128)
cif_embed_kernel(const float* __restrict__ enc, const float* __restrict__ cur, const float* __restrict__ rem,
                 const int* __restrict__ fire_row, const int* __restrict__ row_off, const int* __restrict__ tok_off,
                 int n_seg, float* __restrict__ emb, int2* __restrict__ tok_info, int* __restrict__ tok_frame) {
  pdl_wait();
  pdl_launch_dependents();
  const int g = blockIdx.x;
  if (g >= tok_off[n_seg]) return;
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tok_off[mid] <= g) lo = mid; else hi = mid;
  }
  // segments with zero tokens share an offset with their successor: move to the last one that starts here
  const int seg = lo;
  const int j = g - tok_off[seg];
  const int base = row_off[seg];
  const int f = fire_row[base + j];
  const int fprev = j > 0 ? fire_row[base + j - 1] : base - 1;
  const int c = threadIdx.x * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j > 0) {
    const float r = rem[fprev];
    const float4 hv = *reinterpret_cast<const float4*>(enc + (size_t)fprev * 512 + c);
    acc = make_float4(__fmul_rn(r, hv.x), __fmul_rn(r, hv.y), __fmul_rn(r, hv.z), __fmul_rn(r, hv.w));
  }
  for (int t = fprev + 1; t <= f; ++t) {
    const float cw = cur[t];
    const float4 hv = *reinterpret_cast<const float4*>(enc + (size_t)t * 512 + c);
    acc.x = __fadd_rn(acc.x, __fmul_rn(cw, hv.x));
    acc.y = __fadd_rn(acc.y, __fmul_rn(cw, hv.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(cw, hv.z));
    acc.w = __fadd_rn(acc.w, __fmul_rn(cw, hv.w));
  }
  *reinterpret_cast<float4*>(emb + (size_t)g * 512 + c) = acc;
  if (threadIdx.x == 0) {
    tok_info[g] = make_int2(j, tok_off[seg + 1] - tok_off[seg]);
    tok_frame[g] = f - base;
  }
}

__global__ void argmax_decode_kernel(const unsigned long long* __restrict__ packed, const int* __restrict__ n_dev, int cap,
                                     int* __restrict__ ids) {
  pdl_wait();
  pdl_launch_dependents();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_dev ? min(*n_dev, cap) : cap;
  if (g >= n) return;
  ids[g] = (int)(0xFFFFFFFFu - (uint32_t)(packed[g] & 0xFFFFFFFFull));
}

// K10  log-softmax statistics + top-k per decoder row (the pruned posterior the host-side log-prob consumers read:
// WfstDecoder::Search, wfst-decoder.cpp:27-57; CtcPrefixDecoder).  One warp per row.  The order is total and
// deterministic: value descending, index ascending (so entry 0 is FindMax's first-max-wins argmax, util.cpp:63-74).
__global__ void __launch_bounds__(256)
logprob_topk_kernel(const float* __restrict__ logits, int V, const int* __restrict__ n_dev, int cap, int k,
                    float* __restrict__ lse_out, float* __restrict__ lp_out, int* __restrict__ id_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n = n_dev ? min(*n_dev, cap) : cap;
  if (row >= n) return;
  const float* x = logits + (size_t)row * V;
  float m = -INFINITY;
  for (int i = lane * 4; i < V; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  float sum = 0.f;
  for (int i = lane * 4; i < V; i += 128) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    sum += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
  }
  sum = warp_sum(sum);
  const float lse = m + logf(sum);
  if (lane == 0) lse_out[row] = lse;
  // k rounds of "largest key strictly below the previous winner"; key = (value desc, index asc)
  float pv = INFINITY;
  int pi = -1;
  for (int j = 0; j < k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane * 4; i < V; i += 128) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float val = e[q];
        const int idx = i + q;
        const bool below_prev = (val < pv) || (val == pv && idx > pi);
        const bool better = (val > bv) || (val == bv && idx < bi);
        if (below_prev && better) { bv = val; bi = idx; }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      lp_out[(size_t)row * k + j] = bv - lse;
      id_out[(size_t)row * k + j] = bi;
    }
    pv = bv;
    pi = bi;
  }
}

template <bool F16>
__global__ void f32_to_h16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = pack_h1<F16>(in[i]);
}
template <bool F16>
__global__ void h16_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = unpack_h2<F16>((uint32_t)in[i]).x;
}

// LayerNorm over 512 fp32 columns with the rows staged through shared memory by cp.async: a warp owns 4 consecutive rows, every
// lane issues the 16 16-byte copies of its own elements up front (one commit group per row: 8 KB per warp, 64 KB per CTA and
// 192 KB per SM in flight -- three times what register loads allowed) and normalises row k as soon as group k has landed.  A
// lane reads back only what it copied itself, so no barrier is needed; the arithmetic is that of layernorm_kernel<512, false>.
constexpr int LNS_ROWS_PER_WARP = 4, LNS_WARPS = 8, LNS_SMEM = LNS_WARPS * LNS_ROWS_PER_WARP * 512 * 4;
template <bool F16>
__global__ void __launch_bounds__(256)
layernorm512_staged_kernel(const float* __restrict__ in, int rows, const int* __restrict__ rows_dev, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                           const int2* __restrict__ row_info, int zero_gap) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float lns_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = rows_dev ? *rows_dev : rows;
  const int row0 = (blockIdx.x * LNS_WARPS + warp) * LNS_ROWS_PER_WARP;
  if (row0 >= nrows) return;
  float* mine = lns_smem + (size_t)warp * LNS_ROWS_PER_WARP * 512;
#pragma unroll
  for (int k = 0; k < LNS_ROWS_PER_WARP; ++k) {
    if (row0 + k < nrows) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (lane + 32 * i) * 4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(mine + k * 512 + c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(in + (size_t)(row0 + k) * 512 + c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float4 g[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    g[i] = *reinterpret_cast<const float4*>(gamma + (lane + 32 * i) * 4);
    b[i] = *reinterpret_cast<const float4*>(beta + (lane + 32 * i) * 4);
  }
#pragma unroll
  for (int k = 0; k < LNS_ROWS_PER_WARP; ++k) {
    if (k == 0) asm volatile("cp.async.wait_group 3;" ::: "memory");
    else if (k == 1) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (k == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int row = row0 + k;
    if (row >= nrows) break;
    const bool gap = zero_gap && row_info && row_info[row].x < 0;
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 f = *reinterpret_cast<const float4*>(mine + k * 512 + (lane + 32 * i) * 4);
      v[i * 4] = f.x; v[i * 4 + 1] = f.y; v[i * 4 + 2] = f.z; v[i * 4 + 3] = f.w;
#pragma unroll
      for (int e = 0; e < 4; ++e) sum += v[i * 4 + e];
    }
    const float mean = warp_sum(sum) / 512.0f;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float d = v[i * 4 + e] - mean; sq += d * d; }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / 512.0f + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (lane + 32 * i) * 4;
      float o[4];
      o[0] = (v[i * 4] - mean) * rstd * g[i].x + b[i].x;
      o[1] = (v[i * 4 + 1] - mean) * rstd * g[i].y + b[i].y;
      o[2] = (v[i * 4 + 2] - mean) * rstd * g[i].z + b[i].z;
      o[3] = (v[i * 4 + 3] - mean) * rstd * g[i].w + b[i].w;
      if (gap) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = 0.f; }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * 512 + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (out_bf16) {
        *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * 512 + c) = make_uint2(pack_h2<F16>(o[0], o[1]), pack_h2<F16>(o[2], o[3]));
      }
    }
  }
}

// The decoder's LayerNorm(2048) on the bf16 FFN hidden rows, staged the same way: a warp owns 2 rows (8 KB), copies them with
// cp.async (one commit group per row) and normalises row k when it has landed; arithmetic of layernorm_kernel<2048, true>.
constexpr int LNB_ROWS_PER_WARP = 2, LNB_SMEM = LNS_WARPS * LNB_ROWS_PER_WARP * 2048 * 2;
template <bool F16>
__global__ void __launch_bounds__(256)
layernorm2048_staged_kernel(const __nv_bfloat16* in, int rows, const int* __restrict__ rows_dev, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float eps, __nv_bfloat16* out_bf16, float* __restrict__ out_f32) {
  // in and out_bf16 may be the same buffer (the decoder normalises its FFN hidden rows in place): a lane writes only the
  // elements it copied itself, after it has read them from shared memory
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float lns_smem[];
  __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(lns_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = rows_dev ? *rows_dev : rows;
  const int row0 = (blockIdx.x * LNS_WARPS + warp) * LNB_ROWS_PER_WARP;
  if (row0 >= nrows) return;
  __nv_bfloat16* mine = base + (size_t)warp * LNB_ROWS_PER_WARP * 2048;
#pragma unroll
  for (int k = 0; k < LNB_ROWS_PER_WARP; ++k) {
    if (row0 + k < nrows) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = (lane + 32 * i) * 8;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(mine + k * 2048 + c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(in + (size_t)(row0 + k) * 2048 + c) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll
  for (int k = 0; k < LNB_ROWS_PER_WARP; ++k) {
    if (k == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int row = row0 + k;
    if (row >= nrows) break;
    float v[64];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = *reinterpret_cast<const uint4*>(mine + k * 2048 + (lane + 32 * i) * 8);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = unpack_h2<F16>(uw[e]); v[i * 8 + 2 * e] = f.x; v[i * 8 + 2 * e + 1] = f.y; }
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i * 8 + e];
    }
    const float mean = warp_sum(sum) / 2048.0f;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[i * 8 + e] - mean; sq += d * d; }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / 2048.0f + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = (lane + 32 * i) * 8;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + c + e);
        const float4 b = *reinterpret_cast<const float4*>(beta + c + e);
        o[e] = (v[i * 8 + e] - mean) * rstd * g.x + b.x;
        o[e + 1] = (v[i * 8 + e + 1] - mean) * rstd * g.y + b.y;
        o[e + 2] = (v[i * 8 + e + 2] - mean) * rstd * g.z + b.z;
        o[e + 3] = (v[i * 8 + e + 3] - mean) * rstd * g.w + b.w;
      }
      if (out_f32) {
        *reinterpret_cast<float4*>(out_f32 + (size_t)row * 2048 + c) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(out_f32 + (size_t)row * 2048 + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (out_bf16) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pk[e] = pack_h2<F16>(o[2 * e], o[2 * e + 1]);
        *reinterpret_cast<uint4*>(out_bf16 + (size_t)row * 2048 + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

template <int D, bool F16>
int ln_dispatch(const void* in, int in_is_bf16, int rows, const int* rows_dev, const float* gamma, const float* beta,
                float eps, __nv_bfloat16* ob, float* of, const int2* ri, int zg, cudaStream_t s) {
  const int blocks = (rows + 7) / 8;
  static const bool plain = getenv("B200PF_LN_PLAIN") != nullptr;
  if (D == 512 && !in_is_bf16 && rows >= 4096 && !plain) {   // large fp32 LayerNorms (the residual stream): staged loads
    static PerDeviceOnce once;
    const int rc = once_per_device(once, [] {
      return (int)cudaFuncSetAttribute(layernorm512_staged_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, LNS_SMEM);
    });
    if (rc) return rc;
    const int per = LNS_WARPS * LNS_ROWS_PER_WARP;
    return launch_kernel(layernorm512_staged_kernel<F16>, dim3((rows + per - 1) / per), dim3(256), LNS_SMEM, s, (const float*)in, rows, rows_dev, gamma,
                         beta, eps, ob, of, ri, zg);
  }
  if (D == 2048 && in_is_bf16 && rows >= 4096 && !ri && !plain) {   // the decoder's FFN LayerNorm on large batches: staged loads
    static PerDeviceOnce once;
    const int rc = once_per_device(once, [] {
      return (int)cudaFuncSetAttribute(layernorm2048_staged_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, LNB_SMEM);
    });
    if (rc) return rc;
    const int per = LNS_WARPS * LNB_ROWS_PER_WARP;
    return launch_kernel(layernorm2048_staged_kernel<F16>, dim3((rows + per - 1) / per), dim3(256), LNB_SMEM, s, (const __nv_bfloat16*)in, rows,
                         rows_dev, gamma, beta, eps, ob, of);
  }
  if (in_is_bf16)
    return launch_kernel(layernorm_kernel<D, true, F16>, dim3(blocks), dim3(256), 0, s, in, rows, rows_dev, gamma, beta, eps, ob, of, ri, zg);
  return launch_kernel(layernorm_kernel<D, false, F16>, dim3(blocks), dim3(256), 0, s, in, rows, rows_dev, gamma, beta, eps, ob, of, ri, zg);
}

}  // namespace

int layernorm_launch(const void* in, int in_is_bf16, int rows, const int* rows_dev, int D, const float* gamma,
                     const float* beta, float eps, __nv_bfloat16* out_bf16, float* out_f32, const int2* row_info,
                     int zero_gap, cudaStream_t s, int f16) {
  if (rows <= 0) return 0;
#define PF_LN_CASE(DD)                                                                                                          \
  case DD:                                                                                                                      \
    return f16 ? ln_dispatch<DD, true>(in, in_is_bf16, rows, rows_dev, gamma, beta, eps, out_bf16, out_f32, row_info, zero_gap, s) \
               : ln_dispatch<DD, false>(in, in_is_bf16, rows, rows_dev, gamma, beta, eps, out_bf16, out_f32, row_info, zero_gap, s);
  switch (D) {
    PF_LN_CASE(512)
    PF_LN_CASE(560)
    PF_LN_CASE(2048)
    default: return (int)cudaErrorInvalidValue;
  }
#undef PF_LN_CASE
}

int fsmn_launch(const __nv_bfloat16* in, int ld_in, int col0, const float* w_t, const int2* row_info, int rows,
                const int* rows_dev, int mode, __nv_bfloat16* out_bf16, float* y_f32, cudaStream_t s, int f16, int num_sms) {
  if (rows <= 0) return 0;
  static PerDeviceOnce once;
  int rc = once_per_device(once, [] {
    cudaError_t err = cudaFuncSetAttribute(fsmn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FSMN_SMEM);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(fsmn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FSMN_SMEM);
    return (int)err;
  });
  if (rc) return rc;
  CUtensorMap tm;
  rc = make_tmap_bf16_plain(&tm, in, (uint64_t)rows, (uint64_t)(col0 + 512), (uint64_t)ld_in, FSMN_WIN, 256);
  if (rc) return rc;
  const int n_win = (rows + FSMN_RUN - 1) / FSMN_RUN, resident = 2 * (num_sms > 0 ? num_sms : 148);
  const dim3 grid(n_win < resident ? n_win : resident);
  if (f16) return launch_kernel(fsmn_kernel<true>, grid, dim3(FSMN_THREADS), FSMN_SMEM, s, tm, col0, w_t, row_info, rows, rows_dev, mode, out_bf16, y_f32);
  return launch_kernel(fsmn_kernel<false>, grid, dim3(FSMN_THREADS), FSMN_SMEM, s, tm, col0, w_t, row_info, rows, rows_dev, mode, out_bf16, y_f32);
}

int cif_alpha_launch(const float* h, int M, const float* w, const float* b, const int2* row_info, float tail,
                     float* alpha, cudaStream_t s) {
  if (M <= 0) return 0;
  return launch_kernel(cif_alpha_kernel, dim3((M + 7) / 8), dim3(256), 0, s, h, M, w, b, row_info, tail, alpha);
}

int cif_fire_launch(const float* alpha, const int* row_off, const int* seg_T, int n_seg, float threshold, float* cur,
                    float* rem, float* fire_val, int* n_tok, int* fire_row, cudaStream_t s) {
  if (n_seg <= 0) return 0;
  return launch_kernel(cif_fire_kernel, dim3(n_seg), dim3(32), 0, s, alpha, row_off, seg_T, threshold, cur, rem, fire_val, n_tok, fire_row);
}

int cif_scan_launch(const int* n_tok, int n_seg, int* tok_off, int* n_tok_total, cudaStream_t s) {
  if (n_seg <= 0) return 0;
  return launch_kernel(cif_scan_kernel, dim3(1), dim3(1024), 0, s, n_tok, n_seg, tok_off, n_tok_total);
}

int cif_embed_launch(const float* enc_f32, const float* cur, const float* rem, const int* fire_row, const int* row_off,
                     const int* tok_off, int n_seg, int tok_cap, float* emb, int2* tok_info, int* tok_frame,
                     cudaStream_t s) {
  if (tok_cap <= 0) return 0;
  return launch_kernel(cif_embed_kernel, dim3(tok_cap), dim3(128), 0, s, enc_f32, cur, rem, fire_row, row_off, tok_off, n_seg, emb, tok_info, tok_frame);
}

int argmax_decode_launch(const unsigned long long* packed, const int* n_dev, int cap, int* ids, cudaStream_t s) {
  if (cap <= 0) return 0;
  return launch_kernel(argmax_decode_kernel, dim3((cap + 255) / 256), dim3(256), 0, s, packed, n_dev, cap, ids);
}

int logprob_topk_launch(const float* logits, int V, const int* n_dev, int cap, int k, float* lse, float* lp, int* ids, cudaStream_t s) {
  if (cap <= 0 || k <= 0) return 0;
  if (V & 3) return (int)cudaErrorInvalidValue;
  return launch_kernel(logprob_topk_kernel, dim3((cap + 7) / 8), dim3(256), 0, s, logits, V, n_dev, cap, k, lse, lp, ids);
}

int f32_to_bf16_launch(const float* in, __nv_bfloat16* out, int64_t n, cudaStream_t s, int f16) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (f16) return launch_kernel(f32_to_h16_kernel<true>, dim3((int)blocks), dim3(256), 0, s, in, reinterpret_cast<uint16_t*>(out), n);
  return launch_kernel(f32_to_h16_kernel<false>, dim3((int)blocks), dim3(256), 0, s, in, reinterpret_cast<uint16_t*>(out), n);
}

int h16_to_f32_launch(const __nv_bfloat16* in, float* out, int64_t n, cudaStream_t s, int f16) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (f16) return launch_kernel(h16_to_f32_kernel<true>, dim3((int)blocks), dim3(256), 0, s, reinterpret_cast<const uint16_t*>(in), out, n);
  return launch_kernel(h16_to_f32_kernel<false>, dim3((int)blocks), dim3(256), 0, s, reinterpret_cast<const uint16_t*>(in), out, n);
}

}  // namespace pf
