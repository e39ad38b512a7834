// K12  LSTM recurrence (hidden 512) for the config-3 graph pieces the reference runs inside onnxruntime:
//   * the bidirectional LSTM of the timestamp predictor (CifPredictorV3 `cnn_blstm` head; outputs consumed at
//     onnxruntime/src/paraformer.cpp:549-563 and by TimestampOnnx, util.cpp:838-963),
//   * the hotword compiler's LSTM (model_eb.onnx, Paraformer::CompileHotwordEmbedding, paraformer.cpp:592-693).
// plus the two small kernels that turn the BiLSTM output into us_alphas / us_cif_peak.
//
// The input projection x·W_ih^T + b_ih + b_hh is an ordinary GEMM (gemm.cu).  What is left is the sequential part
//   g_t = gx_t + W_hh h_{t-1};  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)        (gate order i, f, g, o)
// which is latency bound.  One CLUSTER of 16 CTAs owns one (group of 32 sequences, direction):
//   * W_hh (2 MB bf16) is distributed over the cluster's shared memory: CTA r keeps the 4 x 32 gate rows of hidden
//     units [32 r, 32 r + 32) (128 KB), loaded once;
//   * every step each CTA computes its [128 gate rows x 32 sequences] slab with mma.sync.m16n8k16 (bf16 in, fp32
//     accumulate; A = W slab, B = h_{t-1} of all 512 units), applies the gate non-linearities in registers (the
//     cell state never leaves registers), and broadcasts its 32 new hidden units to the other 15 CTAs through
//     distributed shared memory (16-byte st.shared::cluster), double buffered, one cluster barrier per step.
// Sequences of one group advance in lock step; shorter ones simply stop early, so callers sort by length.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace pf {
namespace {

constexpr int kCluster = 16;       // CTAs per cluster (non-portable size; 16 x 32 units = 512)
constexpr int kGroup = 32;         // sequences per cluster
constexpr int kThreads = 256;
constexpr int kWBytes = 128 * 1024;       // [128 gate rows][512] bf16, 16-byte chunks XOR-swizzled by (row & 7)
constexpr int kHBytes = 32 * 1024;        // one h buffer: [16 ranks][32 seq][32 units] bf16, chunk ^ ((seq >> 1) & 3)
constexpr int kXBytes = 16 * 1024;        // split-K exchange: [8 warps][16][32 lanes] fp32
constexpr int kSmem = kWBytes + 2 * kHBytes + kXBytes + 2 * kGroup * 4;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_128(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// exp through the MUFU unit (ex2.approx, ~2 ulp) and an approximate reciprocal: ~1e-6 relative, far inside the bf16
// rounding of the recurrent state, at a tenth of the instructions of expf()/tanhf()
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
lstm_kernel(LstmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sH = smem + kWBytes;
  float* sX = reinterpret_cast<float*>(smem + kWBytes + 2 * kHBytes);
  int* sOff = reinterpret_cast<int*>(smem + kWBytes + 2 * kHBytes + kXBytes);
  int* sLen = sOff + kGroup;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int task = blockIdx.x / kCluster;
  const int dir = task % p.n_dir, group = task / p.n_dir;
  const bool rev = (p.reverse_mask >> dir) & 1;
  const int s0 = group * kGroup;

  if (tid < kGroup) {
    const int s = s0 + tid;
    sOff[tid] = s < p.n_seq ? p.seq_off[s] : 0;
    sLen[tid] = s < p.n_seq ? p.seq_len[s] : 0;
  }
  // W_hh slab of this CTA.  Slab row R = ug*32 + mt*16 + half*8 + j holds gate (2 mt + half) of hidden unit
  // 32 rank + 8 ug + j, so that the warp of unit group `ug` finds i,f in its m-tile 0 and g,o in its m-tile 1.
  {
    const __nv_bfloat16* wsrc = p.whh + (size_t)dir * 2048 * 512;
    for (int idx = tid; idx < 128 * 64; idx += kThreads) {
      const int R = idx >> 6, c = idx & 63;
      const int ug = R >> 5, mt = (R >> 4) & 1, half = (R >> 3) & 1, j = R & 7;
      const int grow = (mt * 2 + half) * 512 + (int)rank * 32 + ug * 8 + j;
      const uint4 v = ldg128_nc(wsrc + (size_t)grow * 512 + c * 8);
      *reinterpret_cast<uint4*>(sW + R * 1024 + ((c ^ (R & 7)) << 4)) = v;
    }
    for (int idx = tid; idx < 2 * kHBytes / 16; idx += kThreads) reinterpret_cast<uint4*>(sH)[idx] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  int steps = 0;
#pragma unroll 1
  for (int i = 0; i < kGroup; ++i) steps = max(steps, sLen[i]);
  cluster_sync_all();  // every CTA's h buffers are zero before anyone writes remotely

  const int ug = warp & 3, kh = warp >> 2, gid = lane >> 2, q = lane & 3;
  const int unit_local = ug * 8 + gid;
  // this thread finalises 4 (unit, sequence) cells: sequence 16 kh + 8 jj + 2 q + c
  int c_off[4], c_len[4];
  float c_state[4];
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int sl = 16 * kh + 8 * (x >> 1) + 2 * q + (x & 1);
    c_off[x] = sOff[sl];
    c_len[x] = sLen[sl];
    c_state[x] = 0.f;
  }
  const unsigned short* gx = reinterpret_cast<const unsigned short*>(p.gx);
  const int gcol = dir * 2048 + (int)rank * 32 + unit_local;

  const uint32_t sW_u = smem_u32(sW), sH_u = smem_u32(sH), sX_u = smem_u32(sX);
  // ldmatrix lane addresses (see the fragment layouts of mma.m16n8k16)
  const int a_row0 = ug * 32 + (lane & 7) + 8 * ((lane >> 3) & 1);   // m-tile 0; m-tile 1 is +16 rows
  const int a_csel = lane >> 4;
  const int b_n = (lane & 7) + 8 * (lane >> 4);                      // + 16 np
  const int b_csel = (lane >> 3) & 1;

  for (int k = 0; k < steps; ++k) {
    // gate pre-activations of this step from the input projection (latency hidden behind the MMA phase)
    unsigned short gxv[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const bool act = k < c_len[x];
      const int t = rev ? c_len[x] - 1 - k : k;
      const unsigned short* src = gx + (size_t)(c_off[x] + t) * p.ld_gx + gcol;
#pragma unroll
      for (int g = 0; g < 4; ++g) gxv[x][g] = act ? __ldg(src + g * 512) : (unsigned short)0;
    }

    const uint32_t hb = sH_u + (uint32_t)(k & 1) * kHBytes;          // h_{t-1}
    const uint32_t hn = sH_u + (uint32_t)((k + 1) & 1) * kHBytes;    // h_t
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
      const int ks = kh * 16 + kk;   // this warp's half of K
      uint32_t a0[4], a1[4], b01[4], b23[4];
      {
        const int R0 = a_row0, R1 = a_row0 + 16;
        const int ch = 2 * ks + a_csel;
        ldmatrix_x4(sW_u + R0 * 1024 + ((ch ^ (R0 & 7)) << 4), a0);
        ldmatrix_x4(sW_u + R1 * 1024 + ((ch ^ (R1 & 7)) << 4), a1);
      }
      {
        const int ch = 2 * (ks & 1) + b_csel;
        const int n0 = b_n, n1 = b_n + 16;
        const uint32_t blk = hb + (uint32_t)(ks >> 1) * 2048;
        ldmatrix_x4(blk + n0 * 64 + ((ch ^ ((n0 >> 1) & 3)) << 4), b01);
        ldmatrix_x4(blk + n1 * 64 + ((ch ^ ((n1 >> 1) & 3)) << 4), b23);
      }
      mma_bf16_16816<F16>(acc[0][0], a0, b01[0], b01[1]);
      mma_bf16_16816<F16>(acc[0][1], a0, b01[2], b01[3]);
      mma_bf16_16816<F16>(acc[0][2], a0, b23[0], b23[1]);
      mma_bf16_16816<F16>(acc[0][3], a0, b23[2], b23[3]);
      mma_bf16_16816<F16>(acc[1][0], a1, b01[0], b01[1]);
      mma_bf16_16816<F16>(acc[1][1], a1, b01[2], b01[3]);
      mma_bf16_16816<F16>(acc[1][2], a1, b23[0], b23[1]);
      mma_bf16_16816<F16>(acc[1][3], a1, b23[2], b23[3]);
    }

    // split-K: hand the two n-tiles the partner warp finalises to it, keep n-tiles {2 kh, 2 kh + 1}
    // (static accumulator indices with a runtime predicate keep acc[] in registers)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        if ((nt >> 1) != kh) {
#pragma unroll
          for (int i = 0; i < 4; ++i) sX[(warp * 16 + mt * 8 + (nt & 1) * 4 + i) * 32 + lane] = acc[mt][nt][i];
        }
    __syncthreads();
    const int pw = ug + 4 * (1 - kh);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        if ((nt >> 1) == kh) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][nt][i] += sX[(pw * 16 + mt * 8 + (nt & 1) * 4 + i) * 32 + lane];
        }

    // gates; accumulator rows: m-tile 0 = {i (gid), f (gid + 8)}, m-tile 1 = {g, o}
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int jj = x >> 1, c = x & 1;
      const bool act = k < c_len[x];
      const float gi = (kh ? acc[0][2 + jj][c] : acc[0][jj][c]) + unpack_h2<F16>((uint32_t)gxv[x][0]).x;
      const float gf = (kh ? acc[0][2 + jj][2 + c] : acc[0][jj][2 + c]) + unpack_h2<F16>((uint32_t)gxv[x][1]).x;
      const float gg = (kh ? acc[1][2 + jj][c] : acc[1][jj][c]) + unpack_h2<F16>((uint32_t)gxv[x][2]).x;
      const float go = (kh ? acc[1][2 + jj][2 + c] : acc[1][jj][2 + c]) + unpack_h2<F16>((uint32_t)gxv[x][3]).x;
      float h = 0.f;
      if (act) {
        c_state[x] = sigmoidf_(gf) * c_state[x] + sigmoidf_(gi) * tanhf_(gg);
        h = sigmoidf_(go) * tanhf_(c_state[x]);
        if (p.out_f32) {
          const int t = rev ? c_len[x] - 1 - k : k;
          p.out_f32[(size_t)(c_off[x] + t) * p.ld_out_f32 + dir * 512 + (int)rank * 32 + unit_local] = h;
        }
      }
      const int sl = 16 * kh + 8 * jj + 2 * q + c;
      *reinterpret_cast<uint16_t*>(sH + ((k + 1) & 1) * kHBytes + rank * 2048 + sl * 64 + ((ug ^ ((sl >> 1) & 3)) << 4) + gid * 2) = pack_h1<F16>(h);
    }
    __syncthreads();

    // broadcast this CTA's 2 KB block of h_t to the same place in every other CTA of the cluster
    {
      const int chunk = tid & 127;
      const uint32_t loc = hn + rank * 2048 + chunk * 16;
      const uint4 v = lds128(loc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t dst = (uint32_t)((tid >> 7) + 2 * i);
        if (dst != rank) st_cluster_128(mapa_u32(loc, dst), v);
      }
      if (tid < 128 && p.out_bf16) {
        const int sl = chunk >> 2, pc = chunk & 3;
        const int lc = pc ^ ((sl >> 1) & 3);
        const int len = sLen[sl];
        if (k < len) {
          const int t = rev ? len - 1 - k : k;
          stg128(p.out_bf16 + (size_t)(sOff[sl] + t) * p.ld_out + dir * 512 + (int)rank * 32 + lc * 8, v);
        }
      }
    }
    cluster_sync_all();
  }
}


// alpha2 = relu(sigmoid(h . w + b) * smooth - noise) over [rows, 1024] bf16; one warp per row.
template <bool F16>
__global__ void __launch_bounds__(256)
us_alpha_kernel(const __nv_bfloat16* __restrict__ h, int rows, const float* __restrict__ w, const float* __restrict__ b,
                float smooth, float noise, float* __restrict__ alpha) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 8;
    const uint4 u = ldg128_nc(h + (size_t)row * 1024 + c);
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
    const float4 w0 = *reinterpret_cast<const float4*>(w + c), w1 = *reinterpret_cast<const float4*>(w + c + 4);
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 hv = unpack_h2<F16>(uu[j]);
      s += hv.x * ww[2 * j];
      s += hv.y * ww[2 * j + 1];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const float a = 1.0f / (1.0f + expf(-(s + b[0])));
    alpha[row] = fmaxf(a * smooth - noise, 0.f);
  }
}

// Per segment: rescale alpha2 so that it sums to token_num, then cif_wo_hidden (running sum recorded before the
// subtraction, subtract `threshold` on fire).  One warp per segment; the scan itself is the reference's sequential
// fp32 recurrence.
__global__ void __launch_bounds__(32)
us_peak_kernel(const float* __restrict__ alpha2, const int* __restrict__ seq_off, const int* __restrict__ seq_len,
               const int* __restrict__ n_tok, float threshold, float* __restrict__ us_alphas, float* __restrict__ us_peaks) {
  pdl_wait();
  pdl_launch_dependents();
  const int seg = blockIdx.x, lane = threadIdx.x;
  const int base = seq_off[seg], n = seq_len[seg];
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += alpha2[base + i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float ratio = __fdiv_rn((float)n_tok[seg], s);
  float integrate = 0.f;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int i = c0 + lane;
    const float a = (i < n) ? __fmul_rn(alpha2[base + i], ratio) : 0.f;
    float mine = 0.f;
    const int lim = min(32, n - c0);
    for (int k = 0; k < lim; ++k) {
      const float ak = __shfl_sync(0xffffffffu, a, k);
      integrate = __fadd_rn(integrate, ak);
      if (lane == k) mine = integrate;
      if (integrate >= threshold) integrate = __fsub_rn(integrate, threshold);
    }
    if (i < n) { us_alphas[base + i] = a; us_peaks[base + i] = mine; }
  }
}

// out[j][:] = table[ids[j]][:] as bf16 (hotword Embedding lookup); ids outside [0, vocab) give zeros.
__global__ void __launch_bounds__(128)
embed_gather_kernel(const __nv_bfloat16* __restrict__ table, int vocab, const int* __restrict__ ids, int n, __nv_bfloat16* __restrict__ out) {
  const int j = blockIdx.x;
  if (j >= n) return;
  const int id = ids[j];
  uint2 v = make_uint2(0, 0);
  if (id >= 0 && id < vocab) v = *reinterpret_cast<const uint2*>(table + (size_t)id * 512 + threadIdx.x * 4);
  *reinterpret_cast<uint2*>(out + (size_t)j * 512 + threadIdx.x * 4) = v;
}

}  // namespace

template <bool F16>
static int lstm_launch_fmt(const LstmParams& p, cudaStream_t s) {
  static PerDeviceOnce once;
  const int rc = once_per_device(once, [] {
    cudaError_t err = cudaFuncSetAttribute(lstm_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(lstm_kernel<F16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return (int)err;
  });
  if (rc) return rc;
  const int groups = (p.n_seq + kGroup - 1) / kGroup;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * p.n_dir * kCluster));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, lstm_kernel<F16>, p);
}

// The mma.sync kernel is the product (5.3-5.5 us per step).  A tcgen05 variant of the step was measured slower in round 1
// (6.5 us; csrc/experiments/lstm_tc_kernel.cuh, not built).
int lstm_launch(const LstmParams& p, cudaStream_t s) {
  if (p.n_seq <= 0) return 0;
  if (p.n_dir < 1 || p.n_dir > 2 || (p.ld_gx & 1) || (p.out_bf16 && (p.ld_out & 7))) return (int)cudaErrorInvalidValue;
  return p.f16 ? lstm_launch_fmt<true>(p, s) : lstm_launch_fmt<false>(p, s);
}

int lstm_max_active_clusters() {
  cudaFuncSetAttribute(lstm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(lstm_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCluster * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_kernel<false>, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

int us_alpha_launch(const __nv_bfloat16* h, int rows, const float* w, const float* b, float smooth, float noise, float* alpha,
                    cudaStream_t s, int f16) {
  if (rows <= 0) return 0;
  if (f16) return launch_kernel(us_alpha_kernel<true>, dim3((rows + 7) / 8), dim3(256), 0, s, h, rows, w, b, smooth, noise, alpha);
  return launch_kernel(us_alpha_kernel<false>, dim3((rows + 7) / 8), dim3(256), 0, s, h, rows, w, b, smooth, noise, alpha);
}

int us_peak_launch(const float* alpha2, const int* seq_off, const int* seq_len, const int* n_tok, int n_seg, float threshold,
                   float* us_alphas, float* us_peaks, cudaStream_t s) {
  if (n_seg <= 0) return 0;
  return launch_kernel(us_peak_kernel, dim3(n_seg), dim3(32), 0, s, alpha2, seq_off, seq_len, n_tok, threshold, us_alphas, us_peaks);
}

int embed_gather_launch(const __nv_bfloat16* table, int vocab, const int* ids, int n, __nv_bfloat16* out, cudaStream_t s) {
  if (n <= 0) return 0;
  embed_gather_kernel<<<n, 128, 0, s>>>(table, vocab, ids, n, out);
  return (int)cudaGetLastError();
}

}  // namespace pf
