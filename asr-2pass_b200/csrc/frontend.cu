// K1: Kaldi-compatible fbank + LFR/CMVN + encoder input scaling / position encoding.
// Follows, operation for operation:
//   Paraformer::FbankKaldi           onnxruntime/src/paraformer.cpp:309-323
//   knf ExtractWindow/ProcessWindow  third_party/kaldi-native-fbank/kaldi-native-fbank/csrc/feature-window.cc:121-245
//   knf Rfft (double precision)      .../rfft.cc:35-62
//   knf ComputePowerSpectrum         .../feature-functions.cc:28-47
//   knf MelBanks::Compute, log       .../mel-computations.cc:224-235, feature-fbank.cc:102-108
//   Paraformer::LfrCmvn              onnxruntime/src/paraformer.cpp:421-461
// fp32 steps use explicit round-to-nearest mul/add (no FMA contraction) in the reference's order, so
// everything except the FFT (double, like knf, but a different butterfly order) and logf is bit-identical.
#include <float.h>

#include "kernels.cuh"

namespace pf {
namespace {

constexpr int WIN = 400, SHIFT = 160, NFFT = 512, NBIN = 80, FEAT = 560;
constexpr int FB_WARPS = 8;

struct FbSmem {
  double2 buf[FB_WARPS][NFFT];
  double2 tw[NFFT / 2];
  float window[WIN];
  float power[FB_WARPS][NFFT / 2 + 8];
  float mel_w[1024];
  int2 mel_range[NBIN];
  int mel_w_off[NBIN];
};

template <bool F32>
__device__ __forceinline__ float load_sample(const void* pcm, int64_t i) {
  if (F32) return __fmul_rn(reinterpret_cast<const float*>(pcm)[i], 32768.0f);  // paraformer.cpp:313
  return (float)reinterpret_cast<const int16_t*>(pcm)[i];                        // == (s/32768.f)*32768 exactly
}

template <bool F32>
__global__ void __launch_bounds__(FB_WARPS * 32)
fbank_kernel(const void* __restrict__ pcm, const int64_t* __restrict__ sample_off, const int* __restrict__ fb_off,
             int n_seg, int n_frames, FrontendTables t, int mel_w_total, float* __restrict__ fb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FbSmem& S = *reinterpret_cast<FbSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x) S.tw[i] = t.twiddle[i];
  for (int i = threadIdx.x; i < WIN; i += blockDim.x) S.window[i] = t.window[i];
  for (int i = threadIdx.x; i < mel_w_total; i += blockDim.x) S.mel_w[i] = t.mel_w[i];
  for (int i = threadIdx.x; i < NBIN; i += blockDim.x) { S.mel_range[i] = t.mel_range[i]; S.mel_w_off[i] = t.mel_w_off[i]; }
  __syncthreads();

  double2* buf = S.buf[warp];
  float* pw = S.power[warp];
  for (int fg = blockIdx.x * FB_WARPS + warp; fg < n_frames; fg += gridDim.x * FB_WARPS) {
    // frame -> segment (largest s with fb_off[s] <= fg)
    int lo = 0, hi = n_seg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (fb_off[mid] <= fg) lo = mid; else hi = mid;
    }
    const int64_t s0 = sample_off[lo] + (int64_t)(fg - fb_off[lo]) * SHIFT;

    // RemoveDcOffset (feature-window.cc:179-190)
    float x[13];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
      const int i = lane + 32 * k;
      x[k] = (i < WIN) ? load_sample<F32>(pcm, s0 + i) : 0.f;
      sum += x[k];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float mean = sum / (float)WIN;
    // Preemphasize (:200-211) then window (:57-63); zero pad to 512 and bit-reverse for the DIT FFT
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = lane + 32 * k;
      float w = 0.f;
      if (i < WIN) {
        const float d = __fsub_rn(x[k < 13 ? k : 0], mean);
        const float dp = (i > 0) ? __fsub_rn(load_sample<F32>(pcm, s0 + i - 1), mean) : d;
        w = __fsub_rn(d, __fmul_rn(0.97f, dp));
        w = __fmul_rn(w, S.window[i]);
      }
      buf[__brev((unsigned)i) >> 23] = make_double2((double)w, 0.0);
    }
    __syncwarp();
    // 512-point complex FFT in double (rfft.cc:41-47 narrows the double result to float)
#pragma unroll 1
    for (int s = 1; s <= 9; ++s) {
      const int half = 1 << (s - 1);
      const int tstep = (NFFT / 2) >> (s - 1);
#pragma unroll
      for (int b = lane; b < NFFT / 2; b += 32) {
        const int pos = b & (half - 1);
        const int i0 = ((b >> (s - 1)) << s) + pos;
        const int i1 = i0 + half;
        const double2 tw = S.tw[pos * tstep];
        const double2 u = buf[i0], v = buf[i1];
        const double vr = v.x * tw.x - v.y * tw.y;
        const double vi = v.x * tw.y + v.y * tw.x;
        buf[i0] = make_double2(u.x + vr, u.y + vi);
        buf[i1] = make_double2(u.x - vr, u.y - vi);
      }
      __syncwarp();
    }
    // ComputePowerSpectrum (feature-functions.cc:28-47); bin 256 is never used by the mel bank
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = lane + 32 * k;
      const float re = (float)buf[i].x, im = (float)buf[i].y;
      pw[i] = (i == 0) ? __fmul_rn(re, re) : __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    }
    __syncwarp();
    // MelBanks::Compute (mel-computations.cc:224-235) + log(max(e, eps)) (feature-fbank.cc:102-108)
    for (int b = lane; b < NBIN; b += 32) {
      const int2 rg = S.mel_range[b];
      const float* wv = S.mel_w + S.mel_w_off[b];
      float e = 0.f;
      for (int k = 0; k < rg.y; ++k) e = __fadd_rn(e, __fmul_rn(wv[k], pw[rg.x + k]));
      fb[(size_t)fg * NBIN + b] = logf(fmaxf(e, FLT_EPSILON));
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(FEAT / 4)
lfr_cmvn_posenc_kernel(const float* __restrict__ fb, const int* __restrict__ fb_off, const int* __restrict__ row_seg,
                       const int2* __restrict__ row_info, int M, FrontendTables t, float scale,
                       float* __restrict__ x0, float* __restrict__ feats_tap) {
  const int row = blockIdx.x;
  if (row >= M) return;
  const int c = threadIdx.x * 4;
  const int2 info = row_info[row];
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f), f = o;
  if (info.x >= 0) {
    const int seg = row_seg[row];
    const int f0 = fb_off[seg], n_fb = fb_off[seg + 1] - f0;
    const int j = c / NBIN, m = c - j * NBIN;
    int src = 6 * info.x + j - 3;  // left pad = 3 copies of frame 0, right pad repeats the last frame
    src = src < 0 ? 0 : (src > n_fb - 1 ? n_fb - 1 : src);
    const float4 v = *reinterpret_cast<const float4*>(fb + (size_t)(f0 + src) * NBIN + m);
    const float4 mu = *reinterpret_cast<const float4*>(t.cmvn_mean + c);
    const float4 va = *reinterpret_cast<const float4*>(t.cmvn_var + c);
    f.x = __fmul_rn(__fadd_rn(v.x, mu.x), va.x);
    f.y = __fmul_rn(__fadd_rn(v.y, mu.y), va.y);
    f.z = __fmul_rn(__fadd_rn(v.z, mu.z), va.z);
    f.w = __fmul_rn(__fadd_rn(v.w, mu.w), va.w);
    const int pr = info.x < t.pe_rows ? info.x : t.pe_rows - 1;
    const float4 pe = *reinterpret_cast<const float4*>(t.pos_enc + (size_t)pr * FEAT + c);
    o.x = __fadd_rn(__fmul_rn(f.x, scale), pe.x);
    o.y = __fadd_rn(__fmul_rn(f.y, scale), pe.y);
    o.z = __fadd_rn(__fmul_rn(f.z, scale), pe.z);
    o.w = __fadd_rn(__fmul_rn(f.w, scale), pe.w);
  }
  *reinterpret_cast<float4*>(x0 + (size_t)row * FEAT + c) = o;
  if (feats_tap) *reinterpret_cast<float4*>(feats_tap + (size_t)row * FEAT + c) = f;
}

}  // namespace

int fbank_launch(const void* pcm, int is_f32, const int64_t* sample_off, const int* fb_off, int n_seg,
                 int n_frames_total, const FrontendTables& t, float* fb, cudaStream_t s) {
  if (n_frames_total <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(fbank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbSmem));
    cudaError_t e2 = cudaFuncSetAttribute(fbank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbSmem));
    if (e1 != cudaSuccess) return (int)e1;
    if (e2 != cudaSuccess) return (int)e2;
    attr_set = true;
  }
  int blocks = (n_frames_total + FB_WARPS - 1) / FB_WARPS;
  const int cap = 148 * 2 * 8;
  if (blocks > cap) blocks = cap;
  // total packed mel weights is bounded by 2 * 256; pass the real count through mel_w_off[79] + range
  const int mel_total = 1024;
  if (is_f32)
    fbank_kernel<true><<<blocks, FB_WARPS * 32, sizeof(FbSmem), s>>>(pcm, sample_off, fb_off, n_seg, n_frames_total, t, mel_total, fb);
  else
    fbank_kernel<false><<<blocks, FB_WARPS * 32, sizeof(FbSmem), s>>>(pcm, sample_off, fb_off, n_seg, n_frames_total, t, mel_total, fb);
  return (int)cudaGetLastError();
}

int lfr_cmvn_posenc_launch(const float* fb, const int* fb_off, const int* row_seg, const int2* row_info, int M,
                           const FrontendTables& t, float scale, float* x0, float* feats_tap, cudaStream_t s) {
  if (M <= 0) return 0;
  lfr_cmvn_posenc_kernel<<<M, FEAT / 4, 0, s>>>(fb, fb_off, row_seg, row_info, M, t, scale, x0, feats_tap);
  return (int)cudaGetLastError();
}

}  // namespace pf
