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

#include <math.h>

#include <vector>

#include "kernels.cuh"
#include "launch.cuh"

namespace pf {
namespace {

constexpr int WIN = 400, SHIFT = 160, NFFT = 512, NBIN = 80, FEAT = 560;
constexpr int FB_WARPS = 8;           // 2 frames per warp (16 lanes each)
constexpr int FB_FRAMES = 2 * FB_WARPS;

// W_16^m = exp(-2 pi i m / 16), m = 0..7 (double)
__constant__ double2 c_w16[8] = {
    {1.0, -0.0},
    {0.92387953251128673848, -0.38268343236508978178},
    {0.70710678118654757274, -0.70710678118654757274},
    {0.38268343236508983729, -0.92387953251128673848},
    {0.0, -1.0},
    {-0.38268343236508972627, -0.92387953251128673848},
    {-0.70710678118654746172, -0.70710678118654757274},
    {-0.92387953251128673848, -0.38268343236508989280}};

struct FbSmem {
  double2 buf[FB_FRAMES][16 * 17];      // per frame: 256 complex doubles as 16 rows of 16 (+1 pad: the transposed read of the
                                        // 16 x 16 exchange, one row per lane, would otherwise be an 8-way bank conflict)
  double2 tw[NFFT / 2];                 // exp(-2 pi i k / 512)
  double2 tw1[NFFT / 2];                // pass-1 twiddles W_256^(sub k1) at [k1][sub]: one conflict-free row per k1
  float window[WIN];
  float power[FB_FRAMES][NFFT / 2 + 8];
  float mel_w[1024];
  int2 mel_range[NBIN];
  int mel_w_off[NBIN];
};

template <bool F32>
__device__ __forceinline__ float load_sample(const void* pcm, int64_t i) {
  if (F32) return __fmul_rn(reinterpret_cast<const float*>(pcm)[i], 32768.0f);  // paraformer.cpp:313
  return (float)reinterpret_cast<const int16_t*>(pcm)[i];                        // == (s/32768.f)*32768 exactly
}

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 16-point DFT in registers, radix-2 decimation in frequency; on return v[bitrev4(k)] holds X[k].
__device__ __forceinline__ void fft16(double2 (&v)[16]) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      if ((a & s) == 0) {
        const int b = a + s;
        const int m = (a & (s - 1)) * (8 / s);
        const double2 t = csub(v[a], v[b]);
        v[a] = cadd(v[a], v[b]);
        v[b] = (m == 0) ? t : cmul(t, c_w16[m]);
      }
    }
  }
}
__host__ __device__ constexpr int brev4(int k) { return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3); }

// One frame per 16 lanes.  The 512-point real FFT is a 256-point complex FFT of z[n] = w[2n] + i w[2n+1]
// (16 x 16 decomposition: one radix-16 pass in registers, twiddle, smem transpose, second radix-16 pass)
// followed by the usual even/odd recombination; all in double like knf's Rfft (rfft.cc:41-47).
template <bool F32>
__global__ void __launch_bounds__(FB_WARPS * 32)
fbank_kernel(const void* __restrict__ pcm, const int64_t* __restrict__ sample_off, const int* __restrict__ fb_off,
             int n_seg, int n_frames, FrontendTables t, int mel_w_total, float* __restrict__ fb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FbSmem& S = *reinterpret_cast<FbSmem*>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int sub = lane & 15;                       // lane within the frame's half-warp
  const unsigned hmask = (lane < 16) ? 0x0000ffffu : 0xffff0000u;
  const int slot = (threadIdx.x >> 5) * 2 + (lane >> 4);
  for (int i = threadIdx.x; i < NFFT / 2; i += blockDim.x) {
    S.tw[i] = t.twiddle[i];
    const int j = 2 * (i & 15) * (i >> 4);           // W_256^m = W_512^(2m); W_512^(j) = -W_512^(j-256)
    const double2 w = t.twiddle[j & 255];
    S.tw1[i] = (j >= 256) ? make_double2(-w.x, -w.y) : w;
  }
  for (int i = threadIdx.x; i < WIN; i += blockDim.x) S.window[i] = t.window[i];
  for (int i = threadIdx.x; i < mel_w_total; i += blockDim.x) S.mel_w[i] = t.mel_w[i];
  for (int i = threadIdx.x; i < NBIN; i += blockDim.x) { S.mel_range[i] = t.mel_range[i]; S.mel_w_off[i] = t.mel_w_off[i]; }
  __syncthreads();
  pdl_wait();   // the tables above are constants; PCM and layout come from the stream
  pdl_launch_dependents();

  double2* buf = S.buf[slot];
  float* pw = S.power[slot];
  const int n_iter = (n_frames + gridDim.x * FB_FRAMES - 1) / (gridDim.x * FB_FRAMES);
  for (int it = 0; it < n_iter; ++it) {
    const int fg = (it * gridDim.x + blockIdx.x) * FB_FRAMES + slot;
    const bool active = fg < n_frames;   // both halves of a warp keep executing the shuffles / syncwarps
    int64_t s0 = 0;
    if (active) {
      int lo = 0, hi = n_seg;  // frame -> segment (largest s with fb_off[s] <= fg)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (fb_off[mid] <= fg) lo = mid; else hi = mid;
      }
      s0 = sample_off[lo] + (int64_t)(fg - fb_off[lo]) * SHIFT;
    }
    // lane `sub` owns samples 32 n1 + 2 sub + {0,1}, n1 = 0..15 (zero beyond 400)
    float xe[13], xo[13];
    float sum = 0.f;
#pragma unroll
    for (int n1 = 0; n1 < 13; ++n1) {
      const int i = 32 * n1 + 2 * sub;
      xe[n1] = (active && i < WIN) ? load_sample<F32>(pcm, s0 + i) : 0.f;
      xo[n1] = (active && i + 1 < WIN) ? load_sample<F32>(pcm, s0 + i + 1) : 0.f;
      sum += xe[n1] + xo[n1];
    }
    // RemoveDcOffset (feature-window.cc:179-190); integer-valued samples make the sum exact in any order
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(hmask, sum, off);
    const float mean = sum / (float)WIN;
    double2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      float we = 0.f, wo = 0.f;
      if (n1 < 13) {
        const int i = 32 * n1 + 2 * sub;
        if (active && i < WIN) {
          // Preemphasize (:200-211): w[i] -= 0.97 w[i-1] on the DC-removed samples, w[0] -= 0.97 w[0]
          const float de = __fsub_rn(xe[n1], mean);
          const float dprev = (i > 0) ? __fsub_rn(load_sample<F32>(pcm, s0 + i - 1), mean) : de;
          we = __fmul_rn(__fsub_rn(de, __fmul_rn(0.97f, dprev)), S.window[i]);
          if (i + 1 < WIN) {
            const float dod = __fsub_rn(xo[n1], mean);
            wo = __fmul_rn(__fsub_rn(dod, __fmul_rn(0.97f, de)), S.window[i + 1]);
          }
        }
      }
      v[n1] = make_double2((double)we, (double)wo);
    }
    // pass 1: DFT over n1 (z index = 16 n1 + sub), twiddle W_256^(sub k1), transpose through smem
    fft16(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const double2 y = v[brev4(k1)];
      buf[k1 * 17 + sub] = (sub * k1 == 0) ? y : cmul(y, S.tw1[k1 * 16 + sub]);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf[sub * 17 + n2];   // lane = k1, values over n2
    __syncwarp();
    // pass 2: DFT over n2 -> Z[k1 + 16 k2]
    fft16(v);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) buf[sub + 17 * k2] = v[brev4(k2)];   // Z[k], k = sub + 16 k2, at row k2, column sub
    __syncwarp();
    // real-FFT recombination: X[k] = (Z[k] + conj Z[256-k])/2 - i W_512^k (Z[k] - conj Z[256-k])/2
    // ComputePowerSpectrum (feature-functions.cc:28-47) on the float-narrowed spectrum
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = sub + 16 * j;
      const int kc = (256 - k) & 255;
      const double2 zk = buf[j * 17 + sub];
      const double2 zc = buf[(kc >> 4) * 17 + (kc & 15)];
      const double2 xe_ = make_double2(0.5 * (zk.x + zc.x), 0.5 * (zk.y - zc.y));
      const double2 xo_ = make_double2(0.5 * (zk.y + zc.y), -0.5 * (zk.x - zc.x));   // -i/2 (Z[k] - conj Z[N-k])
      const double2 w = S.tw[k];
      const double2 x = cadd(xe_, cmul(w, xo_));
      const float re = (float)x.x, im = (float)x.y;
      pw[k] = (k == 0) ? __fmul_rn(re, re) : __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    }
    __syncwarp();
    // MelBanks::Compute (mel-computations.cc:224-235) + log(max(e, eps)) (feature-fbank.cc:102-108)
    if (active) {
      for (int b = sub; b < NBIN; b += 16) {
        const int2 rg = S.mel_range[b];
        const float* wv = S.mel_w + S.mel_w_off[b];
        float e = 0.f;
        for (int k = 0; k < rg.y; ++k) e = __fadd_rn(e, __fmul_rn(wv[k], pw[rg.x + k]));
        fb[(size_t)fg * NBIN + b] = logf(fmaxf(e, FLT_EPSILON));
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(FEAT / 4)
lfr_cmvn_posenc_kernel(const float* __restrict__ fb, const int* __restrict__ fb_off, const int* __restrict__ row_seg,
                       const int2* __restrict__ row_info, int M, FrontendTables t, float scale,
                       float* __restrict__ x0, float* __restrict__ feats_tap) {
  pdl_wait();
  pdl_launch_dependents();
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
    float4 pe;
    if (info.x < t.pe_rows) {
      pe = *reinterpret_cast<const float4*>(t.pos_enc + (size_t)info.x * FEAT + c);
    } else {
      // beyond the table (a recording decoded as ONE segment, > 2048 frames = 123 s): the same formula the table was built with
      // (paraformer-online.cpp:240-268: positions from 1, timescales exp(-i ln(1e4) / (depth/2 - 1)), sines then cosines)
      constexpr int half = FEAT / 2;
      const float inc = (float)(-(log(10000.0) / (half - 1)));
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c + q, i = col < half ? col : col - half;
        const float st = (float)(info.x + 1) * expf((float)i * inc);
        v[q] = col < half ? sinf(st) : cosf(st);
      }
      pe = make_float4(v[0], v[1], v[2], v[3]);
    }
    o.x = __fadd_rn(__fmul_rn(f.x, scale), pe.x);
    o.y = __fadd_rn(__fmul_rn(f.y, scale), pe.y);
    o.z = __fadd_rn(__fmul_rn(f.z, scale), pe.z);
    o.w = __fadd_rn(__fmul_rn(f.w, scale), pe.w);
  }
  *reinterpret_cast<float4*>(x0 + (size_t)row * FEAT + c) = o;
  if (feats_tap) *reinterpret_cast<float4*>(feats_tap + (size_t)row * FEAT + c) = f;
}

}  // namespace

// Host-side tables of the Kaldi fbank front end (shared by the acoustic model and the VAD engine): hamming window [400],
// forward twiddles of the 512-point FFT [256] as (cos, sin) doubles, and the 80 triangular mel filters packed as
// {first_bin, size} ranges + weights (at most 1024 weights).
bool fbank_tables_host(std::vector<float>* window_o, std::vector<double>* tw_o, std::vector<int>* range_o, std::vector<float>* w_o,
                       std::vector<int>* woff_o) {
  std::vector<float> window(400);
  const double a = (2.0 * M_PI) / 399.0;
  for (int i = 0; i < 400; ++i) window[i] = (float)(0.54 - 0.46 * cos(a * (double)i));
  std::vector<double> tw(512);
  for (int k = 0; k < 256; ++k) { tw[2 * k] = cos(-2.0 * M_PI * k / 512.0); tw[2 * k + 1] = sin(-2.0 * M_PI * k / 512.0); }
  auto mel = [](float f) { return 1127.0f * logf(1.0f + f / 700.0f); };
  const float sample_freq = 16000.0f, nyquist = 0.5f * sample_freq;
  const float fft_bin_width = sample_freq / 512;
  const float mel_low = mel(20.0f), mel_high = mel(nyquist + 0.0f);
  const float delta = (mel_high - mel_low) / (80 + 1);
  std::vector<int> range(160), woff(80);
  std::vector<float> w(1024, 0.f);
  int used = 0;
  for (int bin = 0; bin < 80; ++bin) {
    const float left = mel_low + bin * delta, center = mel_low + (bin + 1) * delta, right = mel_low + (bin + 2) * delta;
    int first = -1, last = -1;
    std::vector<float> tb(256, 0.f);
    for (int i = 0; i < 256; ++i) {
      const float freq = fft_bin_width * i;
      const float m = mel(freq);
      if (m > left && m < right) {
        float wt;
        if (m <= center) wt = (m - left) / (center - left);
        else wt = (right - m) / (right - center);
        tb[i] = wt;
        if (first == -1) first = i;
        last = i;
      }
    }
    const int size = last + 1 - first;
    if (first < 0 || used + size > 1024) return false;
    range[2 * bin] = first; range[2 * bin + 1] = size; woff[bin] = used;
    for (int k = 0; k < size; ++k) w[used + k] = tb[first + k];
    used += size;
  }
  window_o->swap(window); tw_o->swap(tw); range_o->swap(range); w_o->swap(w); woff_o->swap(woff);
  return true;
}

int fbank_launch(const void* pcm, int is_f32, const int64_t* sample_off, const int* fb_off, int n_seg,
                 int n_frames_total, const FrontendTables& t, float* fb, cudaStream_t s) {
  if (n_frames_total <= 0) return 0;
  static PerDeviceOnce once;
  const int rc = once_per_device(once, [] {
    cudaError_t err = cudaFuncSetAttribute(fbank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbSmem));
    if (err == cudaSuccess) err = cudaFuncSetAttribute(fbank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbSmem));
    return (int)err;
  });
  if (rc) return rc;
  int blocks = (n_frames_total + FB_FRAMES - 1) / FB_FRAMES;
  const int cap = 148 * 2 * 8;
  if (blocks > cap) blocks = cap;
  // total packed mel weights is bounded by 2 * 256; pass the real count through mel_w_off[79] + range
  const int mel_total = 1024;
  if (is_f32)
    return launch_kernel(fbank_kernel<true>, dim3(blocks), dim3(FB_WARPS * 32), sizeof(FbSmem), s, pcm, sample_off, fb_off, n_seg,
                         n_frames_total, t, mel_total, fb);
  return launch_kernel(fbank_kernel<false>, dim3(blocks), dim3(FB_WARPS * 32), sizeof(FbSmem), s, pcm, sample_off, fb_off, n_seg,
                       n_frames_total, t, mel_total, fb);
}

int lfr_cmvn_posenc_launch(const float* fb, const int* fb_off, const int* row_seg, const int2* row_info, int M,
                           const FrontendTables& t, float scale, float* x0, float* feats_tap, cudaStream_t s) {
  if (M <= 0) return 0;
  return launch_kernel(lfr_cmvn_posenc_kernel, dim3(M), dim3(FEAT / 4), 0, s, fb, fb_off, row_seg, row_info, M, t, scale, x0, feats_tap);
}

}  // namespace pf
