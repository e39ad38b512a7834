/*
 * ORACLE (test infrastructure only — never linked or called by the product path).
 *
 * Plain-C restatement of the reference's deterministic front end and greedy back end for
 * Paraformer::Forward.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/onnxruntime).
 *
 *   pf_oracle_fbank      Paraformer::FbankKaldi            src/paraformer.cpp:309-323
 *                        knf::OnlineFbank                  third_party/kaldi-native-fbank/kaldi-native-fbank/csrc/
 *                            NumFrames                     feature-window.cc:73-118
 *                            ExtractWindow/ProcessWindow   feature-window.cc:121-245
 *                            FeatureWindowFunction         feature-window.cc:25-55
 *                            Rfft (double)                 rfft.cc:35-62
 *                            ComputePowerSpectrum          feature-functions.cc:28-47
 *                            MelBanks ctor / Compute       mel-computations.cc:107-255
 *                            FbankComputer::Compute        feature-fbank.cc:73-118
 *   pf_oracle_lfr_cmvn   Paraformer::LfrCmvn               src/paraformer.cpp:421-461
 *   pf_oracle_find_max   FindMax                           src/util.cpp:63-74
 *
 * Pinning: checked against the reference's own compiled knf (oracle/_ref/libknf_ref.so, built by
 * oracle/Makefile from the sources where they lie) and against knf's test-rfft.cc:32-50 known answer.
 * The FFT here is a textbook double-precision radix-2 transform, not Ooura's rdft; both are exact to
 * ~1e-13 relative before the narrowing to float, so results agree to float rounding.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PF_WIN 400
#define PF_SHIFT 160
#define PF_NFFT 512
#define PF_NBIN 80
#define PF_LFR_M 7
#define PF_LFR_N 6

/* feature-window.cc:73-87 (snip_edges == true branch) */
int pf_oracle_num_fbank_frames(int64_t n) {
  if (n < PF_WIN) return 0;
  return (int)(1 + (n - PF_WIN) / PF_SHIFT);
}

/* paraformer.cpp:424  T_lrf = ceil(1.0 * T / lfr_n) */
int pf_oracle_num_lfr_frames(int n_fb) { return (int)ceil(1.0 * n_fb / PF_LFR_N); }

/* double radix-2 DIT FFT of a real 512-vector; out_re/out_im hold bins 0..256 */
static void rfft512_double(const float *in, double *out_re, double *out_im) {
  static double re[PF_NFFT], im[PF_NFFT];
  int n = PF_NFFT, logn = 9;
  for (int i = 0; i < n; i++) {
    int r = 0;
    for (int b = 0; b < logn; b++) r |= ((i >> b) & 1) << (logn - 1 - b);
    re[r] = (double)in[i];
    im[r] = 0.0;
  }
  for (int len = 2; len <= n; len <<= 1) {
    double ang = -2.0 * M_PI / len;
    for (int i = 0; i < n; i += len) {
      for (int j = 0; j < len / 2; j++) {
        double wr = cos(ang * j), wi = sin(ang * j);
        double ur = re[i + j], ui = im[i + j];
        double vr = re[i + j + len / 2] * wr - im[i + j + len / 2] * wi;
        double vi = re[i + j + len / 2] * wi + im[i + j + len / 2] * wr;
        re[i + j] = ur + vr; im[i + j] = ui + vi;
        re[i + j + len / 2] = ur - vr; im[i + j + len / 2] = ui - vi;
      }
    }
  }
  for (int k = 0; k <= n / 2; k++) { out_re[k] = re[k]; out_im[k] = im[k]; }
}

/* generic small real FFT used only by the known-answer test (knf test-rfft.cc:32-50).
 * Output uses knf's packed layout [Re0, Re(n/2), Re1, -Im1(sign as Ooura: +), ...]; the test only
 * checks magnitudes of the imaginary parts via the sign convention stated there. */
void pf_oracle_rfft_packed(const float *in, int n, float *out) {
  for (int k = 0; k <= n / 2; k++) {
    double sr = 0, si = 0;
    for (int t = 0; t < n; t++) {
      double a = -2.0 * M_PI * k * t / n;
      sr += in[t] * cos(a); si += in[t] * sin(a);
    }
    if (k == 0) out[0] = (float)sr;
    else if (k == n / 2) out[1] = (float)sr;
    else { out[2 * k] = (float)sr; out[2 * k + 1] = (float)(-si); /* Ooura rdft: a[2k+1] = -Im */ }
  }
}

/* mel-computations.h:72-74 */
static float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

typedef struct { int first; int size; float w[PF_NFFT / 2]; } mel_bin_t;
static mel_bin_t g_bins[PF_NBIN];
static float g_window[PF_WIN];
static int g_init = 0;

/* mel-computations.cc:107-200 and feature-window.cc:25-55 with the options set in
 * paraformer.cpp:24-31 (16 kHz, 25/10 ms, hamming, 80 bins, low 20 Hz, high = Nyquist). */
static void init_tables(void) {
  if (g_init) return;
  double a = (2.0 * M_PI) / (PF_WIN - 1);
  for (int i = 0; i < PF_WIN; i++) g_window[i] = (float)(0.54 - 0.46 * cos(a * (double)i));
  float sample_freq = 16000.0f;
  int num_fft_bins = PF_NFFT / 2;
  float nyquist = 0.5f * sample_freq;
  float low_freq = 20.0f, high_freq = nyquist + 0.0f;
  float fft_bin_width = sample_freq / PF_NFFT;
  float mel_low = mel_scale(low_freq), mel_high = mel_scale(high_freq);
  float delta = (mel_high - mel_low) / (PF_NBIN + 1);
  for (int bin = 0; bin < PF_NBIN; bin++) {
    float left = mel_low + bin * delta, center = mel_low + (bin + 1) * delta,
          right = mel_low + (bin + 2) * delta;
    float this_bin[PF_NFFT / 2];
    memset(this_bin, 0, sizeof(this_bin));
    int first = -1, last = -1;
    for (int i = 0; i < num_fft_bins; i++) {
      float freq = fft_bin_width * i;
      float mel = mel_scale(freq);
      if (mel > left && mel < right) {
        float w;
        if (mel <= center) w = (mel - left) / (center - left);
        else w = (right - mel) / (right - center);
        this_bin[i] = w;
        if (first == -1) first = i;
        last = i;
      }
    }
    g_bins[bin].first = first;
    g_bins[bin].size = last + 1 - first;
    memcpy(g_bins[bin].w, this_bin + first, sizeof(float) * g_bins[bin].size);
  }
  g_init = 1;
}

/* Dense [80][256] view of the mel weights (what the CUDA side uploads is built by its own host
 * code; this export exists so tests can compare the two tables bit for bit). */
void pf_oracle_mel_matrix(float *out /* [80][256] */) {
  init_tables();
  memset(out, 0, sizeof(float) * PF_NBIN * (PF_NFFT / 2));
  for (int b = 0; b < PF_NBIN; b++)
    for (int k = 0; k < g_bins[b].size; k++) out[b * (PF_NFFT / 2) + g_bins[b].first + k] = g_bins[b].w[k];
}
void pf_oracle_window(float *out /* [400] */) { init_tables(); memcpy(out, g_window, sizeof(g_window)); }

/* pcm: float in [-1,1) as produced by Audio::LoadPcmwav (audio.cpp:787-819).
 * out: [n_fb][80].  Returns n_fb. */
int pf_oracle_fbank(const float *pcm, int64_t n, float *out) {
  init_tables();
  int n_fb = pf_oracle_num_fbank_frames(n);
  float w[PF_NFFT];
  double fre[PF_NFFT / 2 + 1], fim[PF_NFFT / 2 + 1];
  float pw[PF_NFFT / 2 + 1];
  for (int f = 0; f < n_fb; f++) {
    const float *src = pcm + (int64_t)f * PF_SHIFT;
    /* paraformer.cpp:312-314 */
    for (int i = 0; i < PF_WIN; i++) w[i] = src[i] * 32768;
    for (int i = PF_WIN; i < PF_NFFT; i++) w[i] = 0.0f;
    /* RemoveDcOffset feature-window.cc:179-190 */
    float sum = 0;
    for (int i = 0; i != PF_WIN; ++i) sum += w[i];
    float mean = sum / PF_WIN;
    for (int i = 0; i != PF_WIN; ++i) w[i] -= mean;
    /* Preemphasize feature-window.cc:200-211 */
    for (int i = PF_WIN - 1; i > 0; --i) w[i] -= 0.97f * w[i - 1];
    w[0] -= 0.97f * w[0];
    /* window feature-window.cc:57-63 */
    for (int i = 0; i != PF_WIN; ++i) w[i] *= g_window[i];
    /* rfft in double, narrowed to float: rfft.cc:41-47 */
    rfft512_double(w, fre, fim);
    /* ComputePowerSpectrum feature-functions.cc:28-47 */
    {
      float r0 = (float)fre[0];
      pw[0] = r0 * r0;
      for (int k = 1; k < PF_NFFT / 2; k++) {
        float real = (float)fre[k], im = (float)fim[k];
        pw[k] = real * real + im * im;
      }
      float rl = (float)fre[PF_NFFT / 2];
      pw[PF_NFFT / 2] = rl * rl;
    }
    /* MelBanks::Compute mel-computations.cc:224-235 + log feature-fbank.cc:102-108 */
    for (int b = 0; b < PF_NBIN; b++) {
      float energy = 0;
      for (int k = 0; k != g_bins[b].size; ++k) energy += g_bins[b].w[k] * pw[k + g_bins[b].first];
      float t = energy > FLT_EPSILON ? energy : FLT_EPSILON;
      out[(int64_t)f * PF_NBIN + b] = logf(t);
    }
  }
  return n_fb;
}

/* Paraformer::LfrCmvn paraformer.cpp:421-461.  fb: [n_fb][80] -> out [T][560], returns T.
 * means/vars: 560 each (am.mvn <AddShift>/<Rescale>, LoadCmvn paraformer.cpp:325-360). */
int pf_oracle_lfr_cmvn(const float *fb, int n_fb, const float *means, const float *vars, float *out) {
  if (n_fb <= 0) return 0;
  int T_lfr = pf_oracle_num_lfr_frames(n_fb);
  int left = (PF_LFR_M - 1) / 2;
  int T = n_fb + left; /* length after left padding with copies of frame 0 */
  for (int i = 0; i < T_lfr; i++) {
    for (int j = 0; j < PF_LFR_M; j++) {
      int idx = i * PF_LFR_N + j; /* index into the left-padded sequence */
      if (idx >= T) idx = T - 1;   /* right pad by repeating the last frame */
      int srcf = idx - left; if (srcf < 0) srcf = 0;
      memcpy(out + ((int64_t)i * PF_LFR_M + j) * PF_NBIN, fb + (int64_t)srcf * PF_NBIN, sizeof(float) * PF_NBIN);
    }
    float *row = out + (int64_t)i * PF_LFR_M * PF_NBIN;
    for (int c = 0; c < PF_LFR_M * PF_NBIN; c++) row[c] = (row[c] + means[c]) * vars[c];
  }
  return T_lfr;
}

/* FindMax util.cpp:63-74 — first maximum wins (strict >). */
void pf_oracle_find_max(const float *din, int len, float *max_val, int *max_idx) {
  *max_val = -INFINITY; *max_idx = -1;
  for (int i = 0; i < len; i++) if (din[i] > *max_val) { *max_val = din[i]; *max_idx = i; }
}
