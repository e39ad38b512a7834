// FSMN-VAD scores on the GPU: the network FsmnVad::Forward runs through onnxruntime
// (onnxruntime/src/fsmn-vad.cpp:72-135) plus its front end (FbankKaldi / LfrCmvn, fsmn-vad.cpp:137-224), for WHOLE
// recordings at once.  The reference feeds the model in 1 s chunks and carries four [128 x 19] caches
// (audio.cpp:1183-1196, fsmn-vad.cpp:96-100): with a left-context-only memory block (lorder 20, rorder 0) that equals
// one pass over the whole recording with zero history, which is what this file does (the oracle tests the equivalence).
// The E2E VAD state machine that turns scores into segments (e2e-vad.h) stays on the host; it reads only
// scores[t][sil_pdf_id] with sil_pdf_ids = {0} (e2e-vad.h:602-608), so the default output is that one column.
//
//   fbank (80 mel, shared kernel) -> LFR m=5 n=1 + CMVN -> 400->140 -> 140->250 ReLU -> 4 x { 250->128, causal FSMN(20) + id,
//   128->250 ReLU } -> 250->140 -> 140->248 -> softmax.  GEMMs run on the tcgen05 kernel (gemm.cu) with the odd dimensions
//   zero-padded to multiples of 8 at load time (140 -> 144, 250 -> 256); everything else is a bandwidth-bound kernel here.
#include <math.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "engine.h"
#include "launch.cuh"
#include "model_dir.h"

using namespace pf;

namespace {

constexpr int V_IN = 400, V_A1 = 144, V_LIN = 256, V_PROJ = 128, V_O1 = 144, V_OUT = 248, V_LORDER = 20, V_LAYERS = 4;

// LFR m=5 n=1 + CMVN: X[t][80 j + m] = (fb[clamp(t + j - 2, 0, n_fb - 1)][m] + mean) * var  -> bf16 [rows, 400].
// One thread per 4 consecutive mel bins of one output row (float4 in, 4 bf16 out); 8 rows per block.
__global__ void __launch_bounds__(128)
vad_lfr_cmvn_kernel(const float* __restrict__ fb, const int2* __restrict__ row_info, const int* __restrict__ row_start, int rows,
                    const float* __restrict__ mean, const float* __restrict__ var, __nv_bfloat16* __restrict__ out, float* __restrict__ out_f32) {
  pdl_wait();
  pdl_launch_dependents();
  const int q = threadIdx.x;               // 100 float4 groups per row
  if (q >= V_IN / 4) return;
  const int c = q * 4, j = c / 80, m = c - 80 * j;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), va = *reinterpret_cast<const float4*>(var + c);
  for (int rr = 0; rr < 8; ++rr) {
    const int r = blockIdx.x * 8 + rr;
    if (r >= rows) return;
    const int2 inf = row_info[r];          // {t, T}
    int f = inf.x + j - 2;
    f = f < 0 ? 0 : (f > inf.y - 1 ? inf.y - 1 : f);
    const float4 x = *reinterpret_cast<const float4*>(fb + (size_t)(row_start[r] + f) * 80 + m);
    float4 v;
    v.x = __fmul_rn(__fadd_rn(x.x, mu.x), va.x); v.y = __fmul_rn(__fadd_rn(x.y, mu.y), va.y);
    v.z = __fmul_rn(__fadd_rn(x.z, mu.z), va.z); v.w = __fmul_rn(__fadd_rn(x.w, mu.w), va.w);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + (size_t)r * V_IN + c) = pk;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)r * V_IN + c) = v;
  }
}

// causal FSMN memory block: y[t][c] = x[t][c] + sum_{k<20} w[c][k] x[t - 19 + k][c] (rows before the recording start are zero).
// A block stages 64 output rows plus their 19 predecessors in shared memory (one coalesced pass over x) and every thread owns
// two channels of 8 consecutive rows.
constexpr int VF_ROWS = 64;
__global__ void __launch_bounds__(512)
vad_fsmn_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w_t /*[20][128]*/, const int2* __restrict__ row_info, int rows,
                __nv_bfloat16* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) __nv_bfloat162 tile[VF_ROWS + V_LORDER - 1][V_PROJ / 2];
  __shared__ int t_of[VF_ROWS];
  const int r0 = blockIdx.x * VF_ROWS;
  // 83 rows x 256 B as 16-byte pieces: every thread issues its (up to 3) global loads before the first shared-memory store
  constexpr int kPieces = (VF_ROWS + V_LORDER - 1) * (V_PROJ / 8);
  uint4 piece[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = threadIdx.x + u * 512;
    piece[u] = make_uint4(0u, 0u, 0u, 0u);
    if (i < kPieces) {
      const int rr = i / (V_PROJ / 8), c8 = i - rr * (V_PROJ / 8);
      const int r = r0 - (V_LORDER - 1) + rr;
      if (r >= 0 && r < rows) piece[u] = *reinterpret_cast<const uint4*>(x + (size_t)r * V_PROJ + c8 * 8);
    }
  }
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int i = threadIdx.x + u * 512;
    if (i < kPieces) {
      const int rr = i / (V_PROJ / 8), c8 = i - rr * (V_PROJ / 8);
      *reinterpret_cast<uint4*>(&tile[rr][c8 * 4]) = piece[u];
    }
  }
  for (int i = threadIdx.x; i < VF_ROWS; i += blockDim.x) t_of[i] = r0 + i < rows ? row_info[r0 + i].x : 0;
  __syncthreads();
  const int cc = threadIdx.x & 63, g = threadIdx.x >> 6;     // channel pair, row group (8 groups of 8 rows)
  float2 w[V_LORDER];
#pragma unroll
  for (int k = 0; k < V_LORDER; ++k) w[k] = make_float2(w_t[k * V_PROJ + 2 * cc], w_t[k * V_PROJ + 2 * cc + 1]);
#pragma unroll 2
  for (int i = 0; i < 8; ++i) {
    const int rr = g * 8 + i, r = r0 + rr;
    if (r >= rows) return;
    const int t = t_of[rr];
    float2 acc = __bfloat1622float2(tile[rr + V_LORDER - 1][cc]);
#pragma unroll
    for (int k = 0; k < V_LORDER; ++k) {
      const int back = V_LORDER - 1 - k;   // tap k looks `back` frames into the past
      if (back <= t) {
        const float2 v = __bfloat1622float2(tile[rr + k][cc]);
        acc.x += w[k].x * v.x;
        acc.y += w[k].y * v.y;
      }
    }
    reinterpret_cast<__nv_bfloat162*>(y)[(size_t)r * (V_PROJ / 2) + cc] = __floats2bfloat162_rn(acc.x, acc.y);
  }
}

// row softmax over 248 logits: p0 (silence pdf) and, optionally, the whole row
__global__ void __launch_bounds__(256)
vad_softmax_kernel(const float* __restrict__ logits, int rows, float* __restrict__ p0, float* __restrict__ probs) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* x = logits + (size_t)r * V_OUT;
  float v[8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; v[i] = c < V_OUT ? x[c] : -INFINITY; m = fmaxf(m, v[i]); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = (lane + 32 * i) < V_OUT ? expf(v[i] - m) : 0.f; s += v[i]; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float inv = 1.0f / s;
  if (lane == 0) p0[r] = v[0] * inv;
  if (probs) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; if (c < V_OUT) probs[(size_t)r * V_OUT + c] = v[i] * inv; }
  }
}

uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace

struct b200pf_vad {
  int device = 0, num_sms = 148, max_frames = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  uint8_t* wbase = nullptr;
  FrontendTables ft{};
  float *mean = nullptr, *var = nullptr;
  Linear in1, in2, proj[V_LAYERS], aff[V_LAYERS], out1, out2;
  float* fsmn_w[V_LAYERS] = {nullptr, nullptr, nullptr, nullptr};   // [20][128]
  // workspace
  uint8_t* ws = nullptr;
  void* d_pcm = nullptr;
  float* fb = nullptr;
  __nv_bfloat16 *x400 = nullptr, *a1 = nullptr, *a2 = nullptr, *p = nullptr, *mbuf = nullptr, *o1 = nullptr;
  float *logits = nullptr, *p0 = nullptr, *probs = nullptr, *x400_f32 = nullptr;
  int2* d_row_info = nullptr;
  int* d_row_start = nullptr;
  int64_t* d_sample_off = nullptr;
  int* d_fb_off = nullptr;
  int64_t max_samples = 0;
};

#define VCK(call, what) do { int rc_ = check_cuda((call), what); if (rc_) return rc_; } while (0)

extern "C" {

static int b200pf_vad_create_impl(const char* vad_dir, int device, int max_frames, b200pf_vad** out) {
  if (!vad_dir || !out) { set_error("null argument"); return B200PF_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); set_error("no CUDA device: the B200 path has no CPU fallback"); return B200PF_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) { set_error("bad device index"); return B200PF_ERR_INVALID; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) { set_error("device is not sm_100 (Blackwell B200)"); return B200PF_ERR_NO_DEVICE; }
  VCK(cudaSetDevice(device), "cudaSetDevice");
  const std::string dir(vad_dir);
  std::string err;
  WeightFile wf;
  std::vector<float> means, vars;
  if (!read_weight_file(dir + "/vad.b200pf", &wf, &err) || !read_am_mvn(dir + "/am.mvn", &means, &vars, &err)) { set_error(err); return B200PF_ERR_IO; }
  if ((int)means.size() != V_IN || (int)vars.size() != V_IN) { set_error("vad am.mvn dimension != 400"); return B200PF_ERR_IO; }
  std::unique_ptr<b200pf_vad> v(new b200pf_vad);
  v->device = device;
  v->max_frames = max_frames > 0 ? max_frames : 400000;   // a little over one hour of 10 ms frames
  cudaDeviceGetAttribute(&v->num_sms, cudaDevAttrMultiProcessorCount, device);
  VCK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  VCK(cudaMalloc((void**)&v->wbase, 8 << 20), "cudaMalloc(vad weights)");
  size_t used = 0;
  bool ok = true;
  auto take = [&](size_t bytes) -> void* { size_t o = (used + 255) & ~size_t(255); used = o + bytes; if (used > (8u << 20)) { ok = false; return nullptr; } return v->wbase + o; };
  auto up = [&](const void* h, size_t bytes) -> void* { void* d = take(bytes); if (d && cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) ok = false; return d; };
  // [out, in] fp32 -> zero-padded [out_p, in_p] bf16 (+ zero-padded bias)
  auto linear = [&](const std::string& name, int out_d, int in_d, int out_p, int in_p, bool bias) {
    Linear L;
    L.out = out_p; L.in = in_p;
    auto it = wf.tensors.find(name + ".weight");
    if (it == wf.tensors.end() || it->second.shape != std::vector<int64_t>{out_d, in_d}) { ok = false; err = "missing or misshaped tensor " + name + ".weight"; return L; }
    std::vector<uint16_t> w((size_t)out_p * in_p, 0);
    for (int o = 0; o < out_d; ++o)
      for (int i = 0; i < in_d; ++i) w[(size_t)o * in_p + i] = bf16_rne(it->second.data[(size_t)o * in_d + i]);
    L.w = (__nv_bfloat16*)up(w.data(), w.size() * 2);
    if (bias) {
      auto ib = wf.tensors.find(name + ".bias");
      if (ib == wf.tensors.end() || (int)ib->second.data.size() != out_d) { ok = false; err = "missing tensor " + name + ".bias"; return L; }
      std::vector<float> b(out_p, 0.f);
      memcpy(b.data(), ib->second.data.data(), (size_t)out_d * 4);
      L.b = (float*)up(b.data(), b.size() * 4);
    }
    return L;
  };
  v->in1 = linear("encoder.in_linear1.linear", 140, 400, V_A1, V_IN, true);
  v->in2 = linear("encoder.in_linear2.linear", 250, 140, V_LIN, V_A1, true);
  for (int l = 0; l < V_LAYERS && ok; ++l) {
    const std::string p = "encoder.fsmn." + std::to_string(l);
    v->proj[l] = linear(p + ".linear.linear", 128, 250, V_PROJ, V_LIN, false);
    v->aff[l] = linear(p + ".affine.linear", 250, 128, V_LIN, V_PROJ, true);
    auto it = wf.tensors.find(p + ".fsmn_block.conv_left.weight");
    if (it == wf.tensors.end() || it->second.numel() != 128 * 20) { ok = false; err = "missing tensor " + p + ".fsmn_block.conv_left.weight"; break; }
    std::vector<float> wt((size_t)20 * 128);
    for (int c = 0; c < 128; ++c)
      for (int k = 0; k < 20; ++k) wt[(size_t)k * 128 + c] = it->second.data[(size_t)c * 20 + k];
    v->fsmn_w[l] = (float*)up(wt.data(), wt.size() * 4);
  }
  v->out1 = linear("encoder.out_linear1.linear", 140, 250, V_O1, V_LIN, true);
  v->out2 = linear("encoder.out_linear2.linear", 248, 140, V_OUT, V_O1, true);
  {
    std::vector<float> window, w;
    std::vector<double> tw;
    std::vector<int> range, woff;
    if (!fbank_tables_host(&window, &tw, &range, &w, &woff)) { ok = false; err = "mel table overflow"; }
    else {
      v->ft.window = (const float*)up(window.data(), 400 * 4);
      v->ft.twiddle = (const double2*)up(tw.data(), 512 * 8);
      v->ft.mel_range = (const int2*)up(range.data(), 160 * 4);
      v->ft.mel_w = (const float*)up(w.data(), 1024 * 4);
      v->ft.mel_w_off = (const int*)up(woff.data(), 80 * 4);
    }
    v->mean = (float*)up(means.data(), V_IN * 4);
    v->var = (float*)up(vars.data(), V_IN * 4);
  }
  if (!ok) { set_error(err.empty() ? "vad weight upload failed" : err); cudaFree(v->wbase); cudaStreamDestroy(v->stream); return B200PF_ERR_IO; }

  const size_t R = (size_t)v->max_frames;
  v->max_samples = (int64_t)R * 160 + 400 * 64;
  const size_t bytes = R * (80 * 4 + V_IN * 2 + V_A1 * 2 + V_LIN * 2 + V_PROJ * 2 * 2 + V_O1 * 2 + V_OUT * 4 + 4 + 8 + 4) + (size_t)v->max_samples * 2 + (1 << 20);
  VCK(cudaMalloc((void**)&v->ws, bytes), "cudaMalloc(vad workspace)");
  size_t off = 0;
  auto carve = [&](size_t b) { size_t o = (off + 255) & ~size_t(255); off = o + b; return v->ws + o; };
  v->fb = (float*)carve(R * 80 * 4); v->x400 = (__nv_bfloat16*)carve(R * V_IN * 2); v->a1 = (__nv_bfloat16*)carve(R * V_A1 * 2);
  v->a2 = (__nv_bfloat16*)carve(R * V_LIN * 2); v->p = (__nv_bfloat16*)carve(R * V_PROJ * 2); v->mbuf = (__nv_bfloat16*)carve(R * V_PROJ * 2);
  v->o1 = (__nv_bfloat16*)carve(R * V_O1 * 2); v->logits = (float*)carve(R * V_OUT * 4); v->p0 = (float*)carve(R * 4);
  v->d_row_info = (int2*)carve(R * 8); v->d_row_start = (int*)carve(R * 4); v->d_pcm = carve((size_t)v->max_samples * 2);
  v->d_sample_off = (int64_t*)carve(4097 * 8); v->d_fb_off = (int*)carve(4097 * 4);
  VCK(cudaMemset(v->ws, 0, bytes), "cudaMemset(vad workspace)");
  VCK(cudaDeviceSynchronize(), "vad init");
  *out = v.release();
  return 0;
}
// Parsing a hostile or truncated model directory may throw (std::bad_alloc, std::invalid_argument from the text parsers);
// nothing may unwind through the C ABI: it becomes an error code with the text in b200pf_last_error().
int b200pf_vad_create(const char* vad_dir, int device, int max_frames, b200pf_vad** out) {
  try {
    return b200pf_vad_create_impl(vad_dir, device, max_frames, out);
  } catch (const std::exception& ex) {
    set_error(std::string("b200pf_vad_create: ") + ex.what());
    return B200PF_ERR_IO;
  } catch (...) {
    set_error("b200pf_vad_create: unknown exception");
    return B200PF_ERR_IO;
  }
}

void b200pf_vad_destroy(b200pf_vad* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  cudaStreamSynchronize(v->stream);
  cudaFree(v->wbase);
  cudaFree(v->ws);
  cudaFree(v->probs);
  cudaFree(v->x400_f32);
  cudaStreamDestroy(v->stream);
  delete v;
}

int b200pf_vad_scores_s16(b200pf_vad* v, const int16_t* pcm, const int64_t* offsets, int n_rec, float* sil_prob, int64_t cap_frames,
                          int32_t* frame_off, float* all_probs, float* feats) {
  if (!v || !offsets || n_rec < 0 || (n_rec > 0 && !pcm) || !frame_off) { set_error("bad argument"); return B200PF_ERR_INVALID; }
  if (n_rec > 4096) { set_error("too many recordings in one call"); return B200PF_ERR_CAPACITY; }
  VCK(cudaSetDevice(v->device), "cudaSetDevice");
  std::lock_guard<std::mutex> lock(v->mu);
  const int64_t base = n_rec ? offsets[0] : 0, total = n_rec ? offsets[n_rec] - base : 0;
  if (total > v->max_samples) { set_error("audio exceeds the VAD engine's capacity"); return B200PF_ERR_CAPACITY; }
  std::vector<int64_t> soff(n_rec + 1);
  std::vector<int> fboff(n_rec + 1);
  int frames = 0;
  for (int i = 0; i < n_rec; ++i) {
    const int64_t n = offsets[i + 1] - offsets[i];
    if (n < 0) { set_error("offsets not monotone"); return B200PF_ERR_INVALID; }
    soff[i] = offsets[i] - base;
    fboff[i] = frames;
    frame_off[i] = frames;
    frames += num_fbank_frames(n);   // LFR n = 1: one score per fbank frame
  }
  if (n_rec) { soff[n_rec] = 0; fboff[n_rec] = frames; }
  frame_off[n_rec] = frames;
  if (frames > v->max_frames) { set_error("audio exceeds the VAD engine's frame capacity"); return B200PF_ERR_CAPACITY; }
  if (frames > cap_frames) { set_error("score buffer too small"); return B200PF_ERR_CAPACITY; }
  if (frames == 0) return 0;
  std::vector<int2> info(frames);
  std::vector<int> start(frames);
  for (int i = 0; i < n_rec; ++i) {
    const int T = fboff[i + 1] - fboff[i];
    for (int t = 0; t < T; ++t) { info[fboff[i] + t] = make_int2(t, T); start[fboff[i] + t] = fboff[i]; }
  }
  cudaStream_t s = v->stream;
  VCK(cudaMemcpyAsync(v->d_pcm, pcm + base, (size_t)total * 2, cudaMemcpyHostToDevice, s), "H2D pcm");
  VCK(cudaMemcpyAsync(v->d_sample_off, soff.data(), (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice, s), "H2D");
  VCK(cudaMemcpyAsync(v->d_fb_off, fboff.data(), (size_t)(n_rec + 1) * 4, cudaMemcpyHostToDevice, s), "H2D");
  VCK(cudaMemcpyAsync(v->d_row_info, info.data(), (size_t)frames * 8, cudaMemcpyHostToDevice, s), "H2D");
  VCK(cudaMemcpyAsync(v->d_row_start, start.data(), (size_t)frames * 4, cudaMemcpyHostToDevice, s), "H2D");
  if (all_probs && !v->probs) VCK(cudaMalloc((void**)&v->probs, (size_t)v->max_frames * V_OUT * 4), "cudaMalloc(vad probs)");
  if (feats && !v->x400_f32) VCK(cudaMalloc((void**)&v->x400_f32, (size_t)v->max_frames * V_IN * 4), "cudaMalloc(vad feats)");
  int rc = fbank_launch(v->d_pcm, 0, v->d_sample_off, v->d_fb_off, n_rec, frames, v->ft, v->fb, s);
  if (rc) return check_cuda((cudaError_t)rc, "vad fbank");
  rc = launch_kernel(vad_lfr_cmvn_kernel, dim3((frames + 7) / 8), dim3(128), 0, s, (const float*)v->fb, (const int2*)v->d_row_info, (const int*)v->d_row_start,
                     frames, (const float*)v->mean, (const float*)v->var, v->x400, feats ? v->x400_f32 : (float*)nullptr);
  if (rc) return check_cuda((cudaError_t)rc, "vad lfr");
  auto gemm = [&](const __nv_bfloat16* A, int lda, const Linear& W, int relu, __nv_bfloat16* ob, int ldo, float* of, int ldof) {
    GemmProblem p;
    p.A = A; p.lda = lda; p.rows_a = frames; p.W = W.w; p.ldw = W.in; p.M = frames; p.N = W.out; p.K = W.in;
    GemmEpilogue e;
    e.bias = W.b; e.relu = relu; e.out_bf16 = ob; e.ld_out_bf16 = ldo; e.out_f32 = of; e.ld_out_f32 = ldof;
    return gemm_bf16_tcgen05(p, e, v->num_sms, s);
  };
  if ((rc = gemm(v->x400, V_IN, v->in1, 0, v->a1, V_A1, nullptr, 0))) return check_cuda((cudaError_t)rc, "vad gemm in1");
  if ((rc = gemm(v->a1, V_A1, v->in2, 1, v->a2, V_LIN, nullptr, 0))) return check_cuda((cudaError_t)rc, "vad gemm in2");
  for (int l = 0; l < V_LAYERS; ++l) {
    if ((rc = gemm(v->a2, V_LIN, v->proj[l], 0, v->p, V_PROJ, nullptr, 0))) return check_cuda((cudaError_t)rc, "vad gemm proj");
    rc = launch_kernel(vad_fsmn_kernel, dim3((frames + VF_ROWS - 1) / VF_ROWS), dim3(512), 0, s, (const __nv_bfloat16*)v->p, (const float*)v->fsmn_w[l],
                       (const int2*)v->d_row_info, frames, v->mbuf);
    if (rc) return check_cuda((cudaError_t)rc, "vad fsmn");
    if ((rc = gemm(v->mbuf, V_PROJ, v->aff[l], 1, v->a2, V_LIN, nullptr, 0))) return check_cuda((cudaError_t)rc, "vad gemm affine");
  }
  if ((rc = gemm(v->a2, V_LIN, v->out1, 0, v->o1, V_O1, nullptr, 0))) return check_cuda((cudaError_t)rc, "vad gemm out1");
  if ((rc = gemm(v->o1, V_O1, v->out2, 0, nullptr, 0, v->logits, V_OUT))) return check_cuda((cudaError_t)rc, "vad gemm out2");
  rc = launch_kernel(vad_softmax_kernel, dim3((frames + 7) / 8), dim3(256), 0, s, (const float*)v->logits, frames, v->p0, all_probs ? v->probs : (float*)nullptr);
  if (rc) return check_cuda((cudaError_t)rc, "vad softmax");
  if (sil_prob) VCK(cudaMemcpyAsync(sil_prob, v->p0, (size_t)frames * 4, cudaMemcpyDeviceToHost, s), "D2H scores");
  if (all_probs) VCK(cudaMemcpyAsync(all_probs, v->probs, (size_t)frames * V_OUT * 4, cudaMemcpyDeviceToHost, s), "D2H probs");
  if (feats) VCK(cudaMemcpyAsync(feats, v->x400_f32, (size_t)frames * V_IN * 4, cudaMemcpyDeviceToHost, s), "D2H feats");
  VCK(cudaStreamSynchronize(s), "vad forward");
  return 0;
}

}  // extern "C"
