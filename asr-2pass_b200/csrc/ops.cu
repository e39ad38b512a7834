// Single-operator entry points of include/b200pf.h (b200pf_op_*): fp32 host buffers in and out, the product
// kernels in the middle.  They exist so the GPU parity tests can check every kernel against the CPU oracle
// in isolation; the serving path never calls them.
#include <cuda_fp16.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "engine.h"

using namespace pf;

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) { return check_cuda(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc"); }
  template <class T> T* as() { return (T*)p; }
};

// 16-bit operand format the single-operator entry points run in (b200pf_op_set_precision); test infrastructure only
int g_op_f16 = 1;

int select_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device: the B200 path has no CPU fallback");
    return B200PF_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { set_error("bad device index"); return B200PF_ERR_INVALID; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) { set_error("device is not sm_100"); return B200PF_ERR_NO_DEVICE; }
  return check_cuda(cudaSetDevice(device), "cudaSetDevice");
}

// upload fp32 host -> bf16 device
int up_bf16(const float* h, size_t n, DevBuf* tmp, DevBuf* out) {
  int rc = tmp->alloc(n * 4); if (rc) return rc;
  rc = out->alloc(n * 2); if (rc) return rc;
  rc = check_cuda(cudaMemcpy(tmp->p, h, n * 4, cudaMemcpyHostToDevice), "H2D"); if (rc) return rc;
  rc = f32_to_bf16_launch(tmp->as<float>(), out->as<__nv_bfloat16>(), (int64_t)n, 0, g_op_f16);
  return rc ? check_cuda((cudaError_t)rc, "f32_to_bf16") : 0;
}
int up_f32(const float* h, size_t n, DevBuf* out) {
  int rc = out->alloc(n * 4); if (rc) return rc;
  return check_cuda(cudaMemcpy(out->p, h, n * 4, cudaMemcpyHostToDevice), "H2D");
}
template <class T>
int up_raw(const T* h, size_t n, DevBuf* out) {
  int rc = out->alloc(n * sizeof(T)); if (rc) return rc;
  return check_cuda(cudaMemcpy(out->p, h, n * sizeof(T), cudaMemcpyHostToDevice), "H2D");
}

// 16-bit results back to fp32 in the format the ops run in
void widen(const __nv_bfloat16* in, float* out, int64_t n) { h16_to_f32_launch(in, out, n, 0, g_op_f16); }

// pseudo-random bf16 in about [-1, 1): the micro-benchmark must toggle the tensor-core datapath like real activations do
// (constant operands draw far less power and flatter the clock, hence the TFLOP/s)
__global__ void fill_random_h16_kernel(__nv_bfloat16* out, int64_t n, uint32_t seed, int f16) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    const float v = ((float)(x >> 8) * (1.0f / 8388608.0f)) - 1.0f;
    if (f16) reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
    else out[i] = __float2bfloat16(v);
  }
}

int sync_ok(const char* what) { return check_cuda(cudaDeviceSynchronize(), what); }

int sm_count() { int n = 148, d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }

}  // namespace

#define RC(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

extern "C" {

int b200pf_op_set_precision(int precision) {
  if (precision != B200PF_PREC_BF16 && precision != B200PF_PREC_FP16) { set_error("precision must be 0 (bf16) or 1 (fp16)"); return B200PF_ERR_INVALID; }
  g_op_f16 = precision == B200PF_PREC_FP16 ? 1 : 0;
  return 0;
}

int b200pf_op_gemm(int device, const float* A, const float* W, const float* bias, const float* add, const float* res,
                   int M, int N, int K, int relu, int out_bf16_round, float* out, int32_t* argmax_out) {
  RC(select_device(device));
  if (K % 8 || N % 4) { set_error("op_gemm: K % 8 and N % 4 required"); return B200PF_ERR_INVALID; }
  DevBuf ta, tw, dA, dW, dBias, tadd, dAdd, dRes, dOut, dOutB, dAm, dIds;
  RC(up_bf16(A, (size_t)M * K, &ta, &dA));
  RC(up_bf16(W, (size_t)N * K, &tw, &dW));
  GemmProblem p;
  p.f16 = g_op_f16;
  p.A = dA.as<__nv_bfloat16>(); p.lda = K; p.rows_a = M; p.W = dW.as<__nv_bfloat16>(); p.ldw = K; p.M = M; p.N = N; p.K = K;
  GemmEpilogue e;
  if (bias) { RC(up_f32(bias, N, &dBias)); e.bias = dBias.as<float>(); }
  if (add) { RC(up_bf16(add, (size_t)M * N, &tadd, &dAdd)); e.add_bf16 = dAdd.as<__nv_bfloat16>(); e.ld_add = N; }
  e.relu = relu;
  RC(dOut.alloc((size_t)M * N * 4));
  // out_bf16_round: 0 = fp32 out, with `res` IN PLACE (x += ..., the form every residual GEMM of the forward has: TMA
  // reduce-add epilogue); 1 = bf16 out; 2 = fp32 out through the general epilogue with a separate residual buffer
  if (res && out_bf16_round == 0) {
    RC(check_cuda(cudaMemcpy(dOut.p, res, (size_t)M * N * 4, cudaMemcpyHostToDevice), "H2D"));
    e.res_f32 = dOut.as<float>(); e.ld_res = N;
  } else if (res) { RC(up_f32(res, (size_t)M * N, &dRes)); e.res_f32 = dRes.as<float>(); e.ld_res = N; }
  if (out_bf16_round == 2) { e.dbg = 4; out_bf16_round = 0; }
  if (out_bf16_round) { RC(dOutB.alloc((size_t)M * N * 2)); e.out_bf16 = dOutB.as<__nv_bfloat16>(); e.ld_out_bf16 = N; }
  else { e.out_f32 = dOut.as<float>(); e.ld_out_f32 = N; }
  if (argmax_out) {
    RC(dAm.alloc((size_t)M * 8)); RC(dIds.alloc((size_t)M * 4));
    RC(check_cuda(cudaMemset(dAm.p, 0, (size_t)M * 8), "memset"));
    e.argmax = dAm.as<unsigned long long>();
  }
  int rc = gemm_bf16_tcgen05(p, e, sm_count(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "gemm launch");
  if (out_bf16_round) widen(dOutB.as<__nv_bfloat16>(), dOut.as<float>(), (int64_t)M * N);
  if (argmax_out) {
    rc = argmax_decode_launch(dAm.as<unsigned long long>(), nullptr, M, dIds.as<int>(), 0);
    if (rc) return check_cuda((cudaError_t)rc, "argmax decode");
  }
  RC(sync_ok("op_gemm"));
  RC(check_cuda(cudaMemcpy(out, dOut.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost), "D2H"));
  if (argmax_out) RC(check_cuda(cudaMemcpy(argmax_out, dIds.p, (size_t)M * 4, cudaMemcpyDeviceToHost), "D2H"));
  return 0;
}

int b200pf_op_gemm_bench(int device, int M, int N, int K, int mode, int iters, float* ms_out) {
  RC(select_device(device));
  DevBuf dA, dW, dBias, dAdd, dX, dOutB, dAm;
  RC(dA.alloc((size_t)M * K * 2)); RC(dW.alloc((size_t)N * K * 2)); RC(dBias.alloc((size_t)N * 4));
  RC(dAdd.alloc((size_t)M * N * 2)); RC(dX.alloc((size_t)M * N * 4)); RC(dOutB.alloc((size_t)M * N * 2)); RC(dAm.alloc((size_t)M * 8));
  if (getenv("B200PF_GEMM_CONST")) {  // 0x3c3c... is bf16 0.0115: constant operands, for comparison only
    RC(check_cuda(cudaMemset(dA.p, 0x3c, (size_t)M * K * 2), "memset")); RC(check_cuda(cudaMemset(dW.p, 0x3c, (size_t)N * K * 2), "memset"));
    RC(check_cuda(cudaMemset(dAdd.p, 0x3c, (size_t)M * N * 2), "memset"));
  } else {
    fill_random_h16_kernel<<<1024, 256>>>(dA.as<__nv_bfloat16>(), (int64_t)M * K, 1u, g_op_f16);
    fill_random_h16_kernel<<<1024, 256>>>(dW.as<__nv_bfloat16>(), (int64_t)N * K, 2u, g_op_f16);
    fill_random_h16_kernel<<<1024, 256>>>(dAdd.as<__nv_bfloat16>(), (int64_t)M * N, 3u, g_op_f16);
  }
  RC(check_cuda(cudaMemset(dBias.p, 0, (size_t)N * 4), "memset"));
  RC(check_cuda(cudaMemset(dX.p, 0, (size_t)M * N * 4), "memset")); RC(check_cuda(cudaMemset(dAm.p, 0, (size_t)M * 8), "memset"));
  GemmProblem p;
  p.f16 = g_op_f16;
  p.A = dA.as<__nv_bfloat16>(); p.lda = K; p.rows_a = M; p.W = dW.as<__nv_bfloat16>(); p.ldw = K; p.M = M; p.N = N; p.K = K;
  GemmEpilogue e;
  e.bias = dBias.as<float>();
  if (mode == 0 || mode == 1) { e.out_bf16 = dOutB.as<__nv_bfloat16>(); e.ld_out_bf16 = N; e.relu = mode; }
  if (mode == 2 || mode == 3) { e.res_f32 = dX.as<float>(); e.ld_res = N; e.out_f32 = dX.as<float>(); e.ld_out_f32 = N; }
  if (mode == 3) { e.add_bf16 = dAdd.as<__nv_bfloat16>(); e.ld_add = N; }
  if (mode == 4) e.argmax = dAm.as<unsigned long long>();
  if (const char* d = getenv("B200PF_GEMM_DBG")) e.dbg = atoi(d);
  const int sms = sm_count();
  for (int i = 0; i < 3; ++i) { int rc = gemm_bf16_tcgen05(p, e, sms, 0); if (rc) return check_cuda((cudaError_t)rc, "gemm launch"); }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, 0);
  for (int i = 0; i < iters; ++i) { int rc = gemm_bf16_tcgen05(p, e, sms, 0); if (rc) return check_cuda((cudaError_t)rc, "gemm launch"); }
  cudaEventRecord(b, 0);
  RC(sync_ok("op_gemm_bench"));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (ms_out) *ms_out = ms / iters;
  return 0;
}

int b200pf_op_conv3(int device, const float* X, const float* Wr, const float* bias, int M, int C, float* out) {
  RC(select_device(device));
  DevBuf tx, tw, dX, dW, dB, dOut;
  RC(up_bf16(X, (size_t)M * C, &tx, &dX));
  RC(up_bf16(Wr, (size_t)C * 3 * C, &tw, &dW));
  RC(up_f32(bias, C, &dB));
  RC(dOut.alloc((size_t)M * C * 4));
  GemmProblem p;
  p.f16 = g_op_f16;
  p.A = dX.as<__nv_bfloat16>(); p.lda = C; p.rows_a = M; p.W = dW.as<__nv_bfloat16>(); p.ldw = 3 * C; p.M = M; p.N = C; p.K = 3 * C;
  p.a_k_wrap = C; p.a_row_shift0 = -1;
  GemmEpilogue e;
  e.bias = dB.as<float>(); e.out_f32 = dOut.as<float>(); e.ld_out_f32 = C;
  int rc = gemm_bf16_tcgen05(p, e, sm_count(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "conv3 launch");
  RC(sync_ok("op_conv3"));
  return check_cuda(cudaMemcpy(out, dOut.p, (size_t)M * C * 4, cudaMemcpyDeviceToHost), "D2H");
}

int b200pf_op_layernorm(int device, const float* x, int rows, int D, const float* gamma, const float* beta, float eps,
                        int in_bf16, float* out_f32, float* out_bf16_as_f32) {
  RC(select_device(device));
  DevBuf tx, dX, dXb, dG, dB, dO, dOb, dOb32;
  RC(up_f32(gamma, D, &dG)); RC(up_f32(beta, D, &dB));
  const size_t n = (size_t)rows * D;
  RC(dO.alloc(n * 4)); RC(dOb.alloc(n * 2)); RC(dOb32.alloc(n * 4));
  int rc;
  if (in_bf16) {
    RC(up_bf16(x, n, &tx, &dXb));
    rc = layernorm_launch(dXb.p, 1, rows, nullptr, D, dG.as<float>(), dB.as<float>(), eps, dOb.as<__nv_bfloat16>(), dO.as<float>(), nullptr, 0, 0, g_op_f16);
  } else {
    RC(up_f32(x, n, &dX));
    rc = layernorm_launch(dX.p, 0, rows, nullptr, D, dG.as<float>(), dB.as<float>(), eps, dOb.as<__nv_bfloat16>(), dO.as<float>(), nullptr, 0, 0, g_op_f16);
  }
  if (rc) return check_cuda((cudaError_t)rc, "layernorm launch");
  widen(dOb.as<__nv_bfloat16>(), dOb32.as<float>(), (int64_t)n);
  RC(sync_ok("op_layernorm"));
  if (out_f32) RC(check_cuda(cudaMemcpy(out_f32, dO.p, n * 4, cudaMemcpyDeviceToHost), "D2H"));
  if (out_bf16_as_f32) RC(check_cuda(cudaMemcpy(out_bf16_as_f32, dOb32.p, n * 4, cudaMemcpyDeviceToHost), "D2H"));
  return 0;
}

int b200pf_op_attention(int device, const float* q, const float* k, const float* v, const int32_t* q_off,
                        const int32_t* q_len, const int32_t* kv_off, const int32_t* kv_len, int n_seg, int n_heads,
                        int64_t q_rows, int64_t kv_rows, int impl, float* out) {
  RC(select_device(device));
  const int D = n_heads * 128;
  // pack k|v into one [kv_rows, 2D] buffer like the engine's cross-attention layout
  std::vector<float> kvh((size_t)kv_rows * 2 * D);
  for (int64_t r = 0; r < kv_rows; ++r) {
    memcpy(&kvh[(size_t)r * 2 * D], k + (size_t)r * D, (size_t)D * 4);
    memcpy(&kvh[(size_t)r * 2 * D + D], v + (size_t)r * D, (size_t)D * 4);
  }
  DevBuf tq, tkv, dQ, dKV, dOutB, dOut, dqo, dql, dko, dkl, dWork;
  RC(up_bf16(q, (size_t)q_rows * D, &tq, &dQ));
  RC(up_bf16(kvh.data(), kvh.size(), &tkv, &dKV));
  RC(dOutB.alloc((size_t)q_rows * D * 2)); RC(dOut.alloc((size_t)q_rows * D * 4));
  RC(check_cuda(cudaMemset(dOutB.p, 0, (size_t)q_rows * D * 2), "memset"));
  RC(up_raw(q_off, n_seg, &dqo)); RC(up_raw(q_len, n_seg, &dql)); RC(up_raw(kv_off, n_seg, &dko)); RC(up_raw(kv_len, n_seg, &dkl));
  std::vector<AttnWork> work;
  for (int s = 0; s < n_seg; ++s)
    for (int q0 = 0; q0 < q_len[s]; q0 += 128) work.push_back(AttnWork{s, q0});
  RC(up_raw(work.data(), work.size(), &dWork));
  AttnProblem p;
  p.f16 = g_op_f16; p.num_sms = sm_count();
  p.q = dQ.as<__nv_bfloat16>(); p.q_rows = q_rows; p.ldq = D; p.q_col0 = 0;
  p.kv = dKV.as<__nv_bfloat16>(); p.kv_rows = kv_rows; p.ldkv = 2 * D; p.k_col0 = 0; p.v_col0 = D;
  p.out = dOutB.as<__nv_bfloat16>(); p.ldo = D;
  p.q_row_off = dqo.as<int>(); p.q_len = dql.as<int>(); p.kv_row_off = dko.as<int>(); p.kv_len = dkl.as<int>();
  p.work = dWork.as<AttnWork>(); p.n_work = (int)work.size(); p.n_heads = n_heads;
  int rc = impl == 1 ? attention_check_kernel(p, 0) : attention_tcgen05(p, 0);
  if (rc) return check_cuda((cudaError_t)rc, "attention launch");
  widen(dOutB.as<__nv_bfloat16>(), dOut.as<float>(), (int64_t)q_rows * D);
  RC(sync_ok("op_attention"));
  return check_cuda(cudaMemcpy(out, dOut.p, (size_t)q_rows * D * 4, cudaMemcpyDeviceToHost), "D2H");
}

// Times the attention kernel alone on device-resident random operands in the engine's layouts: segments of seg_T[i] frames
// (+ one gap row each).  cross = 0: self-attention over a fused [M,1536] QKV buffer; cross = 1: Lq = (T+1)/2 query rows per
// segment against a [M,1024] K|V buffer.  Work items longest segment first, as the engine orders them.
int b200pf_op_attention_bench(int device, const int32_t* seg_T, int n_seg, int n_heads, int cross, int impl, int iters, float* ms_out,
                              double* flops_out) {
  RC(select_device(device));
  const int D = n_heads * 128;
  std::vector<int> koff(n_seg), klen(n_seg), qoff(n_seg), qlen(n_seg);
  int M = 0, Lq = 0;
  double fl = 0;
  for (int s = 0; s < n_seg; ++s) {
    koff[s] = M; klen[s] = seg_T[s]; M += seg_T[s] + 1;
    qlen[s] = cross ? (seg_T[s] + 1) / 2 : seg_T[s];
    qoff[s] = cross ? Lq : koff[s];
    Lq += qlen[s];
    fl += 4.0 * qlen[s] * klen[s] * D;
  }
  const int q_rows = cross ? Lq : M;
  const int ldq = cross ? D : 3 * D, ldkv = cross ? 2 * D : 3 * D;
  DevBuf dQ, dKV, dO, dqo, dql, dko, dkl, dWork;
  RC(dKV.alloc((size_t)M * ldkv * 2)); RC(dO.alloc((size_t)q_rows * D * 2));
  fill_random_h16_kernel<<<1024, 256>>>(dKV.as<__nv_bfloat16>(), (int64_t)M * ldkv, 11u, g_op_f16);
  if (cross) { RC(dQ.alloc((size_t)q_rows * D * 2)); fill_random_h16_kernel<<<1024, 256>>>(dQ.as<__nv_bfloat16>(), (int64_t)q_rows * D, 12u, g_op_f16); }
  RC(up_raw(qoff.data(), n_seg, &dqo)); RC(up_raw(qlen.data(), n_seg, &dql)); RC(up_raw(koff.data(), n_seg, &dko)); RC(up_raw(klen.data(), n_seg, &dkl));
  std::vector<AttnWork> work;
  for (int s = 0; s < n_seg; ++s)
    for (int q0 = 0; q0 < seg_T[s]; q0 += 128) work.push_back(AttnWork{s, q0});   // the engine builds the list from T for both uses
  std::stable_sort(work.begin(), work.end(), [&](const AttnWork& x, const AttnWork& y) { return seg_T[x.seg] > seg_T[y.seg]; });
  if (cross) std::stable_sort(work.begin(), work.end(), [](const AttnWork& x, const AttnWork& y) { return x.q0 < y.q0; });   // as the engine does
  RC(up_raw(work.data(), work.size(), &dWork));
  AttnProblem p;
  p.f16 = g_op_f16; p.num_sms = sm_count();
  p.q = cross ? dQ.as<__nv_bfloat16>() : dKV.as<__nv_bfloat16>(); p.q_rows = q_rows; p.ldq = ldq; p.q_col0 = 0;
  p.kv = dKV.as<__nv_bfloat16>(); p.kv_rows = M; p.ldkv = ldkv; p.k_col0 = cross ? 0 : D; p.v_col0 = cross ? D : 2 * D;
  p.out = dO.as<__nv_bfloat16>(); p.ldo = D;
  p.q_row_off = dqo.as<int>(); p.q_len = dql.as<int>(); p.kv_row_off = dko.as<int>(); p.kv_len = dkl.as<int>();
  p.work = dWork.as<AttnWork>(); p.n_work = (int)work.size(); p.n_heads = n_heads;
  for (int i = 0; i < 3; ++i) { int rc = attention_tcgen05(p, 0); if (rc) return check_cuda((cudaError_t)rc, "attention launch"); }
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, 0);
  for (int i = 0; i < iters; ++i) { int rc = attention_tcgen05(p, 0); if (rc) return check_cuda((cudaError_t)rc, "attention launch"); }
  cudaEventRecord(b, 0);
  RC(sync_ok("op_attention_bench"));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (ms_out) *ms_out = ms / iters;
  if (flops_out) *flops_out = fl;
  return 0;
}

int b200pf_op_fsmn(int device, const float* x, const float* w, const int32_t* seg_off, int n_seg, float* out) {
  RC(select_device(device));
  const int rows = seg_off[n_seg];
  std::vector<int2> info(rows);
  for (int s = 0; s < n_seg; ++s)
    for (int r = seg_off[s]; r < seg_off[s + 1]; ++r) info[r] = make_int2(r - seg_off[s], seg_off[s + 1] - seg_off[s]);
  std::vector<float> wt((size_t)11 * 512);
  for (int c = 0; c < 512; ++c)
    for (int k = 0; k < 11; ++k) wt[(size_t)k * 512 + c] = w[(size_t)c * 11 + k];
  DevBuf tx, dX, dW, dInfo, dOb, dO;
  RC(up_bf16(x, (size_t)rows * 512, &tx, &dX));
  RC(up_f32(wt.data(), wt.size(), &dW));
  RC(up_raw(info.data(), info.size(), &dInfo));
  RC(dOb.alloc((size_t)rows * 512 * 2)); RC(dO.alloc((size_t)rows * 512 * 4));
  int rc = fsmn_launch(dX.as<__nv_bfloat16>(), 512, 0, dW.as<float>(), dInfo.as<int2>(), rows, nullptr, 0, dOb.as<__nv_bfloat16>(), nullptr, 0, g_op_f16);
  if (rc) return check_cuda((cudaError_t)rc, "fsmn launch");
  widen(dOb.as<__nv_bfloat16>(), dO.as<float>(), (int64_t)rows * 512);
  RC(sync_ok("op_fsmn"));
  return check_cuda(cudaMemcpy(out, dO.p, (size_t)rows * 512 * 4, cudaMemcpyDeviceToHost), "D2H");
}

int b200pf_op_cif(int device, const float* alphas, const float* hidden, const int32_t* seg_off, int n_seg,
                  float threshold, int32_t* n_tok, float* fires, float* embeds, int32_t* fire_frames, int64_t cap_tok) {
  RC(select_device(device));
  const int rows = seg_off[n_seg];
  std::vector<int> row_off(n_seg + 1), seg_T(n_seg);
  for (int s = 0; s <= n_seg; ++s) row_off[s] = seg_off[s];
  for (int s = 0; s < n_seg; ++s) seg_T[s] = seg_off[s + 1] - seg_off[s] - 1;  // rows include the tail frame
  DevBuf dA, dH, dRo, dT, dCur, dRem, dFv, dNt, dFr, dTo, dTot, dEmb, dInfo, dFrame;
  RC(up_f32(alphas, rows, &dA)); RC(up_f32(hidden, (size_t)rows * 512, &dH));
  RC(up_raw(row_off.data(), row_off.size(), &dRo)); RC(up_raw(seg_T.data(), seg_T.size(), &dT));
  RC(dCur.alloc(rows * 4)); RC(dRem.alloc(rows * 4)); RC(dFv.alloc(rows * 4)); RC(dNt.alloc(n_seg * 4)); RC(dFr.alloc(rows * 4));
  RC(dTo.alloc((n_seg + 1) * 4)); RC(dTot.alloc(4)); RC(dEmb.alloc((size_t)rows * 512 * 4)); RC(dInfo.alloc(rows * 8)); RC(dFrame.alloc(rows * 4));
  int rc = cif_fire_launch(dA.as<float>(), dRo.as<int>(), dT.as<int>(), n_seg, threshold, dCur.as<float>(), dRem.as<float>(),
                           dFv.as<float>(), dNt.as<int>(), dFr.as<int>(), 0);
  if (!rc) rc = cif_scan_launch(dNt.as<int>(), n_seg, dTo.as<int>(), dTot.as<int>(), 0);
  if (!rc) rc = cif_embed_launch(dH.as<float>(), dCur.as<float>(), dRem.as<float>(), dFr.as<int>(), dRo.as<int>(), dTo.as<int>(),
                                 n_seg, rows, dEmb.as<float>(), dInfo.as<int2>(), dFrame.as<int>(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "cif launch");
  RC(sync_ok("op_cif"));
  int total = 0;
  RC(check_cuda(cudaMemcpy(&total, dTot.p, 4, cudaMemcpyDeviceToHost), "D2H"));
  if (total > cap_tok) { set_error("op_cif: cap_tok too small"); return B200PF_ERR_CAPACITY; }
  RC(check_cuda(cudaMemcpy(n_tok, dNt.p, (size_t)n_seg * 4, cudaMemcpyDeviceToHost), "D2H"));
  RC(check_cuda(cudaMemcpy(fires, dFv.p, (size_t)rows * 4, cudaMemcpyDeviceToHost), "D2H"));
  if (total) {
    RC(check_cuda(cudaMemcpy(embeds, dEmb.p, (size_t)total * 512 * 4, cudaMemcpyDeviceToHost), "D2H"));
    RC(check_cuda(cudaMemcpy(fire_frames, dFrame.p, (size_t)total * 4, cudaMemcpyDeviceToHost), "D2H"));
  }
  return 0;
}

int b200pf_op_lstm(int device, const float* x, int rows, const int32_t* seq_off, const int32_t* seq_len, int n_seq, int n_dir,
                   const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int bf16_out, float* out) {
  RC(select_device(device));
  if (n_dir < 1 || n_dir > 2 || rows <= 0 || n_seq <= 0) { set_error("op_lstm: bad argument"); return B200PF_ERR_INVALID; }
  const int G = 2048 * n_dir;
  DevBuf tx, dX, tw, dWih, th, dWhh, dB, dGx, dOff, dLen, dOutB, dOut;
  RC(up_bf16(x, (size_t)rows * 512, &tx, &dX));
  RC(up_bf16(w_ih, (size_t)G * 512, &tw, &dWih));
  RC(up_bf16(w_hh, (size_t)G * 512, &th, &dWhh));
  std::vector<float> bsum(G);
  for (int i = 0; i < G; ++i) bsum[i] = b_ih[i] + b_hh[i];
  RC(up_f32(bsum.data(), G, &dB));
  RC(dGx.alloc((size_t)rows * G * 2));
  RC(up_raw(seq_off, n_seq, &dOff)); RC(up_raw(seq_len, n_seq, &dLen));
  RC(dOut.alloc((size_t)rows * 512 * n_dir * 4)); RC(dOutB.alloc((size_t)rows * 512 * n_dir * 2));
  RC(check_cuda(cudaMemset(dOut.p, 0, (size_t)rows * 512 * n_dir * 4), "memset"));
  RC(check_cuda(cudaMemset(dOutB.p, 0, (size_t)rows * 512 * n_dir * 2), "memset"));
  GemmProblem gp;
  gp.f16 = g_op_f16;
  gp.A = dX.as<__nv_bfloat16>(); gp.lda = 512; gp.rows_a = rows; gp.W = dWih.as<__nv_bfloat16>(); gp.ldw = 512; gp.M = rows; gp.N = G; gp.K = 512;
  GemmEpilogue ge;
  ge.bias = dB.as<float>(); ge.out_bf16 = dGx.as<__nv_bfloat16>(); ge.ld_out_bf16 = G;
  int rc = gemm_bf16_tcgen05(gp, ge, sm_count(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "lstm input projection");
  LstmParams lp;
  lp.f16 = g_op_f16;
  lp.gx = dGx.as<__nv_bfloat16>(); lp.ld_gx = G; lp.whh = dWhh.as<__nv_bfloat16>(); lp.seq_off = dOff.as<int>(); lp.seq_len = dLen.as<int>();
  lp.n_seq = n_seq; lp.n_dir = n_dir; lp.reverse_mask = n_dir == 2 ? 2 : 0;
  lp.out_bf16 = dOutB.as<__nv_bfloat16>(); lp.ld_out = 512 * n_dir;
  if (!bf16_out) { lp.out_f32 = dOut.as<float>(); lp.ld_out_f32 = 512 * n_dir; }
  rc = lstm_launch(lp, 0);
  if (rc) return check_cuda((cudaError_t)rc, "lstm launch");
  if (bf16_out) widen(dOutB.as<__nv_bfloat16>(), dOut.as<float>(), (int64_t)rows * 512 * n_dir);
  RC(sync_ok("op_lstm"));
  return check_cuda(cudaMemcpy(out, dOut.p, (size_t)rows * 512 * n_dir * 4, cudaMemcpyDeviceToHost), "D2H");
}

int b200pf_op_lstm_bench(int device, int n_seq, int len, int n_dir, int iters, float* ms_out, int* max_clusters) {
  RC(select_device(device));
  if (n_seq <= 0 || len <= 0 || n_dir < 1 || n_dir > 2) { set_error("op_lstm_bench: bad argument"); return B200PF_ERR_INVALID; }
  const size_t rows = (size_t)n_seq * len;
  const int G = 2048 * n_dir;
  DevBuf dGx, dWhh, dOff, dLen, dOut;
  RC(dGx.alloc(rows * G * 2)); RC(dWhh.alloc((size_t)G * 512 * 2)); RC(dOut.alloc(rows * 512 * n_dir * 2));
  fill_random_h16_kernel<<<1024, 256>>>(dGx.as<__nv_bfloat16>(), (int64_t)rows * G, 5u, g_op_f16);
  fill_random_h16_kernel<<<1024, 256>>>(dWhh.as<__nv_bfloat16>(), (int64_t)G * 512, 6u, g_op_f16);  // |w| < 1: saturating, still finite
  std::vector<int> off(n_seq), ln(n_seq, len);
  for (int i = 0; i < n_seq; ++i) off[i] = i * len;
  RC(up_raw(off.data(), off.size(), &dOff)); RC(up_raw(ln.data(), ln.size(), &dLen));
  LstmParams lp;
  lp.f16 = g_op_f16;
  lp.gx = dGx.as<__nv_bfloat16>(); lp.ld_gx = G; lp.whh = dWhh.as<__nv_bfloat16>(); lp.seq_off = dOff.as<int>(); lp.seq_len = dLen.as<int>();
  lp.n_seq = n_seq; lp.n_dir = n_dir; lp.reverse_mask = n_dir == 2 ? 2 : 0; lp.out_bf16 = dOut.as<__nv_bfloat16>(); lp.ld_out = 512 * n_dir;
  int rc = lstm_launch(lp, 0);
  if (rc) return check_cuda((cudaError_t)rc, "lstm launch");
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, 0);
  for (int i = 0; i < iters; ++i) { rc = lstm_launch(lp, 0); if (rc) return check_cuda((cudaError_t)rc, "lstm launch"); }
  cudaEventRecord(b, 0);
  RC(sync_ok("op_lstm_bench"));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (ms_out) *ms_out = ms / iters;
  if (max_clusters) *max_clusters = lstm_max_active_clusters();
  return 0;
}

int b200pf_op_us_peaks(int device, const float* alpha2, const int32_t* seq_off, const int32_t* seq_len, const int32_t* n_tok,
                       int n_seg, int rows, float threshold, float* us_alphas, float* us_peaks) {
  RC(select_device(device));
  DevBuf dA, dOff, dLen, dTok, dUa, dUp;
  RC(up_f32(alpha2, rows, &dA)); RC(up_raw(seq_off, n_seg, &dOff)); RC(up_raw(seq_len, n_seg, &dLen)); RC(up_raw(n_tok, n_seg, &dTok));
  RC(dUa.alloc((size_t)rows * 4)); RC(dUp.alloc((size_t)rows * 4));
  RC(check_cuda(cudaMemset(dUa.p, 0, (size_t)rows * 4), "memset")); RC(check_cuda(cudaMemset(dUp.p, 0, (size_t)rows * 4), "memset"));
  int rc = us_peak_launch(dA.as<float>(), dOff.as<int>(), dLen.as<int>(), dTok.as<int>(), n_seg, threshold, dUa.as<float>(), dUp.as<float>(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "us_peak launch");
  RC(sync_ok("op_us_peaks"));
  RC(check_cuda(cudaMemcpy(us_alphas, dUa.p, (size_t)rows * 4, cudaMemcpyDeviceToHost), "D2H"));
  return check_cuda(cudaMemcpy(us_peaks, dUp.p, (size_t)rows * 4, cudaMemcpyDeviceToHost), "D2H");
}

int b200pf_op_logprob_topk(int device, const float* logits, int rows, int V, int k, float* lse, float* lp, int32_t* ids) {
  RC(select_device(device));
  if (rows <= 0 || k <= 0 || k > B200PF_MAX_TOPK || (V & 3)) { set_error("op_logprob_topk: bad argument"); return B200PF_ERR_INVALID; }
  DevBuf dL, dLse, dLp, dId;
  RC(up_f32(logits, (size_t)rows * V, &dL));
  RC(dLse.alloc((size_t)rows * 4)); RC(dLp.alloc((size_t)rows * k * 4)); RC(dId.alloc((size_t)rows * k * 4));
  int rc = logprob_topk_launch(dL.as<float>(), V, nullptr, rows, k, dLse.as<float>(), dLp.as<float>(), dId.as<int>(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "logprob_topk launch");
  RC(sync_ok("op_logprob_topk"));
  RC(check_cuda(cudaMemcpy(lse, dLse.p, (size_t)rows * 4, cudaMemcpyDeviceToHost), "D2H"));
  RC(check_cuda(cudaMemcpy(lp, dLp.p, (size_t)rows * k * 4, cudaMemcpyDeviceToHost), "D2H"));
  return check_cuda(cudaMemcpy(ids, dId.p, (size_t)rows * k * 4, cudaMemcpyDeviceToHost), "D2H");
}

int b200pf_op_frontend(b200pf_engine* e, const int16_t* pcm, int64_t n, float* fb_out, float* feats_out) {
  if (!e || !pcm) { set_error("null argument"); return B200PF_ERR_INVALID; }
  RC(check_cuda(cudaSetDevice(e->device), "cudaSetDevice"));
  const int nfb = num_fbank_frames(n), T = num_lfr_frames(n);
  if (nfb <= 0) return 0;
  std::vector<int2> info(T + 1);
  std::vector<int> rseg(T + 1, 0);
  for (int t = 0; t < T; ++t) info[t] = make_int2(t, T);
  info[T] = make_int2(-1, T); rseg[T] = -1;
  const int64_t soff[2] = {0, n};
  const int fboff[2] = {0, nfb};
  DevBuf dP, dS, dF, dInfo, dSeg, dFb, dX0, dFeat;
  RC(up_raw(pcm, (size_t)n, &dP)); RC(up_raw(soff, 2, &dS)); RC(up_raw(fboff, 2, &dF));
  RC(up_raw(info.data(), info.size(), &dInfo)); RC(up_raw(rseg.data(), rseg.size(), &dSeg));
  RC(dFb.alloc((size_t)nfb * 80 * 4)); RC(dX0.alloc((size_t)(T + 1) * 560 * 4)); RC(dFeat.alloc((size_t)(T + 1) * 560 * 4));
  int rc = fbank_launch(dP.p, 0, dS.as<int64_t>(), dF.as<int>(), 1, nfb, e->ft, dFb.as<float>(), 0);
  if (!rc) rc = lfr_cmvn_posenc_launch(dFb.as<float>(), dF.as<int>(), dSeg.as<int>(), dInfo.as<int2>(), T + 1, e->ft,
                                       sqrtf((float)e->cfg.d_model), dX0.as<float>(), dFeat.as<float>(), 0);
  if (rc) return check_cuda((cudaError_t)rc, "frontend launch");
  RC(sync_ok("op_frontend"));
  if (fb_out) RC(check_cuda(cudaMemcpy(fb_out, dFb.p, (size_t)nfb * 80 * 4, cudaMemcpyDeviceToHost), "D2H"));
  if (feats_out) RC(check_cuda(cudaMemcpy(feats_out, dFeat.p, (size_t)T * 560 * 4, cudaMemcpyDeviceToHost), "D2H"));
  return 0;
}

}  // extern "C"
