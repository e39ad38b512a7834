// Engine (one GPU: weights + workspace) and Batch (one batch of segments) behind include/b200pf.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <set>
#include <string>
#include <vector>

#include "../../include/b200pf.h"
#include "attention.cuh"
#include "gemm.cuh"
#include "kernels.cuh"

namespace pf {

void set_error(const std::string& msg);
int check_cuda(cudaError_t e, const char* what);  // 0 or B200PF_ERR_CUDA (and sets the error text)
int num_fbank_frames(int64_t n_samples);  // feature-window.cc:73-87 (snip_edges)
int num_lfr_frames(int64_t n_samples);    // paraformer.cpp:424
uint16_t f32_to_h16(float f, int f16);    // host-side round-to-nearest-even to bf16 (0) or saturating IEEE fp16 (1)
float h16_to_f32(uint16_t h, int f16);    // exact widening of the same formats

struct Linear {
  __nv_bfloat16* w = nullptr;  // [out, in] bf16
  float* b = nullptr;          // [out] or null
  int out = 0, in = 0;
};
struct Norm {
  float* g = nullptr;
  float* b = nullptr;
};
struct EncLayer {
  Norm ln1, ln2;
  Linear qkv, out, w1, w2;
  float* fsmn_wt = nullptr;  // [11][512]
  int din = 512;
};
struct DecLayer {
  Norm ln1, lnff, ln2, ln3;
  Linear w1, w2, q, kv, out;
  // feed_forward.norm folded into w_2 (engine option ffn_ln_fold, gemm.cuh): w2f = w_2 . diag(gamma) in the operand format,
  // w2f.b[n] = sum_k beta[k] w_2[n][k], w2_csum[n] = sum_k w2f[n][k] of the ROUNDED folded weights
  Linear w2f;
  float* w2_csum = nullptr;
  float* fsmn_wt = nullptr;
  bool has_attn = true;
};

struct DeviceArena {
  uint8_t* base = nullptr;
  size_t size = 0, used = 0;
  void* take(size_t bytes) {
    size_t off = (used + 255) & ~size_t(255);
    if (off + bytes > size) return nullptr;
    used = off + bytes;
    return base + off;
  }
};

}  // namespace pf

struct b200pf_engine {
  b200pf_config cfg{};
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;          // FSMN memory block runs here, concurrently with the attention kernel
  cudaStream_t copy = nullptr;          // host->device staging of the NEXT batch (b200pf_engine_copy_stream)
  cudaStream_t d2h = nullptr;           // device->host result reads of a FINISHED batch (b200pf_batch_collect), while the next one computes
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::mutex batches_mu;
  std::set<b200pf_batch*> live_batches; // orphaned (resources freed, b->e = nullptr) by b200pf_engine_destroy
  int overlap = 2;
  int ffn_ln_fold = 1;                  // decoder feed-forward LayerNorm(2048) applied inside the w_2 GEMM epilogue (no pass over the hidden)
  int f16 = 1;                          // 16-bit operand format of weights and activations: 0 bf16, 1 IEEE fp16 (cfg.precision)
  std::string lang = "zh-cn";
  std::vector<std::string> tokens;
  std::mutex mu;  // one forward at a time per engine (the workspace is shared)

  pf::DeviceArena warena;  // weights + tables
  std::vector<pf::EncLayer> enc;
  pf::Norm enc_after;
  pf::Linear pred_conv;    // [512, 3*512] tap-major
  float* pred_out_w = nullptr;
  float* pred_out_b = nullptr;
  std::vector<pf::DecLayer> dec;
  pf::DecLayer dec3;
  pf::Norm dec_after;
  pf::Linear vocab;
  pf::FrontendTables ft{};
  // config 3.  timestamp: CifPredictorV3 upsample head; contextual: bias decoder + hotword LSTM.
  int us_times = 3;
  float smooth2 = 0.25f, noise2 = 0.01f;
  pf::Linear us_cnn;                    // ConvTranspose1d(512,512,k=3,s=3) as [3*512, 512]: row j*512+o = W[i][o][j]
  pf::Linear blstm_ih;                  // [2*2048, 512] forward then reverse; bias = b_ih + b_hh
  __nv_bfloat16* blstm_hh = nullptr;    // [2][2048][512]
  float* us_out_w = nullptr;            // cif_output2 [1024]
  float* us_out_b = nullptr;
  pf::Norm bias_ln3;                    // decoder.bias_decoder.norm3
  pf::Linear bias_q, bias_kv, bias_out; // decoder.bias_decoder.src_attn.*
  pf::Linear bias_output;               // decoder.bias_output [512, 1024] (1x1 conv, no bias)
  __nv_bfloat16* hw_table = nullptr;    // bias_embed.weight [vocab, 512]
  pf::Linear hw_ih;                     // bias_encoder W_ih [2048, 512], bias = b_ih + b_hh
  __nv_bfloat16* hw_hh = nullptr;       // [2048][512]

  // workspace (sized by cfg.max_rows = R)
  pf::DeviceArena ws;
  float* fb = nullptr;              // [6R, 80]
  float* x0 = nullptr;              // [R, 560] fp32   (decoder: t [R,512])
  float* x = nullptr;               // [R, 512] fp32   (after the encoder: predictor hidden)
  __nv_bfloat16* hb = nullptr;      // [R, 560] bf16   LayerNorm output feeding the next GEMM
  __nv_bfloat16* qkv = nullptr;     // [R, 1536] bf16  (decoder: cross k/v [R,1024])
  __nv_bfloat16* mem = nullptr;     // [R, 512] bf16   FSMN memory (decoder: cross q)
  __nv_bfloat16* att = nullptr;     // [R, 512] bf16
  __nv_bfloat16* ffn = nullptr;     // [R, 2048] bf16
  float2* ffn_stats = nullptr;      // [R, 16] (sum, sum of squares) per 128-column span of a decoder feed-forward hidden row
  float* enc_f32 = nullptr;         // [R, 512]
  __nv_bfloat16* enc_bf16 = nullptr;
  float* y = nullptr;               // [R, 512] decoder residual stream
  float *alpha = nullptr, *cif_cur = nullptr, *cif_rem = nullptr, *fire_val = nullptr;
  int* fire_row = nullptr;
  unsigned long long* amax = nullptr;
  int2* tok_info = nullptr;
  // config-3 workspace
  __nv_bfloat16* us_gx = nullptr;       // [3R, 4096] bf16  BiLSTM input projection
  __nv_bfloat16* us_h = nullptr;        // [3R, 1024] bf16  BiLSTM output
  float* us_a2 = nullptr;               // [3R] alpha2 before the per-segment rescale (results: b200pf_batch::d_us_*)
  __nv_bfloat16* hw_kv = nullptr;       // [max_hotwords, 1024] bf16  bias_decoder k/v of the hotword embeddings
  // CUDA graphs of the forward for SMALL batches (options "graphs", "graph_max_rows"): a batch-1 forward is ~600 dependent
  // launches, i.e. bound by launch latency; its layout is padded to a bucket (rows to a multiple of 32, with gap rows) so that
  // batches of similar size share one captured graph, which replays with a single launch call.
  int use_graphs = 1;
  int graph_max_rows = 4096;
  typedef std::tuple<const void*, int, int, int, int, int> GraphKey;   // batch, rows_run, segments, work items, hotwords, flags
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; uint64_t last_use = 0; int seen = 0; };
  std::map<GraphKey, GraphEntry> graphs;
  uint64_t graph_clock = 0;
  long long graph_replays = 0, graph_captures = 0;
  // per-category CUDA-event timing of the launches (option "profile")
  int profile = 0;
  struct ProfRec { cudaEvent_t a, b; int cat; double work; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[16] = {0}, prof_work[16] = {0};
  long long prof_launches[16] = {0};
  // taps
  int taps = 0;
  float* tap_feats = nullptr;   // [R, 560]
  float* tap_emb = nullptr;     // [R, 512]
  float* tap_logits = nullptr;  // [R, vocab]
  // pruned posteriors (option "logprob_topk" = k)
  int topk = 0;
  float* full_logits = nullptr;   // [R, vocab] fp32, only allocated when topk > 0 (scratch of one run; results: b200pf_batch::d_topk_*)
};

struct b200pf_batch {
  b200pf_engine* e = nullptr;
  int64_t max_samples = 0;
  void* d_pcm = nullptr;  // int16 or float
  int pcm_is_f32 = 0;
  // host-pinned meta mirrored on the device
  uint8_t* h_meta = nullptr;
  uint8_t* d_meta = nullptr;
  size_t meta_bytes = 0;
  // views into meta (host / device)
  int64_t* h_sample_off = nullptr; const int64_t* d_sample_off = nullptr;
  int* h_fb_off = nullptr;         const int* d_fb_off = nullptr;
  int* h_row_off = nullptr;        const int* d_row_off = nullptr;
  int* h_seg_T = nullptr;          const int* d_seg_T = nullptr;
  int* h_us_off = nullptr;         const int* d_us_off = nullptr;   // [S] 3 * row_off (upsampled rows)
  int* h_us_len = nullptr;         const int* d_us_len = nullptr;   // [S] 3 * T
  int* h_zero = nullptr;           const int* d_zero = nullptr;     // [S] zeros: every segment reads the same hotword rows
  int* h_hw_len = nullptr;         const int* d_hw_len = nullptr;   // [S] n_hw
  int* h_row_seg = nullptr;        const int* d_row_seg = nullptr;
  int2* h_row_info = nullptr;      const int2* d_row_info = nullptr;
  pf::AttnWork* h_work = nullptr;  const pf::AttnWork* d_work = nullptr;      // self-attention tiles, longest segment first
  pf::AttnWork* h_work_x = nullptr; const pf::AttnWork* d_work_x = nullptr;   // cross-attention: the same tiles by (q0, longest first)
  // device results
  int* d_n_tok = nullptr;     // [S]
  int* d_tok_off = nullptr;   // [S+1]
  int* d_tok_total = nullptr; // [1]
  int* d_ids = nullptr;       // [R]
  int* d_tok_frame = nullptr; // [R]
  uint8_t* h_res = nullptr;   // pinned: n_tok[S] tok_off[S+1] ids[R] tok_frame[R]
  float* h_us = nullptr;      // pinned: us_alphas[3R] us_peaks[3R] (timestamp models)
  uint8_t* h_topk = nullptr;  // pinned: lse[R] lp[R*32] ids[R*32], allocated on first use
  int16_t* h_stage = nullptr; // pinned int16 staging of float input (b200pf_batch_stage_f32 fast path), allocated on first use
  float* d_us_alphas = nullptr;   // [3R] timestamp models: us_alphas / us_cif_peak of THIS batch (one allocation)
  float* d_us_peaks = nullptr;    // [3R]
  float* d_topk_lse = nullptr;    // [R]      pruned posteriors of THIS batch (one allocation, made on the first run that asks)
  float* d_topk_lp = nullptr;     // [R, 32]
  int* d_topk_id = nullptr;       // [R, 32]
  __nv_bfloat16* d_hw = nullptr;  // [max_hotwords, 512] 16-bit hotword embeddings of this batch (contextual models)
  int n_hw = 0;
  int topk_run = 0;   // k the last run produced pruned posteriors with
  // staged state
  int n_seg_in = 0;             // segments the caller passed
  int n_seg = 0;                // segments on the device (T > 0)
  std::vector<int> dev_of_in;   // caller index -> device segment or -1
  std::vector<int> T_in;        // caller index -> T
  int rows = 0, n_frames = 0, n_work = 0;
  int rows_run = 0, n_frames_run = 0, n_work_run = 0;   // what the kernels are launched over: == the real counts, or the padded
                                                        // bucket of a graph-able small batch (extra rows are gap rows)
  bool graph_ok = false;
  int64_t launches = 0;
  double flops = 0.0;
  int64_t last_tokens = -1;     // tokens of the last collected run and the row count it belonged to (profiling: exact decoder FLOPs)
  int last_tokens_rows = -1;
  cudaEvent_t staged = nullptr;
  cudaEvent_t done = nullptr;   // recorded after the last kernel of b200pf_batch_run: collect waits for THIS batch only
  bool collected = false;
};
