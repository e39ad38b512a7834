// Bandwidth-bound kernels of the Paraformer path (front end, LayerNorm, FSMN memory block, CIF).
// Reference anchors are cited per function; paths are relative to /root/reference/onnxruntime.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace pf {

// 16-bit operand storage: every `__nv_bfloat16*` below is 2-byte storage whose FORMAT is chosen by the trailing `f16` argument
// (0 = bf16, 1 = IEEE fp16: the engine's precision 1, see ptx.cuh pack_h2); "bf16" in the parameter names is historical.
//
// Packed row layout shared by all kernels: segment i owns rows [row_off[i], row_off[i] + T_i) followed
// by ONE gap row (zero features; it doubles as the CIF tail frame, tail_threshold 0.45).
// row_info[r] = {t, T}: frame index inside its segment and the segment's frame count; t = -1 on gap rows.

struct FrontendTables {           // device pointers, built once per engine
  const float* window;            // [400]  hamming, feature-window.cc:25-55
  const double2* twiddle;         // [256]  exp(-2 pi i k / 512)
  const int2* mel_range;          // [80]   {first_bin, size}, mel-computations.cc:172-195
  const float* mel_w;             // packed weights, bin b at mel_w_off[b]
  const int* mel_w_off;           // [80]
  const float* cmvn_mean;         // [560]  am.mvn <AddShift>   paraformer.cpp:325-360
  const float* cmvn_var;          // [560]  am.mvn <Rescale>
  const float* pos_enc;           // [pe_rows][560]  sinusoid, positions from 1 (paraformer-online.cpp:240-268)
  int pe_rows;
};

// Host-side tables of the fbank front end (window, FFT twiddles, packed mel filters); false on overflow of the packed table.
bool fbank_tables_host(std::vector<float>* window, std::vector<double>* twiddle, std::vector<int>* mel_range, std::vector<float>* mel_w,
                       std::vector<int>* mel_w_off);

// K1a  Kaldi fbank: Paraformer::FbankKaldi (src/paraformer.cpp:309-323) -> knf::OnlineFbank.
//      pcm is int16 (is_f32 = 0) or float in [-1,1) (is_f32 = 1; multiplied by 32768 as the reference does).
//      sample_off[n_seg+1] (int64), fb_off[n_seg+1] (frame offsets into fb), fb [n_frames_total][80] fp32.
int fbank_launch(const void* pcm, int is_f32, const int64_t* sample_off, const int* fb_off, int n_seg,
                 int n_frames_total, const FrontendTables& t, float* fb, cudaStream_t s);

// K1b  LFR 7/6 stacking + CMVN (Paraformer::LfrCmvn, src/paraformer.cpp:421-461) fused with the encoder's
//      input scale and position encoding (x * sqrt(512) + PE).  feats_tap (optional) gets the pre-scale
//      LFR+CMVN features [M][560].
int lfr_cmvn_posenc_launch(const float* fb, const int* fb_off, const int* row_seg, const int2* row_info, int M,
                           const FrontendTables& t, float scale, float* x0, float* feats_tap, cudaStream_t s);

// K3  LayerNorm over the last dim (eps 1e-12), fp32 statistics.  in: fp32 or bf16 [rows][D]; out bf16 and/or
//     fp32.  zero_gap: rows with row_info.t < 0 are written as zeros.  rows_dev (optional) overrides rows.
int layernorm_launch(const void* in, int in_is_bf16, int rows, const int* rows_dev, int D, const float* gamma,
                     const float* beta, float eps, __nv_bfloat16* out_bf16, float* out_f32, const int2* row_info,
                     int zero_gap, cudaStream_t s, int f16 = 0);

// K6  FSMN memory block: depthwise conv1d k=11 (zero padded at SEGMENT edges) + identity.
//     mode 0 (encoder): out_bf16[r][c] = conv(v)[r][c] + v[r][c]
//     mode 1 (decoder): y_f32[r][c]   += conv(x)[r][c] + x[r][c]
//     in: bf16 [rows][ld_in] starting at column col0; w_t: [11][512] fp32 (tap-major transpose of fsmn_block.weight).
int fsmn_launch(const __nv_bfloat16* in, int ld_in, int col0, const float* w_t, const int2* row_info, int rows,
                const int* rows_dev, int mode, __nv_bfloat16* out_bf16, float* y_f32, cudaStream_t s, int f16 = 0, int num_sms = 0);

// K7  alpha = sigmoid(h . w + b) on frame rows, tail_threshold on gap rows (CifPredictorV2).
int cif_alpha_launch(const float* h, int M, const float* w, const float* b, const int2* row_info, float tail,
                     float* alpha, cudaStream_t s);

// K8  integrate-and-fire (same recurrence as ParaformerOnline::CifSearch, paraformer-online.cpp:270-345):
//     per segment, sequential fp32 integrate exactly as the reference; a warp prefetches alphas and ranks
//     the fires with ballot/popc prefix counts.
//     Outputs per row: cur (weight into the open token), rem (weight carried into the next token when the
//     row fires), fire_val (integrate after adding alpha), tok_of_fire (local token index or -1).
//     Per segment: n_tok.  Per (segment, local token j): fire_row[row_off[seg] + j] = absolute row of the fire.
int cif_fire_launch(const float* alpha, const int* row_off, const int* seg_T, int n_seg, float threshold,
                    float* cur, float* rem, float* fire_val, int* n_tok, int* fire_row, cudaStream_t s);

//     exclusive scan of n_tok -> tok_off[n_seg+1]; total -> *n_tok_total (device)
int cif_scan_launch(const int* n_tok, int n_seg, int* tok_off, int* n_tok_total, cudaStream_t s);

//     token embeddings E[g][:] = rem[f_prev] h[f_prev] + sum_{t in (f_prev, f]} cur[t] h[t]  (fp32, mul then add
//     as the reference does), plus tok_info[g] = {j, L_seg}, tok_frame[g] = frame index of the fire.
int cif_embed_launch(const float* enc_f32, const float* cur, const float* rem, const int* fire_row, const int* row_off,
                     const int* tok_off, int n_seg, int tok_cap, float* emb, int2* tok_info, int* tok_frame,
                     cudaStream_t s);

// K11 decode packed argmax keys -> token ids
int argmax_decode_launch(const unsigned long long* packed, const int* n_dev, int cap, int* ids, cudaStream_t s);

// K12 LSTM recurrence, hidden 512, gate order i,f,g,o (torch.nn.LSTM).  Config 3 only: the timestamp predictor's
//     BiLSTM (outputs read at src/paraformer.cpp:549-563) and the hotword compiler's LSTM (src/paraformer.cpp:592-693).
//     gx = x W_ih^T + b_ih + b_hh comes from the GEMM; sequences are row ranges [seq_off[s], seq_off[s] + seq_len[s]).
struct LstmParams {
  const __nv_bfloat16* gx = nullptr;   // [rows, ld_gx] bf16; direction d at columns [2048 d, 2048 d + 2048)
  int ld_gx = 0;
  const __nv_bfloat16* whh = nullptr;  // [n_dir][2048][512] bf16
  const int* seq_off = nullptr;        // [n_seq] (device)
  const int* seq_len = nullptr;        // [n_seq] (device)
  int n_seq = 0;
  int n_dir = 1;
  int reverse_mask = 0;                // bit d: direction d runs from the last row to the first
  __nv_bfloat16* out_bf16 = nullptr;   // [rows, ld_out]; direction d at columns [512 d, 512 d + 512)   (optional)
  int ld_out = 0;
  float* out_f32 = nullptr;            // same layout, fp32                                             (optional)
  int ld_out_f32 = 0;
  int dbg = 0;                         // micro-benchmark ablations only (B200PF_LSTM_DBG); 0 in the product
  int f16 = 0;                         // gx, whh and out_bf16 are IEEE fp16 instead of bf16
};
int lstm_launch(const LstmParams& p, cudaStream_t s);
int lstm_max_active_clusters();  // co-resident 16-CTA clusters on the current device (diagnostics)

// K13 timestamp head tail (CifPredictorV3.get_upsample_timestmap): alpha2 = relu(sigmoid(h . w + b) * smooth - noise)
//     over the BiLSTM output h [rows, 1024] bf16 ...
int us_alpha_launch(const __nv_bfloat16* h, int rows, const float* w, const float* b, float smooth, float noise, float* alpha,
                    cudaStream_t s, int f16 = 0);
//     ... then per segment: us_alphas = alpha2 * token_num / sum(alpha2); us_cif_peak = cif_wo_hidden(us_alphas, threshold)
//     (the arrays TimestampOnnx consumes, src/util.cpp:838-870).
int us_peak_launch(const float* alpha2, const int* seq_off, const int* seq_len, const int* n_tok, int n_seg, float threshold,
                   float* us_alphas, float* us_peaks, cudaStream_t s);
// hotword Embedding lookup: out[j] = table[ids[j]] (bf16 rows of 512)
int embed_gather_launch(const __nv_bfloat16* table, int vocab, const int* ids, int n, __nv_bfloat16* out, cudaStream_t s);

// K10 pruned posteriors for the host-side log-prob consumers (WfstDecoder::Search src/wfst-decoder.cpp:27-57,
//     CtcPrefixDecoder): per decoder row, logsumexp and the k largest log-softmax values with their token ids, ordered
//     by (value descending, index ascending).  logits [cap, V] fp32; outputs lse [cap], lp / ids [cap, k].
int logprob_topk_launch(const float* logits, int V, const int* n_dev, int cap, int k, float* lse, float* lp, int* ids, cudaStream_t s);

// fp32 <-> 16-bit storage conversion (weight upload, test taps)
int f32_to_bf16_launch(const float* in, __nv_bfloat16* out, int64_t n, cudaStream_t s, int f16 = 0);
int h16_to_f32_launch(const __nv_bfloat16* in, float* out, int64_t n, cudaStream_t s, int f16 = 0);

}  // namespace pf
