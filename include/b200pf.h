/*
 * b200pf — C ABI of the B200-native offline Paraformer acoustic-model path.
 *
 * This is the drop-in boundary for ONE hot path of duj12/ASR-2Pass: `funasr::Model::Forward` for the
 * offline Paraformer (reference: onnxruntime/src/paraformer.cpp:463-589, batched form
 * onnxruntime/src/paraformer-torch.cpp:301-475).  The reference's own exported API is C++
 * (onnxruntime/include/funasrruntime.h:60-138 carries std::map / std::string / std::vector), so the C++
 * shim in asr-2pass_b200/csrc/host/ re-exports those symbols and calls THIS header underneath; see
 * INTEGRATION.md for the binding a maintainer adds on the reference side.
 *
 * Plain pointers and sizes only; no torch or STL types cross this boundary.  Every function returns 0 on
 * success and a non-zero code on failure; b200pf_last_error() describes the last failure on this thread.
 * There is no CPU fallback: without a CUDA device of compute capability 10.x every entry point that
 * computes fails with B200PF_ERR_NO_DEVICE.
 */
#ifndef B200PF_H_
#define B200PF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PF_OK 0
#define B200PF_ERR_INVALID 1
#define B200PF_ERR_NO_DEVICE 2
#define B200PF_ERR_CUDA 3
#define B200PF_ERR_IO 4
#define B200PF_ERR_CAPACITY 5

typedef struct b200pf_engine b200pf_engine; /* one GPU: resident weights + activation workspace */
typedef struct b200pf_batch b200pf_batch;   /* one batch of segments: device PCM, layout, results */

typedef struct b200pf_config {
  int32_t feat_dim, d_model, n_heads, d_ff, n_enc, n_dec, kernel, vocab, pred_residual;
  float cif_threshold, tail_threshold, ln_eps;
  int32_t sample_rate;     /* frontend_conf.fs of config.yaml (paraformer.cpp:188-190) */
  int32_t max_rows;        /* packed LFR rows (frames + one gap row per segment) one batch may hold */
  int32_t max_segments;
  int32_t timestamp;       /* 1: the 4-output model (us_alphas, us_cif_peak; paraformer.cpp:549-563) */
  int32_t contextual;      /* 1: the hotword model (bias_embed input, paraformer.cpp:515-531; model_eb, :592-693) */
  int32_t precision;       /* B200PF_PREC_*: 16-bit operand format of the engine's weights and activations */
} b200pf_config;

/* Tensor-core operand format (tcgen05 kind::f16 takes either at the same rate; accumulation, residual streams, LayerNorm
 * statistics, softmax and CIF are fp32 in both).  FP16 is the default: IEEE half has three more mantissa bits than bf16 and
 * keeps encoder output and logits within the 1e-2 relative tolerance against the fp32 reference graph (bf16 operands give
 * 2-4 % on the logits); conversions saturate at +-65504.  BF16 is north_star's literal format. */
#define B200PF_PREC_BF16 0
#define B200PF_PREC_FP16 1

#define B200PF_MAX_HOTWORDS 4096  /* rows of the hotword embedding a batch may carry (incl. the blank row) */
#define B200PF_HOTWORD_LEN 10     /* max_hotword_len, paraformer.cpp:600 */
#define B200PF_MAX_TOPK 32        /* largest k of the pruned-posterior output */

/* Per-batch results, written into caller-owned host buffers by b200pf_batch_collect.
 * Segment i produced token_counts[i] tokens; its ids are token_ids[token_offsets[i] .. +count).
 * fire_frames holds, per token, the LFR frame index (60 ms units) at which CIF fired; index T_i denotes
 * the tail frame.  lfr_frames[i] = T_i (0 for a segment shorter than one fbank window: the reference
 * returns "" for it, paraformer.cpp:477-480). */
typedef struct b200pf_result {
  int32_t* token_counts;   /* [n_seg]      */
  int32_t* token_offsets;  /* [n_seg + 1]  */
  int32_t* lfr_frames;     /* [n_seg]      */
  int32_t* token_ids;      /* [cap_tokens] */
  int32_t* fire_frames;    /* [cap_tokens] may be NULL */
  int64_t cap_tokens;
  int64_t n_tokens;        /* out: total tokens written */
  /* timestamp models only (all may be NULL): the two extra graph outputs, 3 * T_i values per segment starting at
   * us_offsets[i] (paraformer.cpp:549-563 reads them per item; paraformer-torch.cpp:431-446 the batched form). */
  float* us_alphas;        /* [cap_us] */
  float* us_peaks;         /* [cap_us] */
  int32_t* us_offsets;     /* [n_seg + 1] */
  int64_t cap_us;
  /* pruned posteriors (engine option "logprob_topk" = k > 0; all may be NULL): per token, logsumexp of its logits row and
   * the k largest log-softmax values with their ids, ordered by (value descending, id ascending) - what the reference's
   * log-prob consumers (WfstDecoder::Search, wfst-decoder.cpp:27-57; CtcPrefixDecoder) need instead of [L,8404] rows. */
  float* token_lse;        /* [cap_tokens] */
  float* topk_logprob;     /* [cap_tokens * k] */
  int32_t* topk_ids;       /* [cap_tokens * k] */
  int32_t topk_k;          /* out: k of this result (0 when the option is off) */
} b200pf_result;

const char* b200pf_last_error(void);
int b200pf_version(void);
/* Number of CUDA devices with compute capability 10.x; 0 when there is none (never falls back). */
int b200pf_device_count(void);

/* Host-only: parse <model_dir>/{am.mvn, config.yaml, tokens.json, model.b200pf} exactly as
 * b200pf_engine_create does and report the architecture (no GPU needed; max_rows/max_segments = 0). */
int b200pf_model_dir_probe(const char* model_dir, b200pf_config* out, int* n_tokens, int* n_tensors);

/* Replaces Paraformer::InitAsr (paraformer.cpp:21-53): reads <model_dir>/{am.mvn, config.yaml,
 * tokens.json, model.b200pf}, uploads 16-bit weights (fp16 by default, see b200pf_engine_create_prec) to `device`, allocates workspace for batches of up to
 * max_rows packed rows / max_segments segments (0 -> defaults 32768 / 4096). */
int b200pf_engine_create(const char* model_dir, int device, int max_rows, int max_segments, b200pf_engine** out);
/* Same with the operand format chosen explicitly (B200PF_PREC_BF16 / B200PF_PREC_FP16; -1 = the default, which the environment
 * variable B200PF_PREC=bf16|fp16 overrides). */
int b200pf_engine_create_prec(const char* model_dir, int device, int max_rows, int max_segments, int precision, b200pf_engine** out);
/* Batches of this engine that are still alive are orphaned: their device and pinned memory is released here, every later call
 * on them returns B200PF_ERR_INVALID, and b200pf_batch_destroy only frees the handle.  The creating entry points
 * (b200pf_engine_create*, b200pf_vad_create, b200pf_punc_create, b200pf_model_dir_probe) never unwind: a malformed model
 * directory becomes B200PF_ERR_IO with the text in b200pf_last_error(). */
void b200pf_engine_destroy(b200pf_engine* e);
int b200pf_engine_config(const b200pf_engine* e, b200pf_config* out);
/* Vocabulary access for the host-side detokeniser (Vocab, onnxruntime/src/vocab.cpp:46-63). */
int b200pf_engine_vocab_size(const b200pf_engine* e);
const char* b200pf_engine_token(const b200pf_engine* e, int id);
const char* b200pf_engine_lang(const b200pf_engine* e);
/* Options: "taps" (0/1) keeps fp32 intermediates of the next runs for b200pf_batch_tap; "overlap" (0 off, 1 FSMN enqueued
 * first, 2 = default: attention enqueued first) runs the FSMN memory block on a low-priority side stream concurrently
 * with the attention kernel;
 * "logprob_topk" = k (0..32, default 0) also produces pruned log-softmax posteriors per token (b200pf_result.topk_*);
 * "ffn_ln_fold" (default 1) applies the decoder feed-forward LayerNorm inside the w_1 / w_2 GEMM epilogues (0: as its own pass);
 * "profile" see below. */
int b200pf_engine_set_option(b200pf_engine* e, const char* key, int value);
/* Option "profile" = 1 brackets every launch of b200pf_batch_run with CUDA events on the launching stream.
 * b200pf_engine_profile_read synchronises, folds the finished brackets into 16 categories (names[i]:
 * frontend, layernorm, gemm_other, attention_tcgen05, fsmn, cif, argmax, other, gemm_qkv, gemm_out, gemm_ffn1,
 * gemm_ffn2, gemm_dec, gemm_vocab, two unused) and returns accumulated
 * milliseconds, algorithmic work (FLOPs for the two tensor-core categories, bytes for the rest) and launch
 * counts; `reset` clears the accumulators.  Arrays have 16 entries. */
int b200pf_engine_profile_read(b200pf_engine* e, int reset, const char** names, double* ms, double* work, long long* launches);
/* Small batches (rows <= option "graph_max_rows", default 4096; option "graphs" = 0 turns it off) run as CUDA graphs: the layout
 * is padded to a bucket (rows to a multiple of 32 with gap rows), the first batch of a bucket is launched kernel by kernel, the
 * second is captured and every later one replays the graph with one launch call.  Results are bit-identical either way.
 * Counters: graphs captured / replays so far / graphs currently cached. */
int b200pf_engine_graph_stats(const b200pf_engine* e, long long* captures, long long* replays, int* cached);
/* The CUDA stream (cudaStream_t) the engine enqueues on when `stream` arguments are NULL. */
void* b200pf_engine_stream(b200pf_engine* e);
/* A second stream owned by the engine, meant for b200pf_batch_stage_*: copies of the next batch then overlap the
 * compute of the current one (b200pf_batch_run waits for its batch's staging through an event). */
void* b200pf_engine_copy_stream(b200pf_engine* e);

/* Number of fbank frames / LFR frames for a segment of n samples (feature-window.cc:73-87,
 * paraformer.cpp:424).  Pure host arithmetic. */
int b200pf_num_fbank_frames(int64_t n_samples);
int b200pf_num_lfr_frames(int64_t n_samples);
/* Packed rows a batch with these sample counts occupies (sum of T_i + 1 over segments with T_i > 0). */
int64_t b200pf_rows_for(const int64_t* n_samples, int n_seg);

int b200pf_batch_create(b200pf_engine* e, int64_t max_samples, b200pf_batch** out);
void b200pf_batch_destroy(b200pf_batch* b);

/* Contextual models: the hotword embedding matrix Forward receives (hw_emb [n_hw][dim], model.h:31; the reference
 * replicates it per batch item, paraformer-torch.cpp:373-388).  dim must be d_model; it stays attached to the batch
 * until replaced.  A contextual model run without it fails like the reference ("hw_emb is null", paraformer.cpp:516-520). */
int b200pf_batch_set_hotwords(b200pf_batch* b, const float* hw_emb, int n_hw, int dim);
/* Replaces the model_eb.onnx session of Paraformer::CompileHotwordEmbedding (paraformer.cpp:651-688): Embedding +
 * LSTM over ids [n_words][max_len] (int32, zero padded) and returns the LSTM output at step lengths[j]-1 of word j:
 * out [n_words][d_model] fp32. */
int b200pf_engine_hotword_embed(b200pf_engine* e, const int32_t* ids, const int32_t* lengths, int n_words, int max_len, float* out);

/* Stage inputs (replaces the float copy + feature assembly of Paraformer::Forward, paraformer.cpp:482-532).
 * s16: `pcm` holds the segments back to back, segment i = pcm[offsets[i] .. offsets[i+1]) (the int16 the
 *      reference's LoadPcmwav turns into float/32768, audio.cpp:787-819).
 * f32: the reference's own argument form, din[i] in [-1,1) with len[i] samples (model.h:31).
 * Host buffers may be pageable or pinned; the copy is enqueued on `stream` (NULL -> engine stream). */
int b200pf_batch_stage_s16(b200pf_batch* b, const int16_t* pcm, const int64_t* offsets, int n_seg, void* stream);
int b200pf_batch_stage_f32(b200pf_batch* b, const float* const* din, const int* len, int n_seg, void* stream);
/* s16 from separate segments (seg[i] holds len[i] samples): no host-side gather, each segment is copied straight from where
 * the caller keeps it (e.g. the VAD cut points inside one long recording, in any order). */
int b200pf_batch_stage_s16_ptrs(b200pf_batch* b, const int16_t* const* seg, const int64_t* len, int n_seg, void* stream);
/* Enqueue the whole forward (fbank -> ... -> argmax) for the staged batch.  Asynchronous. */
int b200pf_batch_run(b200pf_batch* b, void* stream);
/* Copy results to the host (device->host on `stream`), synchronise, unpack.  Every result of a batch -- ids, fire frames,
 * us_alphas / us_peaks, pruned posteriors -- lives in the batch object, so run(A), run(B), collect(A) on one engine returns
 * A's own values.  Debug taps (b200pf_batch_tap) are the exception: they read the engine's workspace and are only valid until
 * the next run on that engine. */
int b200pf_batch_collect(b200pf_batch* b, b200pf_result* res, void* stream);
/* stage + run + collect. */
int b200pf_forward_s16(b200pf_batch* b, const int16_t* pcm, const int64_t* offsets, int n_seg, b200pf_result* res);
int b200pf_forward_f32(b200pf_batch* b, const float* const* din, const int* len, int n_seg, b200pf_result* res);
/* Kernel launches enqueued by the last b200pf_batch_run on this batch. */
int64_t b200pf_batch_launches(const b200pf_batch* b);
/* Algorithmic FLOPs (2*M*N*K per contraction) of the last collected batch. */
double b200pf_batch_flops(const b200pf_batch* b);

/* Debug taps (engine option "taps" = 1, after b200pf_batch_collect).  Copies the named fp32 intermediate
 * of segment `seg` to `out` (capacity `cap` floats) and writes its shape.  Names: "fbank" [n_fb,80],
 * "feats" [T,560], "enc" [T,512], "alphas" [T+1], "fires" [T+1], "embeds" [L,512], "logits" [L,vocab]. */
int b200pf_batch_tap(b200pf_batch* b, const char* name, int seg, float* out, int64_t cap, int64_t shape[2]);

/* ---- FSMN-VAD scores ------------------------------------------------------------------------
 * Replaces the onnxruntime session of FsmnVad::Forward (onnxruntime/src/fsmn-vad.cpp:72-135) and its front end (FbankKaldi /
 * LfrCmvn, :137-224) for whole recordings; the E2E state machine that turns scores into segments (e2e-vad.h) stays on the
 * host and reads scores[t][0] (sil_pdf_ids = {0}).  <vad_dir> holds am.mvn (400-dim) and vad.b200pf (upstream FunASR
 * fsmn_vad encoder parameter names).  max_frames: 10 ms frames one call may hold (0 -> 400000). */
typedef struct b200pf_vad b200pf_vad;
int b200pf_vad_create(const char* vad_dir, int device, int max_frames, b200pf_vad** out);
void b200pf_vad_destroy(b200pf_vad* v);
/* pcm: recordings back to back, recording i = pcm[offsets[i] .. offsets[i+1]).  frame_off [n_rec + 1] receives the first
 * frame of every recording (T_i = 1 + (n_i - 400) / 160 frames, 0 when shorter than one window); sil_prob [cap_frames]
 * the probability of pdf 0 per frame; all_probs (optional) [cap_frames, 248]; feats (optional) [cap_frames, 400] the
 * LFR + CMVN features (debug / parity). */
int b200pf_vad_scores_s16(b200pf_vad* v, const int16_t* pcm, const int64_t* offsets, int n_rec, float* sil_prob, int64_t cap_frames,
                          int32_t* frame_off, float* all_probs, float* feats);

/* ---- CT-Transformer punctuation network ------------------------------------------------------
 * Replaces the onnxruntime session of CTTransformer::Infer (onnxruntime/src/ct-transformer.cpp:164-203): int32 token ids in, one
 * punctuation class per token out, for many token sequences per call (the reference runs one 20-token mini-sentence of one
 * request per session call).  <punc_dir> holds punc.b200pf (upstream FunASR CTTransformer parameter names: embed, encoder.*,
 * decoder; config keys vocab, d_model, n_heads, d_ff, n_layers, kernel, n_punc).  max_tokens: tokens one call may hold (0 -> 65536). */
typedef struct b200pf_punc b200pf_punc;
int b200pf_punc_create(const char* punc_dir, int device, int max_tokens, b200pf_punc** out);
void b200pf_punc_destroy(b200pf_punc* p);
int b200pf_punc_info(const b200pf_punc* p, int* vocab, int* n_punc, int* d_model, int* max_tokens);
/* ids: sequences back to back, sequence i = ids[offsets[i] .. offsets[i+1]) (each <= 4096 tokens).  punc_out [offsets[n_seq]]
 * receives the first maximum over classes [0, n_punc - 1) of every token -- the reference's Argmax(row, row + CANDIDATE_NUM - 1)
 * never selects the last class (ct-transformer.cpp:191-195).  logits_out (optional) [offsets[n_seq], n_punc]. */
int b200pf_punc_infer(b200pf_punc* p, const int32_t* ids, const int32_t* offsets, int n_seq, int32_t* punc_out, float* logits_out);
/* The realtime model (CTTransformerOnline::Infer, ct-transformer-online.cpp:139-217): the same call with the reference's VadMask per
 * sequence -- with 0 < vad_pos[i] < T_i, tokens before vad_pos[i] - 1 do not attend to tokens from vad_pos[i] on (:219-233; the
 * reference feeds this mask to BOTH mask inputs of its session).  vad_pos NULL = no mask.  The realtime model's causal FSMN comes
 * from the config key sanm_shift of punc.b200pf (left context (kernel - 1) / 2 + sanm_shift). */
int b200pf_punc_infer_vad(b200pf_punc* p, const int32_t* ids, const int32_t* offsets, const int32_t* vad_pos, int n_seq, int32_t* punc_out,
                          float* logits_out);
/* Kernels launched by b200pf_punc_infer so far. */
long long b200pf_punc_launches(const b200pf_punc* p);

/* ---- single-operator entry points (fp32 host buffers in/out; used by the parity tests) -------------- */
/* 16-bit operand format the b200pf_op_* calls of this process run in (B200PF_PREC_*; default FP16, like the engine).  Wherever
 * the comments below say "bf16" read "the selected 16-bit format". */
int b200pf_op_set_precision(int precision);
/* C = A[M,K] * W[N,K]^T (+bias) (+relu: 1 after bias, 2 after all adds) (+add[M,N] rounded to bf16)
 * (+res[M,N] fp32).  A and W are rounded to bf16 on upload.  out_bf16_round: 1 rounds the result to bf16; 0 = fp32 with the
 * residual applied in place (the TMA epilogues, as in the forward); 2 = fp32 through the general epilogue.
 * argmax_out (optional, [M]) receives the fused greedy argmax (first maximum wins, util.cpp:63-74). */
int b200pf_op_gemm(int device, const float* A, const float* W, const float* bias, const float* add, const float* res,
                   int M, int N, int K, int relu, int out_bf16_round, float* out, int32_t* argmax_out);
/* Times `iters` back-to-back launches of the GEMM on device-resident dummy operands (CUDA events); mode 0: bias ->
 * bf16, 1: bias+ReLU -> bf16, 2: bias + fp32 residual in place, 3: bias + bf16 addend + fp32 residual in place,
 * 4: bias + fused argmax only.  ms_out = average milliseconds per launch. */
int b200pf_op_gemm_bench(int device, int M, int N, int K, int mode, int iters, float* ms_out);
/* conv1d k=3 pad 1 over packed rows with zero rows at segment gaps, as three shifted K passes. */
int b200pf_op_conv3(int device, const float* X, const float* Wr, const float* bias, int M, int C, float* out);
int b200pf_op_layernorm(int device, const float* x, int rows, int D, const float* gamma, const float* beta, float eps,
                        int in_bf16, float* out_f32, float* out_bf16_as_f32);
/* q [sum Tq,H*128], k,v [sum Tk,H*128]; segment s owns rows q_off[s].. and kv_off[s]..; impl 0 = tcgen05
 * kernel (product: persistent, single pass with a running maximum, all heads of a work item pipelined through one CTA),
 * 1 = CUDA-core cross-check. */
int b200pf_op_attention(int device, const float* q, const float* k, const float* v, const int32_t* q_off,
                        const int32_t* q_len, const int32_t* kv_off, const int32_t* kv_len, int n_seg, int n_heads,
                        int64_t q_rows, int64_t kv_rows, int impl, float* out);
/* Times `iters` launches of the attention kernel on device-resident random operands laid out like the engine's buffers:
 * segments of seg_T[i] LFR frames; cross = 0 self-attention (fused QKV rows), 1 = decoder cross-attention with (T+1)/2 query
 * rows per segment.  ms_out = milliseconds per launch, flops_out = 4 * sum(Tq * Tk) * 128 * n_heads. */
int b200pf_op_attention_bench(int device, const int32_t* seg_T, int n_seg, int n_heads, int cross, int impl, int iters, float* ms_out,
                              double* flops_out);
/* depthwise k=11 conv + identity over segments seg_off[n_seg+1]; x [rows,512], w [512,11]. */
int b200pf_op_fsmn(int device, const float* x, const float* w, const int32_t* seg_off, int n_seg, float* out);
/* CIF: alphas [sum (T_i+1)] (tail already appended), hidden [sum (T_i+1), 512]; seg_off[n_seg+1] over those rows.
 * Outputs: n_tok [n_seg], fires [rows], embeds [cap_tok,512], fire_frames [cap_tok]. */
int b200pf_op_cif(int device, const float* alphas, const float* hidden, const int32_t* seg_off, int n_seg,
                  float threshold, int32_t* n_tok, float* fires, float* embeds, int32_t* fire_frames, int64_t cap_tok);
/* LSTM (torch.nn.LSTM, one layer, hidden = input = 512, gate order i,f,g,o) over sequences x[seq_off[s] .. +seq_len[s]);
 * n_dir = 2 adds the reverse direction; weights [n_dir][2048][512], biases [n_dir][2048]; out [rows][512 n_dir] fp32
 * (bf16_out: the bf16 output path, widened). */
int b200pf_op_lstm(int device, const float* x, int rows, const int32_t* seq_off, const int32_t* seq_len, int n_seq, int n_dir,
                   const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int bf16_out, float* out);
/* Times the LSTM recurrence alone: n_seq sequences of `len` steps on random operands; ms_out = milliseconds per launch,
 * max_clusters = co-resident 16-CTA clusters the device grants the kernel. */
int b200pf_op_lstm_bench(int device, int n_seq, int len, int n_dir, int iters, float* ms_out, int* max_clusters);
/* us_alphas / us_cif_peak from raw alpha2 (already relu(sigmoid*s-n)): per segment scale to n_tok and cif_wo_hidden. */
int b200pf_op_us_peaks(int device, const float* alpha2, const int32_t* seq_off, const int32_t* seq_len, const int32_t* n_tok,
                       int n_seg, int rows, float threshold, float* us_alphas, float* us_peaks);
/* logsumexp + top-k log-softmax of logits [rows, V] (V % 4 == 0): lse [rows], lp / ids [rows, k]. */
int b200pf_op_logprob_topk(int device, const float* logits, int rows, int V, int k, float* lse, float* lp, int32_t* ids);
/* fbank + LFR/CMVN of one segment (int16 PCM); fb_out [n_fb,80], feats_out [T,560] (either may be NULL). */
int b200pf_op_frontend(b200pf_engine* e, const int16_t* pcm, int64_t n, float* fb_out, float* feats_out);

#ifdef __cplusplus
}
#endif
#endif /* B200PF_H_ */
