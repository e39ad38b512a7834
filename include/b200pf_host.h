/*
 * C hooks into the C++ host mirror (asr-2pass_b200/csrc/host): the reference-facing layer above b200pf.h.
 * They exist for ctypes-based tests and for bench.py's end-to-end leg; C++ callers use
 * funasrruntime_b200.h / paraformer_b200.h directly.  Exported by libfunasr_b200.so.
 */
#ifndef B200PF_HOST_H_
#define B200PF_HOST_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Vocab::Vector2StringV2 (onnxruntime/src/vocab.cpp:164-305); stateful across calls like the reference. */
void* b200pf_host_detok_create(const char* const* tokens, int n);
void b200pf_host_detok_destroy(void* h);
int b200pf_host_detok_text(void* h, const int32_t* ids, int n, const char* lang, char* out, int cap);
/* The same text for an explicit incoming detokeniser state (did the previous text end on a complete English word); *state_out is
 * the state the call leaves behind.  Does not touch the handle's own state (what the parallel text assembly relies on). */
int b200pf_host_detok_text_state(void* h, const int32_t* ids, int n, const char* lang, int state_in, int* state_out, char* out, int cap);
/* Vector2String + TimestampOnnx + PostProcess (vocab.cpp:98-104, util.cpp:838-963, :720-836). */
int b200pf_host_timestamp_text(void* h, const int32_t* ids, int n, const float* us_alphas, const float* us_peaks, int n_frames,
                               char* out, int cap);
/* Result stitching of FunOfflineInferBuffer (funasrruntime.cpp:291-316). */
int b200pf_host_stitch(const char* const* msgs, const float* start_s, int n, const char* lang, char* text, int text_cap, char* stamp,
                       int stamp_cap);

/* FunOfflineInit / FunOfflineInferBuffer / FunOfflineUninit (funasrruntime.h:100-117) through C types. */
void* b200pf_host_offline_init(const char* model_dir, int device, int max_rows, int max_segments, int batch_size);
/* Same over several GPUs of the box: one engine per device behind one handle, every call's segments are sharded over
 * independent per-GPU queues (funasr_b200::MultiGpuParaformer; no collective). */
void* b200pf_host_offline_init_devices(const char* model_dir, const int* devices, int n_dev, int max_rows, int max_segments, int batch_size);
/* FunOfflineInit with a VAD model ("vad-dir"; funasrruntime.cpp:35-39 / OfflineStream's vad_handle): FunOfflineInferBuffer then cuts
 * every recording with FSMN-VAD scores (GPU) + the E2E state machine (host) as the reference's UseVad() branch does
 * (funasrruntime.cpp:243-245, Audio::CutSplit audio.cpp:1172-1226).  speech_noise_thres <= 0: read <vad_dir>/config.yaml, else 0.6. */
void* b200pf_host_offline_init_vad(const char* model_dir, const char* vad_dir, int device, int max_rows, int max_segments, int batch_size,
                                   float speech_noise_thres);
/* The cut alone: seg_ms receives [start_ms, end_ms) pairs (cap pairs); returns the number of segments, < 0 on error. */
int b200pf_host_offline_vad_cut(void* h, const int16_t* pcm, int64_t n_samples, int vad_tail_sil, int vad_max_len, int* seg_ms, int cap);
/* FunOfflineInferBuffer with the reference's vad_tail_sil / vad_max_len arguments (funasrruntime.h:101-105); text and stamp out. */
int b200pf_host_offline_infer_buffer_vad(void* h, const char* buf, int n_bytes, int vad_tail_sil, int vad_max_len, char* text, int text_cap,
                                         char* stamp, int stamp_cap);
/* FunOfflineInit with the reference's key/value map (com-define.h: "model-dir", "vad-dir", "punc-dir", ... plus the B200 keys
 * "device", "devices", "max-rows", "max-segments", "micro-batch-us", "vad-speech-noise-thres", "punc-max-tokens"). */
void* b200pf_host_offline_init_kv(const char* const* keys, const char* const* values, int n, int batch_size);
/* FunOfflineInferBuffer -> FunASRGetResult / FunASRGetStamp / FunASRGetStampSents (stamp and stamp_sents may be NULL). */
int b200pf_host_offline_infer_full(void* h, const char* buf, int n_bytes, int vad_tail_sil, int vad_max_len, char* text, int text_cap,
                                   char* stamp, int stamp_cap, char* stamp_sents, int sents_cap);
/* 2-pass stream (funasrruntime.h:121-132; the offline leg of FunTpassInferBuffer, funasrruntime.cpp:568-639).  Keys as FunTpassInit
 * reads them ("model-dir", "vad-dir", "punc-dir", ...).  b200pf_host_tpass_infer = one FunTpassInferBuffer call on one connection:
 * mode 0 = ASR_OFFLINE, 1 = ASR_ONLINE, 2 = ASR_TWO_PASS; cache_io = punc_cache[1] as '\n'-joined words, in and out. */
void* b200pf_host_tpass_init_kv(const char* const* keys, const char* const* values, int n);
void* b200pf_host_tpass_online_init(void* tpass);
void b200pf_host_tpass_uninit(void* tpass);
void b200pf_host_tpass_online_uninit(void* online);
int b200pf_host_tpass_infer(void* tpass, void* online, const char* buf, int n_bytes, int input_finished, int mode, int vad_tail_sil,
                            int vad_max_len, char* cache_io, int cache_cap, char* msg, int msg_cap, char* tpass_msg, int tpass_cap, char* stamp,
                            int stamp_cap, char* stamp_sents, int sents_cap);
/* pf::host::StreamingVad fed chunk_len[k] frames per call, the last chunk final: [start_ms, end_ms] pairs, any chunking of the same
 * scores gives b200pf_host_vad_segments' result. */
int b200pf_host_vad_segments_streaming(const float* sil_prob, int n_frames, const int* chunk_len, int n_chunks, int max_end_sil_ms, int max_seg_ms,
                                       float thres, int* out, int cap);
/* `iters` calls of funasr::Model::Forward(float** din, int* len, ...) timed inside the library (mean milliseconds per call in *ms_out,
 * result strings built); returns the number of non-empty strings of the last call, -1 on a bad handle. */
int b200pf_host_model_forward_timed(void* h_offline, const float* const* din, const int* len, int n, int iters, double* ms_out);
/* pf::host::ExpandPrunedPosteriors: b200pf_result.topk_logprob / topk_ids [rows, k] -> dense log-softmax rows [rows, vocab] in the
 * layout WfstDecoder::Search (wfst-decoder.cpp:27-57) and CtcPrefixDecoder::CtcSearch (ctc-prefix-decoder.cpp:157) read. */
int b200pf_host_expand_posteriors(const float* topk_logprob, const int32_t* topk_ids, int rows, int k, int vocab, float* dense);
/* The pool's longest-processing-time-first assignment of segments (sample counts) to n_dev queues; host arithmetic only. */
int b200pf_host_partition(const int* len, int n, int n_dev, int* assign);
/* Segments decoded per device so far; returns the number of devices (0 for a single-GPU handle). */
int b200pf_host_segments_per_device(void* h_offline, long long* out, int cap);
void b200pf_host_offline_uninit(void* h);
int b200pf_host_offline_infer_buffer(void* h, const char* buf, int n_bytes, int vad_max_len, char* text, int text_cap, float* snippet_s);
int b200pf_host_offline_infer_segments(void* h, const int16_t* pcm, int64_t n_samples, const int64_t* seg_begin, const int64_t* seg_end,
                                       int n_seg, char* text, int text_cap);
/* funasr::Model::Forward(float**, int*, ...) (model.h:31) on the handle's ParaformerB200; strings joined by '\n'. */
int b200pf_host_model_forward(void* h_offline, const float* const* din, const int* len, int n, char* out, int cap);
/* Same with the hw_emb argument of Forward: hw [n_hw][dim] (contextual models, paraformer.cpp:515-531). */
int b200pf_host_model_forward_hw(void* h_offline, const float* const* din, const int* len, int n, const float* hw, int n_hw, int dim,
                                 char* out, int cap);
/* FunOfflineInferBuffer with hw_emb; also returns the stitched "[[b,e],...]" stamp string (funasrruntime.cpp:302-316). */
int b200pf_host_offline_infer_buffer_hw(void* h, const char* buf, int n_bytes, int vad_max_len, const float* hw, int n_hw, int dim,
                                        char* text, int text_cap, char* stamp, int stamp_cap);
/* CompileHotwordEmbedding(handle, hotwords) (funasrruntime.h:118; Paraformer::CompileHotwordEmbedding, paraformer.cpp:592-693):
 * writes rows of `dim` floats, returns the row count (hotwords kept + the blank row) or -1. */
int b200pf_host_compile_hotwords(void* h_offline, const char* hotwords, float* out, int cap_rows, int dim);
/* TimestampSentence (util.cpp:569-637): punctuated text + "[[b,e],...]" (ms) -> the stamp_sents JSON array FunASRGetStampSents returns. */
int b200pf_host_sentence_stamps(const char* text, const char* stamp, char* out, int cap);
/* ---- punctuation: funasr::CTTransformer mirror (csrc/host/punc_b200.h) -----------------------
 * CTokenizer::Tokenize without jieba (tokenizer.cpp:312-365): text -> token ids (lower-cased lookup, <unk> otherwise).  No GPU. */
int b200pf_host_punc_tokenize(const char* const* tokens, int n_tokens, const char* text, int32_t* ids, int cap);
/* CTTransformer::AddPunc's mini-sentence walk (ct-transformer.cpp:40-157) with a SCRIPTED network in place of the session (class =
 * hash of token id, position and seed; tests/test_punc.py::scripted_punc): pins the host logic on machines without a GPU. */
int b200pf_host_punc_add_scripted(const char* const* tokens, int n_tokens, const char* const* punc_list, int n_punc, const char* text,
                                  const char* lang, int seed, int every, char* out, int cap);
/* CTTransformerInit / CTTransformerInfer(PUNC_OFFLINE) / CTTransformerUninit (funasrruntime.h:92-96) over the B200 punctuation
 * engine: <punc_dir>/{punc.b200pf, tokens.json, punc_list.json}.  add_batch punctuates n texts in lock step (one engine call per
 * round for all of them); results are written back to back, each NUL-terminated; returns the bytes used. */
/* The realtime model, funasr::CTTransformerOnline (ct-transformer-online.cpp): AddPunc(text, cache).  The word cache travels as one
 * string, every word followed by '\x01'.  _scripted: the walk with the scripted network (seed + 13 * vad_pos), no GPU. */
int b200pf_host_punc_online_add_scripted(const char* const* tokens, int n_tokens, const char* const* punc_list, int n_punc, const char* text,
                                         const char* cache_in, int seed, int every, char* out, int cap, char* cache_out, int cache_cap);
void* b200pf_host_punc_online_create(const char* punc_dir, int device, int max_tokens);
void b200pf_host_punc_online_destroy(void* h);
int b200pf_host_punc_online_add(void* h, const char* text, const char* cache_in, char* out, int cap, char* cache_out, int cache_cap);
void* b200pf_host_punc_create(const char* punc_dir, int device, int max_tokens);
void b200pf_host_punc_destroy(void* h);
/* Engine calls (lock-step rounds) made so far: concurrent add calls share rounds. */
long long b200pf_host_punc_rounds(void* h);
int b200pf_host_punc_add(void* h, const char* text, const char* lang, char* out, int cap);
int b200pf_host_punc_add_batch(void* h, const char* const* texts, int n, const char* lang, char* out, int cap, int* rounds);
/* The host half of CompileHotwordEmbedding alone (paraformer.cpp:600-648; no GPU): hotword string -> ids [n][10] + lengths [n],
 * blank row last.  tokens = tokens.json in id order; seg_dict_path may be NULL.  Returns n, -1 when n > cap_rows. */
int b200pf_host_pack_hotwords(const char* const* tokens, int n_tokens, const char* seg_dict_path, const char* hotwords, int32_t* ids,
                              int32_t* lengths, int cap_rows);
/* Model::InitSegDict (model.h:27; SegDict, seg_dict.cpp:19-38) for English hotwords. */
int b200pf_host_init_seg_dict(void* h_offline, const char* path);

/* FunASRInit -> FunASRInfer (wav_path != NULL) or FunASRInferBuffer (buf) -> FunASRGetResult -> FunASRUninit
 * (funasrruntime.h:60-78; the plain-model API that funasr-onnx-offline style callers use).  Returns the text length,
 * -1 when Init fails, -2 when inference returns nullptr. */
int b200pf_host_funasr_infer(const char* model_dir, int device, int max_rows, const char* wav_path, const char* buf, int n_bytes, char* out, int cap);
/* pf::host::SegmentVad: what funasr::E2EVadModel (onnxruntime/src/e2e-vad.h) does for a whole recording with its default
 * options: sil_prob [n_frames] (probability of pdf 0 per 10 ms frame) -> [start_ms, end_ms] pairs in out[2*cap]; returns the
 * number of segments.  max_end_sil_ms = vad_tail_sil, max_seg_ms = vad_max_len, thres = speech_noise_thres. */
int b200pf_host_vad_segments(const float* sil_prob, int n_frames, int max_end_sil_ms, int max_seg_ms, float thres, int* out, int cap);
/* MicroBatcher (asr-2pass_b200/csrc/host/micro_batcher.h): merges the batch-1 Forward calls that the 2-pass server makes
 * per closed VAD segment (funasrruntime.cpp:570-586) across connections into batched forwards.  `mock` variant: host-only
 * inner model for tests.  forward blocks until the segment is decoded; hw may be NULL (n_hw 0). */
void* b200pf_host_mb_create(void* h_offline, int max_wait_us, int max_batch, int max_rows);
void* b200pf_host_mb_create_mock(int max_wait_us, int max_batch, int max_rows, int latency_us);
void b200pf_host_mb_destroy(void* mb);
int b200pf_host_mb_forward(void* mb, const float* pcm, int len, const float* hw, int n_hw, int dim, char* out, int cap);
/* out7: segments, batches, closed_by_deadline, closed_by_size, max_batch_seen, mean wait us, max wait us */
int b200pf_host_mb_stats(void* mb, double* out7);

#ifdef __cplusplus
}
#endif
#endif
