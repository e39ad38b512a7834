// Offline subset of the reference's exported API (onnxruntime/include/funasrruntime.h:60-138) re-exported with
// IDENTICAL C++ signatures, so callers such as websocket/bin/websocket-server.cpp:81-84 or
// onnxruntime/bin/funasr-onnx-offline-rtf.cpp:70-76 link against libfunasr_b200.so unchanged.
// The reference's header is C++ (std::map / std::string / std::vector cross the boundary); the
// true C ABI lives one level down in include/b200pf.h.
//
// What this shim does NOT contain, by design: VAD, punctuation, ITN, resampling, ffmpeg
// decoding, WFST decoding.  Those stay on the reference's host path; the supported way to keep them is to
// plug ParaformerB200 into the reference's own OfflineStream.  Standalone, this shim treats
// the whole buffer as one segment (the reference's behaviour when no vad-dir is given,
// funasrruntime.cpp:241-245) and hard-cuts audio longer than vad_max_len.
#pragma once
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

typedef void* FUNASR_HANDLE;
typedef void* FUNASR_RESULT;
typedef void* FUNASR_DEC_HANDLE;

typedef enum { RASR_NONE = -1, RASRM_CTC_GREEDY_SEARCH = 0, RASRM_CTC_RPEFIX_BEAM_SEARCH = 1, RASRM_ATTENSION_RESCORING = 2 } FUNASR_MODE;
typedef enum { ASR_OFFLINE = 0, ASR_ONLINE = 1, ASR_TWO_PASS = 2 } ASR_TYPE;
typedef enum { PUNC_OFFLINE = 0, PUNC_ONLINE = 1 } PUNC_TYPE;   // funasrruntime.h:52-55
typedef void (*QM_CALLBACK)(int cur_step, int n_total);

// Plain-model API (funasrruntime.h:60-78): the handle is the funasr::Model itself, and FunASRInfer / FunASRInferBuffer feed
// the WHOLE audio to the single-segment overload Forward(float*, int, bool) (funasrruntime.cpp:57-114).  The reference's
// Paraformer does not override that overload, so it answers "" there; ParaformerB200 does (the batch-1 case).
FUNASR_HANDLE FunASRInit(std::map<std::string, std::string>& model_path, int thread_num, ASR_TYPE type = ASR_OFFLINE);
void FunASRReset(FUNASR_HANDLE handle, FUNASR_DEC_HANDLE dec_handle = nullptr);
FUNASR_RESULT FunASRInferBuffer(FUNASR_HANDLE handle, const char* sz_buf, int n_len, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                                bool input_finished = true, int sampling_rate = 16000, std::string wav_format = "pcm");
FUNASR_RESULT FunASRInfer(FUNASR_HANDLE handle, const char* sz_filename, FUNASR_MODE mode, QM_CALLBACK fn_callback, int sampling_rate = 16000);
void FunASRUninit(FUNASR_HANDLE handle);

// model_path keys (onnxruntime/include/com-define.h:15-38): "model-dir" is required; "quantize", "vad-dir",
// "punc-dir", "itn-dir", "lm-dir" are accepted and ignored here.  Extra keys: "device" (CUDA ordinal), "devices"
// ("0,1,...": one engine per listed GPU behind this handle, segments sharded over independent per-GPU queues),
// "max-rows", "max-segments", "micro-batch-us" (> 0: concurrent FunOfflineInfer* calls on this handle are merged into batched
// forwards by a MicroBatcher with that deadline in microseconds).
FUNASR_HANDLE FunOfflineInit(std::map<std::string, std::string>& model_path, int thread_num, bool use_gpu = false, int batch_size = 1);
void FunOfflineReset(FUNASR_HANDLE handle, FUNASR_DEC_HANDLE dec_handle = nullptr);
FUNASR_RESULT FunOfflineInferBuffer(FUNASR_HANDLE handle, const char* sz_buf, int n_len, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                                    const std::vector<std::vector<float>>& hw_emb, int sampling_rate = 16000,
                                    std::string wav_format = "pcm", bool itn = true, int vad_tail_sil = 800,
                                    int vad_max_len = 60000, FUNASR_DEC_HANDLE dec_handle = nullptr,
                                    std::string svs_lang = "auto", bool svs_itn = true);
FUNASR_RESULT FunOfflineInfer(FUNASR_HANDLE handle, const char* sz_filename, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                              const std::vector<std::vector<float>>& hw_emb, int sampling_rate = 16000, bool itn = true,
                              int vad_tail_sil = 800, int vad_max_len = 60000, FUNASR_DEC_HANDLE dec_handle = nullptr);
const std::vector<std::vector<float>> CompileHotwordEmbedding(FUNASR_HANDLE handle, std::string& hotwords, ASR_TYPE mode = ASR_OFFLINE);
void FunOfflineUninit(FUNASR_HANDLE handle);

// 2-pass stream (funasrruntime.h:121-132).  The OFFLINE leg of FunTpassInferBuffer -- VAD-closed segments through the offline
// acoustic model, realtime punctuation, timestamps and stamp_sents (funasrruntime.cpp:568-639) -- is served here, so
// websocket-server-2pass links against this library and its "offline" and "2pass" modes receive the corrected results
// ("mode":"2pass-offline").  The streaming acoustic model (ParaformerOnline, the "2pass-online" partial results) is outside this
// path: `FunASRGetResult` of a 2-pass result is empty and mode ASR_ONLINE returns no text.  model_path keys: "model-dir",
// "vad-dir" (required), "punc-dir" (the realtime CT-Transformer), plus the extra keys of FunOfflineInit ("device",
// "micro-batch-us": batch-1 calls of all connections merged by the MicroBatcher); "online-model-dir", "itn-dir", "lm-dir" are ignored.
FUNASR_HANDLE FunTpassInit(std::map<std::string, std::string>& model_path, int thread_num);
FUNASR_HANDLE FunTpassOnlineInit(FUNASR_HANDLE tpass_handle, std::vector<int> chunk_size = {5, 10, 5});
FUNASR_RESULT FunTpassInferBuffer(FUNASR_HANDLE handle, FUNASR_HANDLE online_handle, const char* sz_buf, int n_len,
                                  std::vector<std::vector<std::string>>& punc_cache, bool input_finished = true, int sampling_rate = 16000,
                                  std::string wav_format = "pcm", ASR_TYPE mode = ASR_TWO_PASS,
                                  const std::vector<std::vector<float>>& hw_emb = {{0.0}}, bool itn = true, int vad_tail_sil = 800,
                                  int vad_max_len = 60000, FUNASR_DEC_HANDLE dec_handle = nullptr, std::string svs_lang = "auto",
                                  bool svs_itn = true);
void FunTpassUninit(FUNASR_HANDLE handle);
void FunTpassOnlineUninit(FUNASR_HANDLE handle);

const char* FunASRGetResult(FUNASR_RESULT result, int n_index);
const char* FunASRGetStamp(FUNASR_RESULT result);
const char* FunASRGetStampSents(FUNASR_RESULT result);
const char* FunASRGetTpassResult(FUNASR_RESULT result, int n_index);
const int FunASRGetRetNumber(FUNASR_RESULT result);
void FunASRFreeResult(FUNASR_RESULT result);
const float FunASRGetRetSnippetTime(FUNASR_RESULT result);

// PUNC (funasrruntime.h:92-96): model_path["punc-dir"] = <dir>/{punc.b200pf, tokens.json, punc_list.json}; extra keys "device",
// "punc-max-tokens".  PUNC_OFFLINE -> CTTransformerB200, PUNC_ONLINE -> CTTransformerOnlineB200 (the realtime model with its word
// cache in the result object, funasrruntime.cpp:193-198).
FUNASR_HANDLE CTTransformerInit(std::map<std::string, std::string>& model_path, int thread_num, PUNC_TYPE type = PUNC_OFFLINE);
FUNASR_RESULT CTTransformerInfer(FUNASR_HANDLE handle, const char* sz_sentence, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                                 PUNC_TYPE type = PUNC_OFFLINE, FUNASR_RESULT pre_result = nullptr);
const char* CTTransformerGetResult(FUNASR_RESULT result, int n_index);
void CTTransformerFreeResult(FUNASR_RESULT result);
void CTTransformerUninit(FUNASR_HANDLE handle);

// The decoder handle is only meaningful with an LM; greedy search needs none (funasrruntime.cpp:836-894).
FUNASR_DEC_HANDLE FunASRWfstDecoderInit(FUNASR_HANDLE handle, int asr_type, float glob_beam, float lat_beam, float am_scale);
void FunASRWfstDecoderUninit(FUNASR_DEC_HANDLE handle);
void FunWfstDecoderLoadHwsRes(FUNASR_DEC_HANDLE handle, int inc_bias, std::unordered_map<std::string, int>& hws_map);
void FunWfstDecoderUnloadHwsRes(FUNASR_DEC_HANDLE handle);

namespace funasr_b200 { class ParaformerB200; class MultiGpuParaformer; class Model; }
// The acoustic model behind an offline handle (what OfflineStream::asr_handle is in the reference): the model object the
// handle forwards to (single engine or multi-GPU pool), its first engine, and the pool (nullptr on one GPU).
funasr_b200::Model* FunOfflineModel(FUNASR_HANDLE handle);
funasr_b200::ParaformerB200* FunOfflineModelB200(FUNASR_HANDLE handle);
funasr_b200::MultiGpuParaformer* FunOfflinePoolB200(FUNASR_HANDLE handle);

// With "vad-dir" in FunOfflineInit's map, FunOfflineInferBuffer cuts the recording with FSMN-VAD (scores on the GPU, the E2E state
// machine on the host) the way the reference's UseVad() branch does; this returns that cut alone: [start_ms, end_ms) pairs.
int FunOfflineVadSegmentsB200(FUNASR_HANDLE handle, const short* pcm, long long n_samples, int vad_tail_sil, int vad_max_len,
                              std::vector<std::pair<int, int>>* out);

// Extension for callers that already hold VAD cut points (e.g. the reference's Audio::CutSplit output):
// segment i = pcm[seg_begin[i] .. seg_end[i]) in samples.  Segments are length-sorted, batched with the
// reference's FetchDynamic rules, decoded, un-permuted and stitched exactly like FunOfflineInferBuffer.
FUNASR_RESULT FunOfflineInferSegmentsB200(FUNASR_HANDLE handle, const short* pcm, long long n_samples, const long long* seg_begin,
                                          const long long* seg_end, int n_seg, const std::vector<std::vector<float>>& hw_emb = {{0.0}});
