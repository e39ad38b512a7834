// ParaformerB200 — the B200-native offline Paraformer behind the reference's plugin seam.
//
// The reference selects its acoustic model through the virtual class funasr::Model
// (onnxruntime/include/model.h:13-46), instantiated in OfflineStream::OfflineStream
// (onnxruntime/src/offline-stream.cpp:40-57).  `Model` below repeats, with the same names, argument order
// and meaning, the virtuals that the offline Paraformer path uses, so that a maintainer can make
// ParaformerB200 derive from funasr::Model by changing one base-class name (INTEGRATION.md shows the
// patch).  All arithmetic happens in the CUDA engine behind include/b200pf.h; this file is host plumbing:
// argument marshalling, batching by the engine's capacity, greedy-search string assembly.
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/b200pf.h"
#include "text.h"

#ifdef B200PF_WITH_REFERENCE_HEADERS
// Built INSIDE the reference tree: the base classes are the reference's own funasr::Model
// (onnxruntime/include/model.h:13-46) and funasr::WfstDecodable (onnxruntime/src/wfst-decodable.h:17-28), so that
// OfflineStream / TpassStream can hold a ParaformerB200 in their asr_handle and FunASRWfstDecoderInit's dynamic_cast
// (funasrruntime.cpp:841) finds the decoder interface.  tests/test_boundary_cpu.py compiles this variant against the reference's
// headers.
#include "model.h"
#include "wfst-decodable.h"
namespace funasr_b200 {
using Model = funasr::Model;
}
#define B200PF_MODEL_BASES public funasr::Model, public funasr::WfstDecodable
#else
namespace funasr_b200 {

class Model {  // subset of funasr::Model used on the offline Paraformer path (model.h:13-46)
 public:
  virtual ~Model() {}
  virtual void StartUtterance() = 0;
  virtual void EndUtterance() = 0;
  virtual void Reset() = 0;
  virtual void InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config,
                       const std::string& token_file, int thread_num) = 0;
  virtual std::vector<std::string> Forward(float** din, int* len, bool input_finished,
                                           const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                                           void* wfst_decoder = nullptr, int batch_in = 1) = 0;
  virtual std::string Forward(float* din, int len, bool input_finished,
                              const std::vector<std::vector<float>>& hw_emb = {{0.0}}, void* wfst_decoder = nullptr) = 0;
  virtual std::string Rescoring() = 0;
  virtual void InitHwCompiler(const std::string& hw_model, int thread_num) {}
  virtual void InitSegDict(const std::string& seg_dict_model) {}
  virtual std::vector<std::vector<float>> CompileHotwordEmbedding(std::string& hotwords) = 0;
  virtual std::string GetLang() = 0;
  virtual int GetAsrSampleRate() = 0;
  virtual void SetBatchSize(int batch_size) = 0;
  virtual int GetBatchSize() = 0;
};

}  // namespace funasr_b200
#define B200PF_MODEL_BASES public Model
#endif

namespace funasr_b200 {

// Host half of Paraformer::CompileHotwordEmbedding (paraformer.cpp:600-648): hotword string -> id matrix [n, B200PF_HOTWORD_LEN]
// and lengths, blank row last.  Pure host code (pinned against the reference's compiled function, tests/test_host_cpu.py).
void PackHotwords(const std::string& hotwords, const std::unordered_map<std::string, int>& token_id,
                  const std::unordered_map<std::string, std::vector<std::string>>& seg_dict, std::vector<int32_t>* matrix,
                  std::vector<int32_t>* lengths);
// SegDict::SegDict (seg_dict.cpp:19-38): "word \t piece piece ..." per line.
void LoadSegDict(const std::string& path, std::unordered_map<std::string, std::vector<std::string>>* out);

// What the engine returns for one segment before it becomes text: greedy token ids (and, for timestamp models, the upsampled
// alphas / peaks).  MultiGpuParaformer collects these from its per-GPU workers and turns them into text itself, in the CALLER's
// order, through ONE detokeniser -- the reference's Vocab carries state from one call to the next (last_is_complete_english_,
// vocab.cpp:177,259-261,283-288: whether the next text starts with a space), so text built per GPU share would differ from a
// one-GPU handle's in exactly those spaces.
struct SegmentRaw {
  bool has_features = false;   // false: too short for one frame -> "" (paraformer.cpp:477-480)
  std::vector<int> ids;
  std::vector<float> us_alphas, us_peaks;
  // Built by the worker that decoded the segment: its text for either incoming detokeniser state (previous text ended on a complete
  // English word: no / yes) and the state it leaves behind -- the pool's serial pass in the caller's order only picks one.
  bool has_text = false;
  std::string text[2];
  bool ended[2] = {false, true};
};

class ParaformerB200 : B200PF_MODEL_BASES {
 public:
  // device: CUDA ordinal; max_rows / max_segments: engine capacity (0 = defaults)
  explicit ParaformerB200(int device = 0, int max_rows = 0, int max_segments = 0);
  ~ParaformerB200() override;

  // am_model names a file inside the model directory (the reference passes <dir>/model.onnx or
  // model_quant.onnx, offline-stream.cpp:74-89); this implementation reads <dir>/model.b200pf next to it.
  // Exits the process on failure exactly as the reference does (paraformer.cpp:43-46).
  void InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config,
               const std::string& token_file, int thread_num) override;
  // Non-exiting variant used by the C hooks and tests.
  bool Init(const std::string& model_dir, std::string* err);

  std::vector<std::string> Forward(float** din, int* len, bool input_finished,
                                   const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                                   void* wfst_decoder = nullptr, int batch_in = 1) override;
  // The reference's Paraformer does not override the single-segment overload, so FunASRInfer returns ""
  // (SURVEY.md §8(b)); here it is the batch-1 case of the batched Forward.
  std::string Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                      void* wfst_decoder = nullptr) override;
  // int16 entry used by the buffer API when no resampling is needed (skips the float round trip of
  // Audio::LoadPcmwav, audio.cpp:787-819, which is exact: int16 -> float/32768 -> *32768).
  std::vector<std::string> ForwardPcm16(const int16_t* pcm, const int64_t* offsets, int n_seg,
                                        const std::vector<std::vector<float>>& hw_emb = {{0.0}});
  // int16 segments that live anywhere (e.g. VAD cut points of one recording, any order): copied to the GPU from where they
  // are, without a host-side gather.
  std::vector<std::string> ForwardSegments16(const int16_t* const* seg, const int64_t* len, int n_seg,
                                             const std::vector<std::vector<float>>& hw_emb = {{0.0}});

  // The same forwards stopped before detokenisation (greedy path only; an LM decoder handle is not taken here).
  std::vector<SegmentRaw> ForwardRaw(float** din, int* len, int batch_in, const std::vector<std::vector<float>>& hw_emb = {{0.0}});
  std::vector<SegmentRaw> ForwardSegments16Raw(const int16_t* const* seg, const int64_t* len, int n_seg,
                                               const std::vector<std::vector<float>>& hw_emb = {{0.0}});
  // GreedySearch's text for one segment (paraformer.cpp:386-407); advances this object's detokeniser state like the reference's Vocab.
  std::string TextOf(const SegmentRaw& seg);

  void StartUtterance() override {}
  void EndUtterance() override {}
  void Reset() override {}
  std::string Rescoring() override { return ""; }
  // Contextual models: hotword string -> ids -> Embedding + LSTM on the GPU -> one 512-vector per hotword plus the
  // blank row (paraformer.cpp:592-693).  Other models: {{0 x 512}} (paraformer.cpp:595-599).
  std::vector<std::vector<float>> CompileHotwordEmbedding(std::string& hotwords) override;
  // The reference loads model_eb.onnx here and flips use_hotword (paraformer.cpp:243-272).  The B200 model file
  // already carries the hotword compiler's weights, so this only checks that it does.
  void InitHwCompiler(const std::string& hw_model, int thread_num) override;
  void InitSegDict(const std::string& seg_dict_model) override;  // SegDict::SegDict, seg_dict.cpp:19-38
  bool use_hotword() const { return use_hotword_; }
  bool has_timestamp() const { return has_timestamp_; }
  std::string GetLang() override { return language_; }
  int GetAsrSampleRate() override { return sample_rate_; }
  void SetBatchSize(int batch_size) override { batch_size_ = batch_size; }
  int GetBatchSize() override { return batch_size_; }

#ifdef B200PF_WITH_REFERENCE_HEADERS
  // funasr::WfstDecodable (wfst-decodable.h:17-28), as the reference's Paraformer implements it (paraformer.h:77-80): the
  // reference's own Vocab / PhoneSet over the model directory's tokens.json; no LM is loaded here, so FunASRWfstDecoderInit
  // builds the CtcPrefixDecoder branch (funasrruntime.cpp:843-850).
  std::shared_ptr<fst::Fst<fst::StdArc>> GetLm() const override { return lm_; }
  funasr::Vocab* GetVocab() const override { return ref_vocab_.get(); }
  funasr::PhoneSet* GetPhoneSet() const override { return ref_phone_set_.get(); }
  funasr::Vocab* GetLmVocab() const override { return lm_vocab; }
  funasr::Vocab* GetVocab() override { return ref_vocab_.get(); }             // funasr::Model's non-const accessors (model.h:40-42)
  funasr::PhoneSet* GetPhoneSet() override { return ref_phone_set_.get(); }
  // Paraformer::InitLm (paraformer.cpp:156-176): TLG.fst + lexicon.  With an LM loaded, Forward(..., wfst_decoder) feeds the
  // reference's own WfstDecoder::Search with rows rebuilt from the engine's pruned posteriors (logprob_adapter.h) instead of
  // the greedy ids -- the branch at paraformer.cpp:565-578.
  void InitLm(const std::string& lm_file, const std::string& lm_cfg_file, const std::string& lex_file, const std::string& lm_units_file) override;
#endif

  b200pf_engine* engine() { return engine_; }
  pf::host::Detokenizer* vocab() { return vocab_.get(); }
  // Segments of this THREAD's last Forward / ForwardPcm16 / ForwardSegments16 call that could not be decoded (their strings are
  // "", as the reference returns for an inference exception, paraformer.cpp:582-587) -- e.g. a segment that does not fit the
  // engine's capacity.  Lets the buffer APIs tell "nothing recognised" from "not decoded".
  static int last_failed_segments();
  // token ids / CIF fire frames of the last Forward on this thread's call (debug / tests)
  const std::vector<std::vector<int>>& last_ids() const { return last_ids_; }

 private:
  std::vector<std::string> Decode(const b200pf_result& r, int n_seg, void* wfst_decoder, SegmentRaw* raw = nullptr);
  // engine-sized sub-batches through the C ABI, double buffered: pcm16 (offsets) or float (din/len)
  struct Slot {
    b200pf_batch* batch = nullptr;
    int64_t samples = 0;
    bool hw_valid = false;           // hw_flat is what is attached to `batch`
    std::vector<float> hw_flat;
  };
  bool StageSlot(int k, const int16_t* pcm, const int64_t* offsets, float** din, int* len, int n, int64_t samples,
                 const std::vector<std::vector<float>>& hw_emb, const int16_t* const* seg16 = nullptr, const int64_t* len16 = nullptr);
  bool CollectSlot(int k, int n, std::vector<std::string>* out, void* wfst_decoder, SegmentRaw* raw = nullptr);
  // raw != nullptr: [n_seg] records are filled instead of text (the returned strings stay empty)
  std::vector<std::string> RunAll(const int16_t* pcm, const int64_t* offsets, float** din, int* len, int n_seg,
                                  const std::vector<std::vector<float>>& hw_emb, const int16_t* const* seg16 = nullptr,
                                  const int64_t* len16 = nullptr, void* wfst_decoder = nullptr, SegmentRaw* raw = nullptr);
  bool UseLmDecoder(void* wfst_decoder) const;   // an LM is loaded and the caller passed a decoder handle (paraformer.cpp:565)

  int device_, max_rows_, max_segments_;
  b200pf_engine* engine_ = nullptr;
  Slot slots_[2];
  std::unique_ptr<pf::host::Detokenizer> vocab_;
  std::string language_ = "zh-cn";
  int sample_rate_ = 16000;
  int batch_size_ = 1;
  std::mutex mu_;  // Forward is called concurrently by decoder threads on one handle (websocket-server.cpp:387-403)
  std::vector<std::vector<int>> last_ids_;
  bool use_hotword_ = false, has_timestamp_ = false;
  int topk_on_ = 0;   // engine option "logprob_topk" as this object last set it
  int d_model_ = 512;
  std::unordered_map<std::string, std::vector<std::string>> seg_dict_;
  std::unordered_map<std::string, int> token_id_;  // PhoneSet::String2Id (phone-set.cpp:38-68)
#ifdef B200PF_WITH_REFERENCE_HEADERS
  std::unique_ptr<funasr::Vocab> ref_vocab_;
  std::unique_ptr<funasr::PhoneSet> ref_phone_set_;
#endif
};

}  // namespace funasr_b200
