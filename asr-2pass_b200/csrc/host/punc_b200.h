// CTTransformerB200 — the reference's punctuation model interface (funasr::PuncModel / funasr::CTTransformer,
// onnxruntime/include/punc-model.h:10-21, onnxruntime/src/ct-transformer.{h,cpp}) over the B200 punctuation engine
// (include/b200pf.h b200pf_punc_*).
//
// The reference punctuates one request at a time: AddPunc walks the text in 20-token mini-sentences, and every step is one
// onnxruntime call whose input depends on the previous step's output (the unfinished sentence is carried over).  That chain
// cannot be shortened, but it can be shared: every request -- from AddPunc on any thread, or the many of one AddPuncBatch call --
// becomes a job of ONE dispatcher thread that advances all jobs in lock step, one engine call per round carrying the current
// mini-sentence of every job that still has one.  Requests join at any round and leave when their text is finished (continuous
// batching), so the decoder threads of a server that reach punc_handle->AddPunc at the same time (funasrruntime.cpp:317-320)
// share the network's ~40 launches per round instead of queueing behind each other.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../../include/b200pf.h"

namespace funasr_b200 {

class PuncModel {  // funasr::PuncModel, punc-model.h:10-21
 public:
  virtual ~PuncModel() {}
  virtual void InitPunc(const std::string& punc_model, const std::string& punc_config, const std::string& token_file, int thread_num) = 0;
  virtual std::string AddPunc(const char* sz_input, std::string language = "zh-cn") { (void)sz_input; (void)language; return ""; }
  virtual std::string AddPunc(const char* sz_input, std::vector<std::string>& arr_cache, std::string language = "zh-cn") {
    (void)sz_input; (void)arr_cache; (void)language;
    return "";
  }
  bool is_online = false;
};

// funasr::CTokenizer without the jieba branch (tokenizer.cpp): text -> pieces (original case) + ids (lower-cased lookup, <unk> otherwise).
class PuncTokenizer {
 public:
  void Open(const std::vector<std::string>& tokens, const std::vector<std::string>& punc_list);
  void Tokenize(const char* text, std::vector<std::string>* pieces, std::vector<int32_t>* ids) const;
  const std::string& Id2Punc(int id) const { return punc_[(size_t)id]; }
  int NumPunc() const { return (int)punc_.size(); }
  int NumTokens() const { return n_tokens_; }

 private:
  std::unordered_map<std::string, int> token2id_;
  std::vector<std::string> punc_;
  int unk_ = 0, n_tokens_ = 0;
};

// One request inside AddPunc / AddPuncBatch: the mini-sentence walk of ct-transformer.cpp:40-157 as a resumable state machine.
class PuncJob {
 public:
  PuncJob(const PuncTokenizer* tok, const char* text, const std::string& language);
  bool Active() const { return pos_ < ids_.size(); }
  const std::vector<int32_t>& Input() const { return in_ids_; }   // what the next network call must see (cache + next 20 tokens)
  void Consume(const int32_t* punc, int n);                         // the network's classes for Input()
  std::string Result() const;

 private:
  void Prepare();
  const PuncTokenizer* tok_;
  std::string language_;
  std::vector<std::string> pieces_, remain_str_, in_str_, new_str_, sent_out_;
  std::vector<int32_t> ids_, remain_ids_, in_ids_;
  size_t pos_ = 0;
  int n_total_ = 0;
};

class CTTransformerB200 : public PuncModel {
 public:
  explicit CTTransformerB200(int device = 0, int max_tokens = 0) : device_(device), max_tokens_(max_tokens) {}
  ~CTTransformerB200() override;
  // <punc_dir>/{punc.b200pf, tokens.json, punc_list.json}
  bool Init(const std::string& punc_dir, std::string* err);
  // reference signature: the directory of punc_model is used; config.yaml's punc_list must have been converted to punc_list.json
  void InitPunc(const std::string& punc_model, const std::string& punc_config, const std::string& token_file, int thread_num) override;
  std::string AddPunc(const char* sz_input, std::string language = "zh-cn") override;
  std::string AddPunc(const char* sz_input, std::vector<std::string>& arr_cache, std::string language = "zh-cn") override {
    (void)arr_cache;                      // ct-transformer.cpp:159-161: the offline model ignores the cache
    return AddPunc(sz_input, language);
  }
  // Many requests in lock step; results in input order.  `rounds` (optional) receives the number of engine calls made.
  std::vector<std::string> AddPuncBatch(const std::vector<std::string>& texts, const std::string& language = "zh-cn", int* rounds = nullptr);
  const PuncTokenizer& tokenizer() const { return tok_; }
  b200pf_punc* engine() const { return engine_; }
  long long rounds() const;   // engine calls made so far

 private:
  struct Ticket;
  void Run();
  void RunRound(std::vector<Ticket*>* active);
  int device_, max_tokens_;
  b200pf_punc* engine_ = nullptr;
  PuncTokenizer tok_;
  mutable std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  std::deque<Ticket*> pending_;
  bool stop_ = false;
  long long rounds_ = 0;
  std::thread worker_;
};

// funasr::CTTransformerOnline (ct-transformer-online.{h,cpp}): the realtime punctuation model -- what the 2-pass server applies to
// both its online and its offline-leg results (funasrruntime.cpp:545,610).  AddPunc(text, cache) punctuates `cache + text`, returns
// the part that belongs to `text` and leaves in `cache` the words after the last sentence end.  The network sees the reference's
// VadMask (vad_pos = the number of cached words) in every layer and needs sanm_shift in punc.b200pf (b200pf_punc_infer_vad).
class CTTransformerOnlineB200 : public PuncModel {
 public:
  explicit CTTransformerOnlineB200(int device = 0, int max_tokens = 0) : device_(device), max_tokens_(max_tokens) { is_online = true; }
  ~CTTransformerOnlineB200() override;
  bool Init(const std::string& punc_dir, std::string* err);
  void InitPunc(const std::string& punc_model, const std::string& punc_config, const std::string& token_file, int thread_num) override;
  std::string AddPunc(const char* sz_input, std::vector<std::string>& arr_cache, std::string language = "zh-cn") override;
  const PuncTokenizer& tokenizer() const { return tok_; }

 private:
  int device_, max_tokens_;
  b200pf_punc* engine_ = nullptr;
  PuncTokenizer tok_;
};

// The realtime walk with any network: infer(ids, vad_pos) -> class per token.
std::string AddPuncOnlineWith(const PuncTokenizer& tok, const char* text, std::vector<std::string>* cache,
                              const std::function<std::vector<int32_t>(const std::vector<int32_t>&, int)>& infer);

// The walk with any network (tests drive it with a scripted one on machines without a GPU).
std::string AddPuncWith(const PuncTokenizer& tok, const char* text, const std::string& language,
                        const std::function<std::vector<int32_t>(const std::vector<int32_t>&)>& infer);

}  // namespace funasr_b200
