#include "paraformer_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace funasr_b200 {

namespace {
std::string DirOf(const std::string& path) {
  const size_t p = path.find_last_of('/');
  return p == std::string::npos ? std::string(".") : path.substr(0, p);
}
}  // namespace

ParaformerB200::ParaformerB200(int device, int max_rows, int max_segments)
    : device_(device), max_rows_(max_rows), max_segments_(max_segments) {}

ParaformerB200::~ParaformerB200() {
  if (batch_) b200pf_batch_destroy(batch_);
  if (engine_) b200pf_engine_destroy(engine_);
}

bool ParaformerB200::Init(const std::string& model_dir, std::string* err) {
  if (b200pf_engine_create(model_dir.c_str(), device_, max_rows_, max_segments_, &engine_) != 0) {
    if (err) *err = b200pf_last_error();
    return false;
  }
  b200pf_config cfg;
  b200pf_engine_config(engine_, &cfg);
  sample_rate_ = cfg.sample_rate;
  language_ = b200pf_engine_lang(engine_);
  std::vector<std::string> toks(b200pf_engine_vocab_size(engine_));
  for (size_t i = 0; i < toks.size(); ++i) toks[i] = b200pf_engine_token(engine_, (int)i);
  vocab_.reset(new pf::host::Detokenizer(std::move(toks)));
  max_rows_ = cfg.max_rows;
  max_segments_ = cfg.max_segments;
  return true;
}

void ParaformerB200::InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config,
                             const std::string& token_file, int thread_num) {
  (void)am_cmvn; (void)am_config; (void)token_file; (void)thread_num;  // same directory; no host threads needed
  std::string err;
  if (!Init(DirOf(am_model), &err)) {
    fprintf(stderr, "Error when load am b200pf model: %s\n", err.c_str());
    exit(-1);
  }
}

std::vector<std::vector<float>> ParaformerB200::CompileHotwordEmbedding(std::string& hotwords) {
  (void)hotwords;
  return std::vector<std::vector<float>>(1, std::vector<float>(512, 0.0f));
}

std::vector<std::string> ParaformerB200::Decode(const b200pf_result& r, int n_seg) {
  std::vector<std::string> out(n_seg);
  last_ids_.assign(n_seg, std::vector<int>());
  for (int i = 0; i < n_seg; ++i) {
    const int cnt = r.token_counts[i];
    if (r.lfr_frames[i] <= 0) continue;  // empty features -> "" (paraformer.cpp:477-480)
    std::vector<int> ids(r.token_ids + r.token_offsets[i], r.token_ids + r.token_offsets[i] + cnt);
    out[i] = vocab_->ToText(ids, language_);  // GreedySearch, paraformer.cpp:386-397
    last_ids_[i].swap(ids);
  }
  return out;
}

std::vector<std::string> ParaformerB200::Forward(float** din, int* len, bool input_finished,
                                                 const std::vector<std::vector<float>>& hw_emb, void* wfst_decoder,
                                                 int batch_in) {
  (void)input_finished; (void)hw_emb; (void)wfst_decoder;
  std::vector<std::string> results(batch_in > 0 ? batch_in : 0);
  if (batch_in <= 0 || !engine_) return results;
  std::lock_guard<std::mutex> lock(mu_);
  try {
    // split the caller's batch by the engine's capacity; results keep the caller's order
    int start = 0;
    while (start < batch_in) {
      int64_t rows = 0, samples = 0;
      int end = start;
      while (end < batch_in && end - start < max_segments_) {
        const int T = b200pf_num_lfr_frames(len[end]);
        const int64_t r = T > 0 ? T + 1 : 0;
        if (end > start && rows + r > max_rows_) break;
        rows += r;
        samples += len[end];
        ++end;
      }
      const int n = end - start;
      if (!batch_ || samples > batch_samples_) {
        if (batch_) b200pf_batch_destroy(batch_);
        batch_ = nullptr;
        batch_samples_ = samples + samples / 4 + 16000;
        if (b200pf_batch_create(engine_, batch_samples_, &batch_) != 0) throw std::string(b200pf_last_error());
      }
      std::vector<int32_t> counts(n), offs(n + 1), frames(n), ids((size_t)max_rows_), fire((size_t)max_rows_);
      b200pf_result r;
      r.token_counts = counts.data(); r.token_offsets = offs.data(); r.lfr_frames = frames.data();
      r.token_ids = ids.data(); r.fire_frames = fire.data(); r.cap_tokens = max_rows_; r.n_tokens = 0;
      if (b200pf_forward_f32(batch_, din + start, len + start, n, &r) != 0) throw std::string(b200pf_last_error());
      std::vector<std::string> part = Decode(r, n);
      for (int i = 0; i < n; ++i) results[start + i] = part[i];
      start = end;
    }
  } catch (const std::string& e) {
    // the reference logs and returns "" for the failing call, never throws (paraformer.cpp:582-587)
    fprintf(stderr, "ParaformerB200::Forward: %s\n", e.c_str());
  }
  return results;
}

std::string ParaformerB200::Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb,
                                    void* wfst_decoder) {
  float* one[1] = {din};
  int l[1] = {len};
  std::vector<std::string> r = Forward(one, l, input_finished, hw_emb, wfst_decoder, 1);
  return r.empty() ? std::string() : r[0];
}

std::vector<std::string> ParaformerB200::ForwardPcm16(const int16_t* pcm, const int64_t* offsets, int n_seg) {
  std::vector<std::string> results(n_seg > 0 ? n_seg : 0);
  if (n_seg <= 0 || !engine_) return results;
  std::lock_guard<std::mutex> lock(mu_);
  int start = 0;
  while (start < n_seg) {
    int64_t rows = 0;
    int end = start;
    while (end < n_seg && end - start < max_segments_) {
      const int T = b200pf_num_lfr_frames(offsets[end + 1] - offsets[end]);
      const int64_t r = T > 0 ? T + 1 : 0;
      if (end > start && rows + r > max_rows_) break;
      rows += r;
      ++end;
    }
    const int n = end - start;
    const int64_t samples = offsets[end] - offsets[start];
    if (!batch_ || samples > batch_samples_) {
      if (batch_) b200pf_batch_destroy(batch_);
      batch_ = nullptr;
      batch_samples_ = samples + samples / 4 + 16000;
      if (b200pf_batch_create(engine_, batch_samples_, &batch_) != 0) { fprintf(stderr, "%s\n", b200pf_last_error()); return results; }
    }
    std::vector<int32_t> counts(n), offs(n + 1), frames(n), ids((size_t)max_rows_), fire((size_t)max_rows_);
    b200pf_result r;
    r.token_counts = counts.data(); r.token_offsets = offs.data(); r.lfr_frames = frames.data();
    r.token_ids = ids.data(); r.fire_frames = fire.data(); r.cap_tokens = max_rows_; r.n_tokens = 0;
    if (b200pf_forward_s16(batch_, pcm, offsets + start, n, &r) != 0) { fprintf(stderr, "%s\n", b200pf_last_error()); return results; }
    std::vector<std::string> part = Decode(r, n);
    for (int i = 0; i < n; ++i) results[start + i] = part[i];
    start = end;
  }
  return results;
}

}  // namespace funasr_b200
