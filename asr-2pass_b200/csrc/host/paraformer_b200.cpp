#include "paraformer_b200.h"

#include <chrono>
#include <thread>

#include "logprob_adapter.h"
#ifdef B200PF_WITH_REFERENCE_HEADERS
#include "wfst-decoder.h"   // funasr::WfstDecoder (onnxruntime/src/wfst-decoder.h:59-84)
#endif

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace funasr_b200 {

namespace {
thread_local int tl_failed_segments = 0;
std::string DirOf(const std::string& path) {
  const size_t p = path.find_last_of('/');
  return p == std::string::npos ? std::string(".") : path.substr(0, p);
}
}  // namespace

int ParaformerB200::last_failed_segments() { return tl_failed_segments; }

ParaformerB200::ParaformerB200(int device, int max_rows, int max_segments)
    : device_(device), max_rows_(max_rows), max_segments_(max_segments) {}

ParaformerB200::~ParaformerB200() {
  for (Slot& sl : slots_) if (sl.batch) b200pf_batch_destroy(sl.batch);
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
  d_model_ = cfg.d_model;
  use_hotword_ = cfg.contextual != 0;
  has_timestamp_ = cfg.timestamp != 0;
  for (size_t i = 0; i < vocab_->tokens().size(); ++i) token_id_.emplace(vocab_->tokens()[i], (int)i);  // PhoneSet: first id wins
#ifdef B200PF_WITH_REFERENCE_HEADERS
  // the reference's decoders read its own Vocab / PhoneSet (Paraformer::InitAsr, paraformer.cpp:36-40)
  const std::string token_file = model_dir + "/tokens.json";
  ref_vocab_.reset(new funasr::Vocab(token_file.c_str()));
  ref_phone_set_.reset(new funasr::PhoneSet(token_file.c_str()));
  phone_set_ = ref_phone_set_.get();
  lm_vocab = nullptr;
#endif
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

namespace {

// UTF-8 -> code points (BMP is all the reference's UTF-16 path distinguishes here)
std::vector<std::pair<uint32_t, std::string>> CodePoints(const std::string& s) {
  std::vector<std::pair<uint32_t, std::string>> out;
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = (unsigned char)s[i];
    int n = 1;
    uint32_t cp = c;
    if ((c & 0xe0) == 0xc0) { n = 2; cp = c & 0x1f; }
    else if ((c & 0xf0) == 0xe0) { n = 3; cp = c & 0x0f; }
    else if ((c & 0xf8) == 0xf0) { n = 4; cp = c & 0x07; }
    if (i + n > s.size()) n = (int)(s.size() - i);
    for (int k = 1; k < n; ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3f);
    out.emplace_back(cp, s.substr(i, n));
    i += n;
  }
  return out;
}

std::vector<std::string> SplitChar(const std::string& s, char delim) {  // util.cpp:639-647 (std::getline semantics)
  std::vector<std::string> elems;
  size_t start = 0;
  if (s.empty()) return elems;
  while (true) {
    const size_t p = s.find(delim, start);
    if (p == std::string::npos) {
      if (start < s.size()) elems.push_back(s.substr(start));
      break;
    }
    elems.push_back(s.substr(start, p - start));
    start = p + 1;
  }
  return elems;
}

}  // namespace

void ParaformerB200::InitHwCompiler(const std::string& hw_model, int thread_num) {
  (void)hw_model; (void)thread_num;
  if (!use_hotword_) fprintf(stderr, "InitHwCompiler: model.b200pf carries no hotword compiler (contextual = 0); hotwords are ignored\n");
}

void ParaformerB200::InitSegDict(const std::string& seg_dict_model) { LoadSegDict(seg_dict_model, &seg_dict_); }

// Hotword string -> id matrix [n, 10] + lengths, the host half of Paraformer::CompileHotwordEmbedding (paraformer.cpp:600-648):
// split on ' '; an all-CJK word becomes its characters, anything else goes through the seg_dict; words with no pieces or with a
// piece missing from the token list (PhoneSet::String2Id == -1 among the first 10) are dropped; the blank row {1, 0, ...} of
// length 1 is appended last.
void PackHotwords(const std::string& hotwords, const std::unordered_map<std::string, int>& token_id,
                  const std::unordered_map<std::string, std::vector<std::string>>& seg_dict, std::vector<int32_t>* matrix,
                  std::vector<int32_t>* lengths) {
  const int max_len = B200PF_HOTWORD_LEN;
  matrix->clear();
  lengths->clear();
  if (!hotwords.empty()) {
    for (const std::string& hotword : SplitChar(hotwords, ' ')) {
      std::vector<std::string> chars;
      const auto cps = CodePoints(hotword);
      bool all_cjk = !hotword.empty();
      for (const auto& cp : cps) all_cjk = all_cjk && cp.first >= 0x4e00 && cp.first <= 0x9fff;   // IsAllChineseCharactor
      if (all_cjk) {
        for (const auto& cp : cps)  // KeepChineseCharacterAndSplit (util.cpp:192-209)
          if ((cp.first >= 0x4e00 && cp.first <= 0x9fff) || (cp.first >= 0x3400 && cp.first <= 0x4dff)) chars.push_back(cp.second);
      } else {
        for (const std::string& word : SplitChar(hotword, ' ')) {
          auto it = seg_dict.find(word);  // SegDict::GetTokensByWord: OOV -> no tokens
          if (it != seg_dict.end()) chars.insert(chars.end(), it->second.begin(), it->second.end());
        }
      }
      if (chars.empty()) continue;
      std::vector<int32_t> row(max_len, 0);
      const int len = std::min(max_len, (int)chars.size());
      bool oov = false;
      for (int i = 0; i < len && !oov; ++i) {
        auto it = token_id.find(chars[i]);
        if (it == token_id.end()) oov = true; else row[i] = it->second;
      }
      if (oov) continue;
      lengths->push_back(len);
      matrix->insert(matrix->end(), row.begin(), row.end());
    }
  }
  std::vector<int32_t> blank(max_len, 0);
  blank[0] = 1;
  matrix->insert(matrix->end(), blank.begin(), blank.end());
  lengths->push_back(1);
}

void LoadSegDict(const std::string& path, std::unordered_map<std::string, std::vector<std::string>>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "%s open failed !!\n", path.c_str()); return; }
  std::string text;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  for (const std::string& line : SplitChar(text, '\n')) {
    const std::vector<std::string> item = SplitChar(line, '\t');
    if (item.size() > 1) (*out)[item[0]] = SplitChar(item[1], ' ');
  }
}

std::vector<std::vector<float>> ParaformerB200::CompileHotwordEmbedding(std::string& hotwords) {
  const int dim = d_model_;
  if (!use_hotword_) return std::vector<std::vector<float>>(1, std::vector<float>(dim, 0.0f));  // paraformer.cpp:595-599
  const int max_len = B200PF_HOTWORD_LEN;
  std::vector<int32_t> matrix, lengths;
  PackHotwords(hotwords, token_id_, seg_dict_, &matrix, &lengths);
  const int n = (int)lengths.size();
  std::vector<float> flat((size_t)n * dim);
  std::vector<std::vector<float>> result;
  {
    std::lock_guard<std::mutex> lock(mu_);
    if (b200pf_engine_hotword_embed(engine_, matrix.data(), lengths.data(), n, max_len, flat.data()) != 0) {
      fprintf(stderr, "CompileHotwordEmbedding: %s\n", b200pf_last_error());  // the reference logs and returns {} too
      return result;
    }
  }
  for (int j = 0; j < n; ++j) result.emplace_back(flat.begin() + (size_t)j * dim, flat.begin() + (size_t)(j + 1) * dim);
  return result;
}

bool ParaformerB200::UseLmDecoder(void* wfst_decoder) const {
#ifdef B200PF_WITH_REFERENCE_HEADERS
  return wfst_decoder != nullptr && lm_ != nullptr;
#else
  (void)wfst_decoder;
  return false;   // the stand-alone shim has no LM decoder: FunASRWfstDecoderInit returns nullptr there
#endif
}

#ifdef B200PF_WITH_REFERENCE_HEADERS
void ParaformerB200::InitLm(const std::string& lm_file, const std::string& lm_cfg_file, const std::string& lex_file,
                            const std::string& lm_units_file) {
  lm_ = std::shared_ptr<fst::Fst<fst::StdArc>>(fst::Fst<fst::StdArc>::Read(lm_file));
  if (!lm_) { fprintf(stderr, "Failed to load lm file %s\n", lm_file.c_str()); return; }
  lm_vocab = new funasr::Vocab(lm_cfg_file.c_str(), lex_file.c_str());
  if (!lm_units_file.empty()) phone_set_ = new funasr::PhoneSet(lm_units_file.c_str());
}
#endif

std::string ParaformerB200::TextOf(const SegmentRaw& seg) {
  if (!seg.has_features) return std::string();
  if (!has_timestamp_) return vocab_->ToText(seg.ids, language_);  // GreedySearch, paraformer.cpp:386-397
  // GreedySearch with is_stamp (paraformer.cpp:398-407): Vector2String -> TimestampOnnx -> PostProcess
  std::vector<std::string> pieces = vocab_->ToPieces(seg.ids);
  std::vector<std::string> raw(pieces);
  std::vector<float> us_alphas(seg.us_alphas), us_peaks(seg.us_peaks);
  std::string dbg;
  std::vector<pf::host::Span> spans = pf::host::TimestampFromPeaks(&us_alphas, us_peaks, &pieces, &dbg);
  return pf::host::MergeWithStamps(raw, spans);
}

std::vector<std::string> ParaformerB200::Decode(const b200pf_result& r, int n_seg, void* wfst_decoder, SegmentRaw* raw_out) {
  std::vector<std::string> out(n_seg);
  last_ids_.assign(n_seg, std::vector<int>());
  // Large greedy batches without stamps: host threads build each segment's text for both incoming detokeniser states, then one
  // serial pass follows the state chain -- the same strings as the loop below, without 2-3 ms of serial string work behind the
  // last (largest) sub-batch of a call.
  if (!raw_out && !has_timestamp_ && !UseLmDecoder(wfst_decoder) && n_seg >= 64) {
    struct Both { std::string text[2]; bool ended[2]; bool used = false; };
    std::vector<Both> both(n_seg);
    const int nth = (int)std::min<unsigned>(4u, std::max(1u, std::thread::hardware_concurrency() / 4));
    auto work = [&](int t) {
      for (int i = t; i < n_seg; i += nth) {
        if (r.lfr_frames[i] <= 0) continue;
        std::vector<int> ids(r.token_ids + r.token_offsets[i], r.token_ids + r.token_offsets[i] + r.token_counts[i]);
        for (int v = 0; v < 2; ++v) both[i].text[v] = vocab_->ToText(ids, language_, v != 0, &both[i].ended[v]);
        both[i].used = true;
        last_ids_[i].swap(ids);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nth; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    bool st = vocab_->ended_on_english_word();
    for (int i = 0; i < n_seg; ++i) {
      if (!both[i].used) continue;
      out[i] = std::move(both[i].text[st ? 1 : 0]);
      st = both[i].ended[st ? 1 : 0];
    }
    vocab_->set_ended_on_english_word(st);
    return out;
  }
  for (int i = 0; i < n_seg; ++i) {
    const int cnt = r.token_counts[i];
    if (raw_out) raw_out[i] = SegmentRaw();
    if (r.lfr_frames[i] <= 0) continue;  // empty features -> "" (paraformer.cpp:477-480)
    std::vector<int> ids(r.token_ids + r.token_offsets[i], r.token_ids + r.token_offsets[i] + cnt);
#ifdef B200PF_WITH_REFERENCE_HEADERS
    if (UseLmDecoder(wfst_decoder) && r.topk_k > 0) {
      // BeamSearch + FinalizeDecode (paraformer.cpp:565-578) on log-softmax rows rebuilt from the pruned posteriors
      std::vector<float> dense;
      const int V = b200pf_engine_vocab_size(engine_);
      pf::host::ExpandPrunedPosteriors(r.topk_logprob + (size_t)r.token_offsets[i] * r.topk_k, r.topk_ids + (size_t)r.token_offsets[i] * r.topk_k,
                                       cnt, r.topk_k, V, &dense);
      funasr::WfstDecoder* dec = (funasr::WfstDecoder*)wfst_decoder;
      out[i] = dec->Search(dense.data(), cnt, V);
      if (has_timestamp_) {
        std::vector<float> us_alphas(r.us_alphas + r.us_offsets[i], r.us_alphas + r.us_offsets[i + 1]);
        std::vector<float> us_peaks(r.us_peaks + r.us_offsets[i], r.us_peaks + r.us_offsets[i + 1]);
        out[i] = dec->FinalizeDecode(true, us_alphas, us_peaks);
      } else {
        out[i] = dec->FinalizeDecode();
      }
      last_ids_[i].swap(ids);
      continue;
    }
#endif
    {
      SegmentRaw seg;
      seg.has_features = true;
      seg.ids = ids;
      if (has_timestamp_) {
        seg.us_alphas.assign(r.us_alphas + r.us_offsets[i], r.us_alphas + r.us_offsets[i + 1]);
        seg.us_peaks.assign(r.us_peaks + r.us_offsets[i], r.us_peaks + r.us_offsets[i + 1]);
      }
      if (raw_out) {   // the caller assembles the text (MultiGpuParaformer: in the caller's order); both variants are prepared here
        if (has_timestamp_) {
          seg.text[0] = seg.text[1] = TextOf(seg);   // the stamp path does not go through the stateful detokeniser
        } else {
          for (int v = 0; v < 2; ++v) seg.text[v] = vocab_->ToText(seg.ids, language_, v != 0, &seg.ended[v]);
        }
        seg.has_text = true;
        raw_out[i] = std::move(seg);
      } else {
        out[i] = TextOf(seg);
      }
    }
    last_ids_[i].swap(ids);
  }
  return out;
}

// One engine-sized sub-batch is staged into slot `k` (its own b200pf_batch: device PCM + layout) on the engine's COPY
// stream, so that the host-to-device copies of sub-batch i+1 run while sub-batch i computes.
bool ParaformerB200::StageSlot(int k, const int16_t* pcm, const int64_t* offsets, float** din, int* len, int n, int64_t samples,
                               const std::vector<std::vector<float>>& hw_emb, const int16_t* const* seg16, const int64_t* len16) {
  Slot& sl = slots_[k];
  if (!sl.batch || samples > sl.samples) {
    if (sl.batch) b200pf_batch_destroy(sl.batch);
    sl.batch = nullptr;
    sl.hw_valid = false;
    // Sized ONCE for anything the engine can hold (a packed row is 6 fbank shifts = 960 samples; every segment adds at most
    // one partial row and the 240-sample window overlap): re-creating a batch object costs tens of milliseconds of
    // cudaMalloc / cudaMallocHost, which showed up as 100 ms stalls when batch sizes grew from call to call.
    const int64_t full = (int64_t)max_rows_ * 960 + (int64_t)max_segments_ * 1200;
    sl.samples = samples > full ? samples + 16000 : full;
    if (b200pf_batch_create(engine_, sl.samples, &sl.batch) != 0) { fprintf(stderr, "ParaformerB200: %s\n", b200pf_last_error()); return false; }
  }
  if (use_hotword_) {
    if (hw_emb.empty()) { fprintf(stderr, "hw_emb is null\n"); return false; }  // paraformer.cpp:516-520
    // flatten [n_hw][dim] like the reference (paraformer.cpp:521-526); re-upload only when the matrix changed
    std::vector<float> flat;
    flat.reserve(hw_emb.size() * hw_emb[0].size());
    for (const auto& row : hw_emb) flat.insert(flat.end(), row.begin(), row.end());
    if (!sl.hw_valid || flat != sl.hw_flat) {
      if (flat.size() != hw_emb.size() * (size_t)d_model_ ||
          b200pf_batch_set_hotwords(sl.batch, flat.data(), (int)hw_emb.size(), (int)hw_emb[0].size()) != 0) {
        fprintf(stderr, "ParaformerB200: bad hotword embedding (%zu x %zu): %s\n", hw_emb.size(), hw_emb[0].size(), b200pf_last_error());
        return false;  // ORT would throw on the shape mismatch -> "" (paraformer.cpp:582-587)
      }
      sl.hw_flat.swap(flat);
      sl.hw_valid = true;
    }
  }
  void* cs = b200pf_engine_copy_stream(engine_);
  const int rc = seg16 ? b200pf_batch_stage_s16_ptrs(sl.batch, seg16, len16, n, cs)
                       : (pcm ? b200pf_batch_stage_s16(sl.batch, pcm, offsets, n, cs) : b200pf_batch_stage_f32(sl.batch, din, len, n, cs));
  if (rc != 0) { fprintf(stderr, "ParaformerB200::Forward: %s\n", b200pf_last_error()); return false; }
  return true;
}

bool ParaformerB200::CollectSlot(int k, int n, std::vector<std::string>* out, void* wfst_decoder, SegmentRaw* raw) {
  std::vector<int32_t> counts(n), offs(n + 1), frames(n), ids((size_t)max_rows_), fire((size_t)max_rows_), us_offs(n + 1), tk_ids;
  std::vector<float> us_a, us_p, tk_lp, tk_lse;
  b200pf_result r;
  memset(&r, 0, sizeof(r));
  r.token_counts = counts.data(); r.token_offsets = offs.data(); r.lfr_frames = frames.data();
  r.token_ids = ids.data(); r.fire_frames = fire.data(); r.cap_tokens = max_rows_;
  r.us_offsets = us_offs.data();
  if (has_timestamp_) {
    us_a.resize((size_t)3 * max_rows_); us_p.resize((size_t)3 * max_rows_);
    r.us_alphas = us_a.data(); r.us_peaks = us_p.data(); r.cap_us = (int64_t)3 * max_rows_;
  }
  if (UseLmDecoder(wfst_decoder)) {   // the LM decoder reads pruned posteriors instead of greedy ids
    tk_lp.resize((size_t)max_rows_ * B200PF_MAX_TOPK); tk_ids.resize((size_t)max_rows_ * B200PF_MAX_TOPK); tk_lse.resize((size_t)max_rows_);
    r.topk_logprob = tk_lp.data(); r.topk_ids = tk_ids.data(); r.token_lse = tk_lse.data();
  }
  if (b200pf_batch_collect(slots_[k].batch, &r, nullptr) != 0) { fprintf(stderr, "ParaformerB200::Forward: %s\n", b200pf_last_error()); return false; }
  *out = Decode(r, n, wfst_decoder, raw);
  return true;
}

// Shared driver of both Forward flavours: split the caller's batch by the engine's capacity, then run the sub-batches
// through two slots: stage(i+1) overlaps compute(i).  A failing sub-batch logs and leaves "" for its items, never throws
// (paraformer.cpp:582-587).  Results keep the caller's order.
std::vector<std::string> ParaformerB200::RunAll(const int16_t* pcm, const int64_t* offsets, float** din, int* len, int n_seg,
                                                const std::vector<std::vector<float>>& hw_emb, const int16_t* const* seg16,
                                                const int64_t* len16, void* wfst_decoder, SegmentRaw* raw) {
  std::vector<std::string> results(n_seg > 0 ? n_seg : 0);
  tl_failed_segments = 0;
  if (n_seg <= 0 || !engine_) return results;
  std::lock_guard<std::mutex> lock(mu_);
  {  // pruned posteriors are only computed for calls that decode with the LM
    const int want = UseLmDecoder(wfst_decoder) ? B200PF_MAX_TOPK : 0;
    if (want != topk_on_) { b200pf_engine_set_option(engine_, "logprob_topk", want); topk_on_ = want; }
  }
  struct Sub { int start, end; int64_t samples; };
  std::vector<Sub> subs;
  // Host buffers are usually pageable: the driver copies them through its own staging buffer at a fraction of the PCIe rate and
  // the calling thread waits for it.  A call whose segments would fit ONE engine-sized batch would expose that whole copy (708 MB
  // of float samples for the configs[1] workload).  Larger calls are therefore cut into sub-batches that GROW geometrically
  // from 8192 rows (x 4 each, up to the engine's capacity): only the first, small copy is exposed, every later copy runs while
  // the GPU computes the sub-batch before it, and most rows still travel in large batches.  B200PF_SUB_ROWS fixes the cap instead.
  static const int64_t env_rows = getenv("B200PF_SUB_ROWS") ? atoll(getenv("B200PF_SUB_ROWS")) : 0;
  int64_t total_rows = 0;
  for (int i = 0; i < n_seg; ++i) {
    const int64_t ns = seg16 ? len16[i] : (pcm ? offsets[i + 1] - offsets[i] : (int64_t)len[i]);
    const int T = b200pf_num_lfr_frames(ns);
    total_rows += T > 0 ? T + 1 : 0;
  }
  const bool grow = env_rows <= 0 && total_rows > 16384;
  int64_t row_cap = grow ? std::min<int64_t>(max_rows_, 8192) : (env_rows > 0 ? std::min<int64_t>(max_rows_, env_rows) : (int64_t)max_rows_);
  int start = 0;
  int64_t rows_done = 0;
  while (start < n_seg) {
    int64_t rows = 0, samples = 0;
    int end = start;
    auto seg_rows = [&](int k, int64_t* ns_out) {
      const int64_t ns = seg16 ? len16[k] : (pcm ? offsets[k + 1] - offsets[k] : (int64_t)len[k]);
      const int T = b200pf_num_lfr_frames(ns);
      *ns_out = ns;
      return (int64_t)(T > 0 ? T + 1 : 0);
    };
    while (end < n_seg && end - start < max_segments_) {
      int64_t ns;
      const int64_t r = seg_rows(end, &ns);
      if (end > start && rows + r > row_cap) break;
      rows += r;
      samples += ns;
      ++end;
    }
    // A small remainder (less than a quarter of this sub-batch) would be one more launch-bound forward: a forward of a few thousand
    // rows costs ~8 ms whatever its size.  If it still fits the engine, this sub-batch takes it along.
    const int64_t rest = total_rows - rows_done - rows;
    if (grow && end < n_seg && rest * 4 <= rows && rows + rest <= max_rows_ && n_seg - start <= max_segments_) {
      for (; end < n_seg; ++end) {
        int64_t ns;
        rows += seg_rows(end, &ns);
        samples += ns;
      }
    }
    rows_done += rows;
    subs.push_back(Sub{start, end, samples});
    start = end;
    if (grow) row_cap = std::min<int64_t>(max_rows_, row_cap * 4);
  }
  auto stage = [&](size_t i) {
    const Sub& sb = subs[i];
    if (seg16) return StageSlot((int)(i & 1), nullptr, nullptr, nullptr, nullptr, sb.end - sb.start, sb.samples, hw_emb, seg16 + sb.start, len16 + sb.start);
    return StageSlot((int)(i & 1), pcm, pcm ? offsets + sb.start : nullptr, pcm ? nullptr : din + sb.start, pcm ? nullptr : len + sb.start,
                     sb.end - sb.start, sb.samples, hw_emb);
  };
  // Software pipeline over two slots: sub-batch i+1 is staged AND enqueued before sub-batch i is collected, so the GPU goes
  // from one forward straight into the next while the host reads i's results and builds its strings (a batch's results live in
  // its own buffers; collect waits for that batch's completion event only).
  std::vector<char> ok(subs.size(), 0), running(subs.size(), 0);
  auto launch = [&](size_t i) {
    ok[i] = stage(i);
    if (ok[i]) {
      running[i] = b200pf_batch_run(slots_[i & 1].batch, nullptr) == 0;   // asynchronous; waits for the slot's staged event
      if (!running[i]) fprintf(stderr, "ParaformerB200::Forward: %s\n", b200pf_last_error());
    }
  };
  static const bool trace = getenv("B200PF_HOST_TRACE") != nullptr;   // per-sub-batch wall times on stderr
  using clk = std::chrono::steady_clock;
  const clk::time_point t0 = clk::now();
  auto ms_since = [&](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
  launch(0);
  if (trace) fprintf(stderr, "[b200pf dev %d] %d segments in %zu sub-batches; first staged + enqueued at %.2f ms\n", device_, n_seg, subs.size(), ms_since(t0));
  for (size_t i = 0; i < subs.size(); ++i) {
    if (i + 1 < subs.size()) launch(i + 1);          // slot (i+1)&1 was collected one iteration ago
    if (trace && i + 1 < subs.size()) fprintf(stderr, "[b200pf dev %d]   sub-batch %zu (%d segments) staged + enqueued at %.2f ms\n", device_, i + 1, subs[i + 1].end - subs[i + 1].start, ms_since(t0));
    bool done = false;
    if (running[i]) {
      std::vector<std::string> part;
      if (CollectSlot((int)(i & 1), subs[i].end - subs[i].start, &part, wfst_decoder, raw ? raw + subs[i].start : nullptr)) {
        for (int k = 0; k < subs[i].end - subs[i].start; ++k) results[subs[i].start + k] = part[k];
        done = true;
      }
    }
    if (!done) tl_failed_segments += subs[i].end - subs[i].start;
    if (trace) fprintf(stderr, "[b200pf dev %d]   sub-batch %zu (%d segments) collected + decoded at %.2f ms\n", device_, i, subs[i].end - subs[i].start, ms_since(t0));
  }
  return results;
}

std::vector<std::string> ParaformerB200::Forward(float** din, int* len, bool input_finished,
                                                 const std::vector<std::vector<float>>& hw_emb, void* wfst_decoder,
                                                 int batch_in) {
  (void)input_finished;
  return RunAll(nullptr, nullptr, din, len, batch_in, hw_emb, nullptr, nullptr, wfst_decoder);
}

std::string ParaformerB200::Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb,
                                    void* wfst_decoder) {
  float* one[1] = {din};
  int l[1] = {len};
  std::vector<std::string> r = Forward(one, l, input_finished, hw_emb, wfst_decoder, 1);
  return r.empty() ? std::string() : r[0];
}

std::vector<std::string> ParaformerB200::ForwardSegments16(const int16_t* const* seg, const int64_t* len, int n_seg,
                                                           const std::vector<std::vector<float>>& hw_emb) {
  return RunAll(nullptr, nullptr, nullptr, nullptr, n_seg, hw_emb, seg, len);
}

std::vector<SegmentRaw> ParaformerB200::ForwardRaw(float** din, int* len, int batch_in, const std::vector<std::vector<float>>& hw_emb) {
  std::vector<SegmentRaw> raw(batch_in > 0 ? batch_in : 0);
  RunAll(nullptr, nullptr, din, len, batch_in, hw_emb, nullptr, nullptr, nullptr, raw.data());
  return raw;
}

std::vector<SegmentRaw> ParaformerB200::ForwardSegments16Raw(const int16_t* const* seg, const int64_t* len, int n_seg,
                                                             const std::vector<std::vector<float>>& hw_emb) {
  std::vector<SegmentRaw> raw(n_seg > 0 ? n_seg : 0);
  RunAll(nullptr, nullptr, nullptr, nullptr, n_seg, hw_emb, seg, len, nullptr, raw.data());
  return raw;
}

std::vector<std::string> ParaformerB200::ForwardPcm16(const int16_t* pcm, const int64_t* offsets, int n_seg,
                                                      const std::vector<std::vector<float>>& hw_emb) {
  return RunAll(pcm, offsets, nullptr, nullptr, n_seg, hw_emb);
}

}  // namespace funasr_b200
