#include "punc_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>

namespace funasr_b200 {

namespace {
constexpr int kMiniSentence = 20;     // TOKEN_LEN
constexpr int kCachePopLimit = 200;   // CACHE_POP_TRIGGER_LIMIT
constexpr int kNotPunc = 1, kComma = 2, kPeriod = 3, kQuestion = 4, kDun = 5;

bool HighBit(const std::string& s) { return !s.empty() && (static_cast<unsigned char>(s[0]) & 0x80) != 0; }

// one UTF-8 sequence per element; the length is the run of leading one-bits of the first byte (tokenizer.cpp:268-284)
void SplitUtf8(const std::string& s, std::vector<std::string>* out) {
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = static_cast<unsigned char>(s[i]);
    size_t len = 1;
    for (int j = 0; j < 6 && (c & (0x80 >> j)); ++j) len = (size_t)j + 1;
    out->push_back(s.substr(i, len));
    i += len;
  }
}

// tokens.json / punc_list.json: a JSON array of strings
bool ReadStringArray(const std::string& path, std::vector<std::string>* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string s = ss.str();
  size_t i = s.find('[');
  if (i == std::string::npos) return false;
  ++i;
  while (i < s.size()) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t' || s[i] == ',')) ++i;
    if (i >= s.size()) return false;
    if (s[i] == ']') return true;
    if (s[i] != '"') return false;
    ++i;
    std::string cur;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        const char e = s[i + 1];
        if (e == 'u' && i + 5 < s.size()) {
          unsigned cp = (unsigned)strtoul(s.substr(i + 2, 4).c_str(), nullptr, 16);
          i += 6;
          if (cp >= 0xD800 && cp <= 0xDBFF && i + 5 < s.size() && s[i] == '\\' && s[i + 1] == 'u') {
            const unsigned lo = (unsigned)strtoul(s.substr(i + 2, 4).c_str(), nullptr, 16);
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            i += 6;
          }
          if (cp < 0x80) cur += (char)cp;
          else if (cp < 0x800) { cur += (char)(0xC0 | (cp >> 6)); cur += (char)(0x80 | (cp & 0x3F)); }
          else if (cp < 0x10000) { cur += (char)(0xE0 | (cp >> 12)); cur += (char)(0x80 | ((cp >> 6) & 0x3F)); cur += (char)(0x80 | (cp & 0x3F)); }
          else { cur += (char)(0xF0 | (cp >> 18)); cur += (char)(0x80 | ((cp >> 12) & 0x3F)); cur += (char)(0x80 | ((cp >> 6) & 0x3F)); cur += (char)(0x80 | (cp & 0x3F)); }
          continue;
        }
        cur += e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e == 'b' ? '\b' : e == 'f' ? '\f' : e;
        i += 2;
        continue;
      }
      cur += s[i++];
    }
    if (i >= s.size()) return false;
    ++i;
    out->push_back(cur);
  }
  return false;
}

std::string DirOf(const std::string& path) {
  const size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string(".") : path.substr(0, k);
}
}  // namespace

void PuncTokenizer::Open(const std::vector<std::string>& tokens, const std::vector<std::string>& punc_list) {
  token2id_.clear();
  for (size_t i = 0; i < tokens.size(); ++i) token2id_[tokens[i]] = (int)i;   // a repeated token keeps its LAST id (tokenizer.cpp:168-172)
  n_tokens_ = (int)tokens.size();
  auto it = token2id_.find("<unk>");
  unk_ = it == token2id_.end() ? 0 : it->second;   // operator[] on a missing key yields 0 in the reference
  punc_ = punc_list;
}

void PuncTokenizer::Tokenize(const char* text, std::vector<std::string>* pieces, std::vector<int32_t>* ids) const {
  pieces->clear();
  ids->clear();
  const std::string s(text ? text : "");
  // words are separated by ' '; inside a word, runs of bytes with the high bit set are split into UTF-8 characters and runs of
  // ASCII bytes stay whole (tokenizer.cpp:312-365)
  size_t begin = 0;
  while (begin <= s.size() && !s.empty()) {
    size_t end = s.find(' ', begin);
    if (end == std::string::npos) end = s.size();
    std::string ascii, wide;
    for (size_t i = begin; i < end; ++i) {
      const char ch = s[i];
      if (!(static_cast<unsigned char>(ch) & 0x80)) {
        if (!wide.empty()) { SplitUtf8(wide, pieces); wide.clear(); }
        ascii += ch;
      } else {
        if (!ascii.empty()) { pieces->push_back(ascii); ascii.clear(); }
        wide += ch;
      }
    }
    if (!wide.empty()) SplitUtf8(wide, pieces);
    if (!ascii.empty()) pieces->push_back(ascii);
    begin = end + 1;
  }
  ids->reserve(pieces->size());
  for (const std::string& p : *pieces) {
    std::string low(p);
    for (char& c : low)
      if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    auto it = token2id_.find(low);
    ids->push_back(it == token2id_.end() ? unk_ : it->second);
  }
}

PuncJob::PuncJob(const PuncTokenizer* tok, const char* text, const std::string& language) : tok_(tok), language_(language) {
  tok_->Tokenize(text, &pieces_, &ids_);
  n_total_ = (int)std::ceil((float)ids_.size() / kMiniSentence);
  Prepare();
}

void PuncJob::Prepare() {
  in_ids_.clear();
  in_str_.clear();
  if (!Active()) return;
  const size_t end = std::min(ids_.size(), pos_ + kMiniSentence);
  in_ids_ = remain_ids_;
  in_ids_.insert(in_ids_.end(), ids_.begin() + pos_, ids_.begin() + end);
  in_str_ = remain_str_;
  in_str_.insert(in_str_.end(), pieces_.begin() + pos_, pieces_.begin() + end);
}

void PuncJob::Consume(const int32_t* punc_in, int n) {
  std::vector<int32_t> punc(punc_in, punc_in + n);
  const int cur = (int)(pos_ / kMiniSentence);
  const bool last = cur == n_total_ - 1;
  if (!last) {
    // cut after the last sentence end (。 or ？, compared by their strings) strictly inside the window; what follows is carried over
    int sent_end = -1, last_comma = -1;
    for (int k = n - 2; k > 0; --k) {
      const std::string& p = tok_->Id2Punc(punc[k]);
      if (p == tok_->Id2Punc(kPeriod) || p == tok_->Id2Punc(kQuestion)) { sent_end = k; break; }
      if (last_comma < 0 && p == tok_->Id2Punc(kComma)) last_comma = k;
    }
    if (sent_end < 0 && (int)in_str_.size() > kCachePopLimit && last_comma > 0) {
      sent_end = last_comma;      // no sentence end within 200 cached tokens: the last comma becomes a period
      punc[sent_end] = kPeriod;
    }
    remain_str_.assign(in_str_.begin() + (sent_end + 1), in_str_.end());
    remain_ids_.assign(in_ids_.begin() + (sent_end + 1), in_ids_.end());
    in_str_.resize((size_t)(sent_end + 1));
    punc.resize((size_t)(sent_end + 1));
  }
  for (size_t k = 0; k < in_str_.size(); ++k) {
    // two neighbouring ASCII pieces are separated by a space; the space is written INTO the piece, so a third one sees ' ' (ASCII) too
    if (k > 0 && !HighBit(in_str_[k - 1]) && !HighBit(in_str_[k])) in_str_[k] = " " + in_str_[k];
    new_str_.push_back(in_str_[k]);
    if (punc[k] != kNotPunc) new_str_.push_back(tok_->Id2Punc(punc[k]));
  }
  // the reference copies the whole output after every mini-sentence (NewSentenceOut = NewString, quadratic in the text length); only
  // the copy made after the LAST one is ever read, and only that one gets the tail fix-up
  if (last) {
    sent_out_ = new_str_;
    if (!new_str_.empty()) {
      const std::string& tail = new_str_.back();
      if (tail == tok_->Id2Punc(kComma) || tail == tok_->Id2Punc(kDun)) {
        sent_out_.back() = tok_->Id2Punc(kPeriod);
      } else if (tail != tok_->Id2Punc(kPeriod) && tail != tok_->Id2Punc(kQuestion)) {
        sent_out_.push_back(tok_->Id2Punc(kPeriod));
      }
    }
  }
  pos_ += kMiniSentence;
  Prepare();
}

std::string PuncJob::Result() const {
  std::string res;
  for (const std::string& s : sent_out_) res += s;
  if (language_ == "en-bpe") {
    static const char* zh[4] = {"\xEF\xBC\x8C", "\xE3\x80\x82", "\xE3\x80\x81", "\xEF\xBC\x9F"};   // ， 。 、 ？
    static const char en[4] = {',', '.', ',', '?'};
    for (int i = 0; i < 4; ++i) {
      size_t p = 0;
      while ((p = res.find(zh[i], p)) != std::string::npos) {
        res.replace(p, 3, 1, en[i]);
        ++p;
      }
    }
  }
  return res;
}

std::string AddPuncWith(const PuncTokenizer& tok, const char* text, const std::string& language,
                        const std::function<std::vector<int32_t>(const std::vector<int32_t>&)>& infer) {
  PuncJob job(&tok, text, language);
  while (job.Active()) {
    const std::vector<int32_t> punc = infer(job.Input());
    if (punc.size() != job.Input().size()) return "";
    job.Consume(punc.data(), (int)punc.size());
  }
  return job.Result();
}

struct CTTransformerB200::Ticket {
  Ticket(const PuncTokenizer* tok, const char* text, const std::string& language) : job(tok, text, language) {}
  PuncJob job;
  bool done = false, failed = false;
};

CTTransformerB200::~CTTransformerB200() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  if (worker_.joinable()) worker_.join();   // finishes the jobs it still holds
  if (engine_) b200pf_punc_destroy(engine_);
}

long long CTTransformerB200::rounds() const {
  std::lock_guard<std::mutex> lk(mu_);
  return rounds_;
}

bool CTTransformerB200::Init(const std::string& punc_dir, std::string* err) {
  std::vector<std::string> tokens, punc;
  if (!ReadStringArray(punc_dir + "/tokens.json", &tokens) || tokens.empty()) { if (err) *err = punc_dir + "/tokens.json: not a JSON array of strings"; return false; }
  if (!ReadStringArray(punc_dir + "/punc_list.json", &punc) || punc.empty()) { if (err) *err = punc_dir + "/punc_list.json: not a JSON array of strings"; return false; }
  if (b200pf_punc_create(punc_dir.c_str(), device_, max_tokens_, &engine_) != 0) { if (err) *err = b200pf_last_error(); return false; }
  int vocab = 0, n_punc = 0;
  b200pf_punc_info(engine_, &vocab, &n_punc, nullptr, &max_tokens_);
  if (vocab != (int)tokens.size() || n_punc != (int)punc.size() || n_punc < 6) {
    if (err) *err = "punctuation model / tokens.json / punc_list.json sizes disagree";
    b200pf_punc_destroy(engine_);
    engine_ = nullptr;
    return false;
  }
  tok_.Open(tokens, punc);
  worker_ = std::thread(&CTTransformerB200::Run, this);
  return true;
}

void CTTransformerB200::InitPunc(const std::string& punc_model, const std::string& punc_config, const std::string& token_file, int thread_num) {
  (void)punc_config; (void)token_file; (void)thread_num;
  std::string err;
  if (!Init(DirOf(punc_model), &err)) {
    fprintf(stderr, "Error when load punc model: %s\n", err.c_str());   // the reference exits here too (ct-transformer.cpp:24-27)
    exit(-1);
  }
}

// One engine call: the current mini-sentence of as many active jobs as fit, in arrival order; jobs that do not fit wait a round.
void CTTransformerB200::RunRound(std::vector<Ticket*>* active) {
  std::vector<int32_t> ids, offs(1, 0), punc;
  std::vector<Ticket*> who;
  for (Ticket* t : *active) {
    const std::vector<int32_t>& in = t->job.Input();
    if ((int)in.size() > max_tokens_) {
      fprintf(stderr, "CTTransformerB200: a sentence cache outgrew the engine (%zu tokens)\n", in.size());
      t->failed = true;
      continue;
    }
    if ((int)(ids.size() + in.size()) > max_tokens_ || who.size() >= 16384) break;
    ids.insert(ids.end(), in.begin(), in.end());
    offs.push_back((int32_t)ids.size());
    who.push_back(t);
  }
  if (who.empty()) return;
  punc.assign(ids.size(), 0);
  if (b200pf_punc_infer(engine_, ids.data(), offs.data(), (int)who.size(), punc.data(), nullptr) != 0) {
    fprintf(stderr, "Error when run punc forword: %s\n", b200pf_last_error());   // the reference logs and answers "" (ct-transformer.cpp:197-201)
    for (Ticket* t : who) t->failed = true;
    return;
  }
  for (size_t k = 0; k < who.size(); ++k) who[k]->job.Consume(punc.data() + offs[k], offs[k + 1] - offs[k]);
}

void CTTransformerB200::Run() {
  std::vector<Ticket*> active;
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_work_.wait(lk, [&] { return stop_ || !pending_.empty() || !active.empty(); });
    if (pending_.empty() && active.empty()) return;   // stop_ and nothing left
    while (!pending_.empty()) { active.push_back(pending_.front()); pending_.pop_front(); }
    lk.unlock();
    RunRound(&active);
    lk.lock();
    ++rounds_;
    bool any_done = false;
    for (size_t i = 0; i < active.size();) {
      Ticket* t = active[i];
      if (t->failed || !t->job.Active()) {
        t->done = true;
        any_done = true;
        active.erase(active.begin() + i);
      } else {
        ++i;
      }
    }
    if (any_done) cv_done_.notify_all();
  }
}

std::vector<std::string> CTTransformerB200::AddPuncBatch(const std::vector<std::string>& texts, const std::string& language, int* rounds) {
  std::vector<std::string> out(texts.size());
  if (rounds) *rounds = 0;
  if (!engine_) { fprintf(stderr, "CTTransformerB200: not initialised\n"); return out; }
  std::vector<std::unique_ptr<Ticket>> tickets;
  tickets.reserve(texts.size());
  for (const std::string& t : texts) tickets.emplace_back(new Ticket(&tok_, t.c_str(), language));   // tokenised on the caller's thread
  long long r0;
  {
    std::unique_lock<std::mutex> lk(mu_);
    r0 = rounds_;
    for (auto& t : tickets) {
      if (t->job.Active()) pending_.push_back(t.get()); else t->done = true;   // empty text: nothing to run
    }
    cv_work_.notify_all();
    cv_done_.wait(lk, [&] {
      for (auto& t : tickets)
        if (!t->done) return false;
      return true;
    });
    if (rounds) *rounds = (int)(rounds_ - r0);
  }
  for (size_t j = 0; j < tickets.size(); ++j) out[j] = tickets[j]->failed ? std::string() : tickets[j]->job.Result();
  return out;
}

std::string CTTransformerB200::AddPunc(const char* sz_input, std::string language) {
  std::vector<std::string> r = AddPuncBatch(std::vector<std::string>(1, sz_input ? sz_input : ""), language, nullptr);
  return r.empty() ? std::string() : r[0];
}

// ---- realtime model -------------------------------------------------------------------------------------------------------
std::string AddPuncOnlineWith(const PuncTokenizer& tok, const char* text, std::vector<std::string>* cache,
                              const std::function<std::vector<int32_t>(const std::vector<int32_t>&, int)>& infer) {
  // full text = cached words + this text, with a space between two ASCII neighbours (ct-transformer-online.cpp:46-53)
  std::string full;
  for (const std::string& w : *cache) full += w;
  const char* in = text ? text : "";
  if (!full.empty() && !(static_cast<unsigned char>(full.back()) & 0x80) && in[0] != 0 && !(static_cast<unsigned char>(in[0]) & 0x80)) full += " ";
  full += in;
  std::vector<std::string> pieces;
  std::vector<int32_t> ids;
  tok.Tokenize(full.c_str(), &pieces, &ids);
  const int n_cache = (int)cache->size();   // the VAD position of every network call of this request
  const int n_total = (int)std::ceil((float)ids.size() / kMiniSentence);
  std::vector<int32_t> remain_ids, punc_all;
  std::vector<std::string> remain_str, words;
  for (size_t pos = 0; pos < ids.size(); pos += kMiniSentence) {
    const size_t end = std::min(ids.size(), pos + (size_t)kMiniSentence);
    std::vector<int32_t> in_ids(remain_ids);
    in_ids.insert(in_ids.end(), ids.begin() + pos, ids.begin() + end);
    std::vector<std::string> in_str(remain_str);
    in_str.insert(in_str.end(), pieces.begin() + pos, pieces.begin() + end);
    std::vector<int32_t> punc = infer(in_ids, n_cache);
    if (punc.size() != in_ids.size()) return "";
    if ((int)(pos / kMiniSentence) < n_total - 1) {
      int sent_end = -1, last_comma = -1;
      for (int k = (int)punc.size() - 2; k > 0; --k) {
        const std::string& p = tok.Id2Punc(punc[k]);
        if (p == tok.Id2Punc(kPeriod) || p == tok.Id2Punc(kQuestion)) { sent_end = k; break; }
        if (last_comma < 0 && p == tok.Id2Punc(kComma)) last_comma = k;
      }
      if (sent_end < 0 && (int)in_str.size() > kCachePopLimit && last_comma > 0) {
        sent_end = last_comma;
        punc[sent_end] = kPeriod;
      }
      remain_str.assign(in_str.begin() + (sent_end + 1), in_str.end());
      remain_ids.assign(in_ids.begin() + (sent_end + 1), in_ids.end());
      in_str.resize((size_t)(sent_end + 1));
      punc.resize((size_t)(sent_end + 1));
    }
    words.insert(words.end(), in_str.begin(), in_str.end());
    punc_all.insert(punc_all.end(), punc.begin(), punc.end());
  }
  // output: the words after the cached ones, each followed by its mark unless that is "_"; the mark of the LAST cached word is
  // emitted too (the skip counter reaches the cache size on that word, ct-transformer-online.cpp:106-121)
  std::string out_last;          // the last piece appended, to drop a trailing mark
  std::vector<std::string> out;
  int skipped = 0;
  for (size_t i = 0; i < words.size(); ++i) {
    if (!HighBit(words[i]) && i + 1 < words.size() && !HighBit(words[i + 1])) words[i] += " ";
    if (skipped < n_cache) ++skipped; else out.push_back(words[i]);
    if (skipped >= n_cache) {
      const std::string& mark = tok.Id2Punc(punc_all[i]);
      if (mark != "_") out.push_back(mark);
    }
  }
  int sent_end = -1;
  for (int i = (int)punc_all.size() - 2; i > 0; --i)
    if (punc_all[i] == kPeriod || punc_all[i] == kQuestion) { sent_end = i; break; }
  cache->assign(words.begin() + (sent_end + 1), words.end());
  if (!out.empty()) {
    bool is_mark = false;
    for (int k = 0; k < tok.NumPunc(); ++k) is_mark = is_mark || out.back() == tok.Id2Punc(k);
    if (is_mark) out.pop_back();   // a trailing mark waits for the next call
  }
  std::string res;
  for (const std::string& w : out) res += w;
  return res;
}

CTTransformerOnlineB200::~CTTransformerOnlineB200() {
  if (engine_) b200pf_punc_destroy(engine_);
}

bool CTTransformerOnlineB200::Init(const std::string& punc_dir, std::string* err) {
  std::vector<std::string> tokens, punc;
  if (!ReadStringArray(punc_dir + "/tokens.json", &tokens) || tokens.empty()) { if (err) *err = punc_dir + "/tokens.json: not a JSON array of strings"; return false; }
  if (!ReadStringArray(punc_dir + "/punc_list.json", &punc) || punc.empty()) { if (err) *err = punc_dir + "/punc_list.json: not a JSON array of strings"; return false; }
  if (b200pf_punc_create(punc_dir.c_str(), device_, max_tokens_, &engine_) != 0) { if (err) *err = b200pf_last_error(); return false; }
  int vocab = 0, n_punc = 0;
  b200pf_punc_info(engine_, &vocab, &n_punc, nullptr, &max_tokens_);
  if (vocab != (int)tokens.size() || n_punc != (int)punc.size() || n_punc < 6) {
    if (err) *err = "punctuation model / tokens.json / punc_list.json sizes disagree";
    b200pf_punc_destroy(engine_);
    engine_ = nullptr;
    return false;
  }
  tok_.Open(tokens, punc);
  return true;
}

void CTTransformerOnlineB200::InitPunc(const std::string& punc_model, const std::string& punc_config, const std::string& token_file, int thread_num) {
  (void)punc_config; (void)token_file; (void)thread_num;
  std::string err;
  if (!Init(DirOf(punc_model), &err)) {
    fprintf(stderr, "Error when load punc model: %s\n", err.c_str());
    exit(-1);
  }
}

std::string CTTransformerOnlineB200::AddPunc(const char* sz_input, std::vector<std::string>& arr_cache, std::string language) {
  (void)language;   // the realtime model does not map symbols for en-bpe (ct-transformer-online.cpp:40-137)
  if (!engine_) { fprintf(stderr, "CTTransformerOnlineB200: not initialised\n"); return ""; }
  b200pf_punc* eng = engine_;
  return AddPuncOnlineWith(tok_, sz_input, &arr_cache, [eng](const std::vector<int32_t>& ids, int vad_pos) {
    std::vector<int32_t> punc(ids.size(), 0);
    const int32_t offs[2] = {0, (int32_t)ids.size()};
    const int32_t vp[1] = {vad_pos};
    if (ids.empty() || b200pf_punc_infer_vad(eng, ids.data(), offs, vp, 1, punc.data(), nullptr) != 0) {
      fprintf(stderr, "Error when run punc onnx forword: %s\n", b200pf_last_error());
      punc.clear();
    }
    return punc;
  });
}

}  // namespace funasr_b200
