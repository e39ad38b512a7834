// C hooks (declared in include/b200pf_host.h) that expose the C++ host mirror to ctypes-based tests and to
// bench.py's end-to-end leg: detokeniser, timestamp/post-processing, and the FunOffline* call sequence.
#include <algorithm>
#include <cstring>
#include <memory>

#include "../../../include/b200pf_host.h"
#include <chrono>
#include <thread>

#include "funasrruntime_b200.h"
#include "logprob_adapter.h"
#include "micro_batcher.h"
#include "multi_gpu.h"
#include "paraformer_b200.h"
#include "punc_b200.h"
#include "vad_segmenter.h"

namespace {
int CopyOut(const std::string& s, char* out, int cap) {
  if (!out || cap <= 0) return (int)s.size();
  const int n = std::min<int>((int)s.size(), cap - 1);
  memcpy(out, s.data(), n);
  out[n] = 0;
  return (int)s.size();
}
struct Detok { pf::host::Detokenizer d; explicit Detok(std::vector<std::string> t) : d(std::move(t)) {} };
}  // namespace

extern "C" {

void* b200pf_host_detok_create(const char* const* tokens, int n) {
  std::vector<std::string> t(tokens, tokens + n);
  return new Detok(std::move(t));
}
void b200pf_host_detok_destroy(void* h) { delete (Detok*)h; }
int b200pf_host_detok_text(void* h, const int32_t* ids, int n, const char* lang, char* out, int cap) {
  std::vector<int> v(ids, ids + n);
  return CopyOut(((Detok*)h)->d.ToText(v, lang ? lang : ""), out, cap);
}
int b200pf_host_detok_text_state(void* h, const int32_t* ids, int n, const char* lang, int state_in, int* state_out, char* out, int cap) {
  std::vector<int> v(ids, ids + n);
  bool ended = state_in != 0;
  const std::string t = static_cast<const pf::host::Detokenizer&>(((Detok*)h)->d).ToText(v, lang ? lang : "", state_in != 0, &ended);
  if (state_out) *state_out = ended ? 1 : 0;
  return CopyOut(t, out, cap);
}
int b200pf_host_timestamp_text(void* h, const int32_t* ids, int n, const float* us_alphas, const float* us_peaks, int n_frames,
                               char* out, int cap) {
  Detok* d = (Detok*)h;
  std::vector<int> v(ids, ids + n);
  std::vector<std::string> pieces = d->d.ToPieces(v);
  std::vector<std::string> raw = pieces;
  std::vector<float> al(us_alphas, us_alphas + n_frames), pk(us_peaks, us_peaks + n_frames);
  std::vector<pf::host::Span> spans = pf::host::TimestampFromPeaks(&al, pk, &pieces, nullptr);
  return CopyOut(pf::host::MergeWithStamps(raw, spans), out, cap);
}
int b200pf_host_stitch(const char* const* msgs, const float* start_s, int n, const char* lang, char* text, int text_cap, char* stamp,
                       int stamp_cap) {
  std::vector<std::string> m(msgs, msgs + n);
  std::vector<float> st(start_s, start_s + n);
  std::string t, s;
  pf::host::StitchSegments(m, st, lang ? lang : "", &t, &s);
  CopyOut(t, text, text_cap);
  CopyOut(s, stamp, stamp_cap);
  return 0;
}

void* b200pf_host_offline_init(const char* model_dir, int device, int max_rows, int max_segments, int batch_size) {
  std::map<std::string, std::string> mp;
  mp["model-dir"] = model_dir;
  mp["device"] = std::to_string(device);
  mp["max-rows"] = std::to_string(max_rows);
  mp["max-segments"] = std::to_string(max_segments);
  return FunOfflineInit(mp, 1, true, batch_size);
}
void* b200pf_host_offline_init_devices(const char* model_dir, const int* devices, int n_dev, int max_rows, int max_segments, int batch_size) {
  std::map<std::string, std::string> mp;
  mp["model-dir"] = model_dir;
  std::string d;
  for (int i = 0; i < n_dev; ++i) d += (i ? "," : "") + std::to_string(devices[i]);
  mp["devices"] = d;
  mp["max-rows"] = std::to_string(max_rows);
  mp["max-segments"] = std::to_string(max_segments);
  return FunOfflineInit(mp, 1, true, batch_size);
}
void* b200pf_host_offline_init_vad(const char* model_dir, const char* vad_dir, int device, int max_rows, int max_segments, int batch_size,
                                   float speech_noise_thres) {
  std::map<std::string, std::string> mp;
  mp["model-dir"] = model_dir;
  mp["vad-dir"] = vad_dir;
  mp["device"] = std::to_string(device);
  mp["max-rows"] = std::to_string(max_rows);
  mp["max-segments"] = std::to_string(max_segments);
  if (speech_noise_thres > 0.f) mp["vad-speech-noise-thres"] = std::to_string(speech_noise_thres);
  return FunOfflineInit(mp, 1, true, batch_size);
}
int b200pf_host_offline_vad_cut(void* h, const int16_t* pcm, int64_t n_samples, int vad_tail_sil, int vad_max_len, int* seg_ms, int cap) {
  std::vector<std::pair<int, int>> segs;
  if (FunOfflineVadSegmentsB200(h, pcm, n_samples, vad_tail_sil, vad_max_len, &segs) != 0) return -1;
  if ((int)segs.size() > cap) return -2;
  for (size_t i = 0; i < segs.size(); ++i) { seg_ms[2 * i] = segs[i].first; seg_ms[2 * i + 1] = segs[i].second; }
  return (int)segs.size();
}
int b200pf_host_offline_infer_buffer_vad(void* h, const char* buf, int n_bytes, int vad_tail_sil, int vad_max_len, char* text, int text_cap,
                                         char* stamp, int stamp_cap) {
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  FUNASR_RESULT r = FunOfflineInferBuffer(h, buf, n_bytes, RASR_NONE, nullptr, hw, 16000, "pcm", true, vad_tail_sil, vad_max_len);
  if (!r) return -1;
  const int n = CopyOut(FunASRGetResult(r, 0), text, text_cap);
  CopyOut(FunASRGetStamp(r), stamp, stamp_cap);
  FunASRFreeResult(r);
  return n;
}
void* b200pf_host_offline_init_kv(const char* const* keys, const char* const* values, int n, int batch_size) {
  std::map<std::string, std::string> mp;
  for (int i = 0; i < n; ++i) mp[keys[i]] = values[i];
  return FunOfflineInit(mp, 1, true, batch_size);
}
int b200pf_host_offline_infer_full(void* h, const char* buf, int n_bytes, int vad_tail_sil, int vad_max_len, char* text, int text_cap,
                                   char* stamp, int stamp_cap, char* stamp_sents, int sents_cap) {
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  FUNASR_RESULT r = FunOfflineInferBuffer(h, buf, n_bytes, RASR_NONE, nullptr, hw, 16000, "pcm", true, vad_tail_sil, vad_max_len);
  if (!r) return -1;
  const int n = CopyOut(FunASRGetResult(r, 0), text, text_cap);
  if (stamp) CopyOut(FunASRGetStamp(r), stamp, stamp_cap);
  if (stamp_sents) CopyOut(FunASRGetStampSents(r), stamp_sents, sents_cap);
  FunASRFreeResult(r);
  return n;
}
// ---- 2-pass stream hooks (FunTpassInit / FunTpassOnlineInit / FunTpassInferBuffer) ----
void* b200pf_host_tpass_init_kv(const char* const* keys, const char* const* values, int n) {
  std::map<std::string, std::string> mp;
  for (int i = 0; i < n; ++i) mp[keys[i]] = values[i];
  return FunTpassInit(mp, 1);
}
void* b200pf_host_tpass_online_init(void* tpass) { return FunTpassOnlineInit(tpass); }
void b200pf_host_tpass_uninit(void* tpass) { FunTpassUninit(tpass); }
void b200pf_host_tpass_online_uninit(void* online) { FunTpassOnlineUninit(online); }
// One FunTpassInferBuffer call.  cache_io carries punc_cache[1] (the offline leg's realtime-punctuation word cache) in and out as
// '\n'-joined words.  Returns the length of tpass_msg (>= 0) or -1 when the call returned nullptr.
int b200pf_host_tpass_infer(void* tpass, void* online, const char* buf, int n_bytes, int input_finished, int mode, int vad_tail_sil,
                            int vad_max_len, char* cache_io, int cache_cap, char* msg, int msg_cap, char* tpass_msg, int tpass_cap, char* stamp,
                            int stamp_cap, char* stamp_sents, int sents_cap) {
  std::vector<std::vector<std::string>> cache(2);
  if (cache_io) {
    std::string c(cache_io), cur;
    for (char ch : c) { if (ch == '\n') { cache[1].push_back(cur); cur.clear(); } else cur.push_back(ch); }
    if (!cur.empty()) cache[1].push_back(cur);
  }
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  FUNASR_RESULT r = FunTpassInferBuffer(tpass, online, buf, n_bytes, cache, input_finished != 0, 16000, "pcm", (ASR_TYPE)mode, hw, true, vad_tail_sil,
                                        vad_max_len);
  if (!r) return -1;
  if (msg) CopyOut(FunASRGetResult(r, 0), msg, msg_cap);
  const int n = CopyOut(FunASRGetTpassResult(r, 0), tpass_msg, tpass_cap);
  if (stamp) CopyOut(FunASRGetStamp(r), stamp, stamp_cap);
  if (stamp_sents) CopyOut(FunASRGetStampSents(r), stamp_sents, sents_cap);
  FunASRFreeResult(r);
  if (cache_io) {
    std::string joined;
    for (size_t i = 0; i < cache[1].size(); ++i) { if (i) joined.push_back('\n'); joined += cache[1][i]; }
    CopyOut(joined.c_str(), cache_io, cache_cap);
  }
  return n;
}
// The VAD state machine fed in chunks (pf::host::StreamingVad): chunk_len[k] frames per Push, the last one flagged final.
// out receives [start_ms, end_ms] pairs; returns the number of segments.
int b200pf_host_vad_segments_streaming(const float* sil_prob, int n_frames, const int* chunk_len, int n_chunks, int max_end_sil_ms, int max_seg_ms,
                                       float thres, int* out, int cap) {
  pf::host::VadOptions vo;
  vo.max_end_silence_ms = max_end_sil_ms; vo.max_single_segment_ms = max_seg_ms; vo.speech_noise_thres = thres;
  pf::host::StreamingVad sv(vo);
  int pos = 0, n_out = 0;
  for (int k = 0; k < n_chunks; ++k) {
    const int len = std::min(chunk_len[k], n_frames - pos);
    const bool fin = k == n_chunks - 1;
    for (const auto& sg : sv.Push(sil_prob + pos, len, fin)) {
      if (n_out < cap) { out[2 * n_out] = sg.first; out[2 * n_out + 1] = sg.second; }
      ++n_out;
    }
    pos += len;
  }
  return n_out;
}
int b200pf_host_expand_posteriors(const float* topk_logprob, const int32_t* topk_ids, int rows, int k, int vocab, float* dense) {
  std::vector<float> d;
  pf::host::ExpandPrunedPosteriors(topk_logprob, topk_ids, rows, k, vocab, &d);
  memcpy(dense, d.data(), d.size() * sizeof(float));
  return 0;
}
int b200pf_host_partition(const int* len, int n, int n_dev, int* assign) {
  std::vector<int> a;
  funasr_b200::PartitionSegments(len, n, n_dev, &a);
  for (int i = 0; i < n; ++i) assign[i] = a[i];
  return 0;
}
int b200pf_host_segments_per_device(void* h_offline, long long* out, int cap) {
  funasr_b200::MultiGpuParaformer* p = FunOfflinePoolB200(h_offline);
  if (!p) return 0;
  const std::vector<long long> v = p->segments_per_device();
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
void b200pf_host_offline_uninit(void* h) { FunOfflineUninit(h); }
int b200pf_host_offline_infer_buffer(void* h, const char* buf, int n_bytes, int vad_max_len, char* text, int text_cap, float* snippet_s) {
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  FUNASR_RESULT r = FunOfflineInferBuffer(h, buf, n_bytes, RASR_NONE, nullptr, hw, 16000, "pcm", true, 800, vad_max_len);
  if (!r) return -1;
  const int n = CopyOut(FunASRGetResult(r, 0), text, text_cap);
  if (snippet_s) *snippet_s = FunASRGetRetSnippetTime(r);
  FunASRFreeResult(r);
  return n;
}
int b200pf_host_offline_infer_segments(void* h, const int16_t* pcm, int64_t n_samples, const int64_t* seg_begin, const int64_t* seg_end,
                                       int n_seg, char* text, int text_cap) {
  FUNASR_RESULT r = FunOfflineInferSegmentsB200(h, pcm, n_samples, (const long long*)seg_begin, (const long long*)seg_end, n_seg);
  if (!r) return -1;
  const int n = CopyOut(FunASRGetResult(r, 0), text, text_cap);
  FunASRFreeResult(r);
  return n;
}
int b200pf_host_offline_infer_buffer_hw(void* h, const char* buf, int n_bytes, int vad_max_len, const float* hw, int n_hw, int dim,
                                        char* text, int text_cap, char* stamp, int stamp_cap) {
  std::vector<std::vector<float>> emb;
  for (int j = 0; j < n_hw; ++j) emb.emplace_back(hw + (size_t)j * dim, hw + (size_t)(j + 1) * dim);
  FUNASR_RESULT r = FunOfflineInferBuffer(h, buf, n_bytes, RASR_NONE, nullptr, emb, 16000, "pcm", true, 800, vad_max_len);
  if (!r) return -1;
  const int n = CopyOut(FunASRGetResult(r, 0), text, text_cap);
  CopyOut(FunASRGetStamp(r), stamp, stamp_cap);
  FunASRFreeResult(r);
  return n;
}
// CompileHotwordEmbedding(handle, hotwords) (funasrruntime.h:118): rows are written to out [cap_rows][dim]; returns the row count.
int b200pf_host_compile_hotwords(void* h_offline, const char* hotwords, float* out, int cap_rows, int dim) {
  std::string hw(hotwords ? hotwords : "");
  const std::vector<std::vector<float>> emb = CompileHotwordEmbedding(h_offline, hw);
  if ((int)emb.size() > cap_rows) return -1;
  for (size_t j = 0; j < emb.size(); ++j) {
    if ((int)emb[j].size() != dim) return -1;
    memcpy(out + j * dim, emb[j].data(), (size_t)dim * sizeof(float));
  }
  return (int)emb.size();
}
int b200pf_host_pack_hotwords(const char* const* tokens, int n_tokens, const char* seg_dict_path, const char* hotwords, int32_t* ids,
                              int32_t* lengths, int cap_rows) {
  std::unordered_map<std::string, int> token_id;
  for (int i = 0; i < n_tokens; ++i) token_id.emplace(tokens[i], i);   // PhoneSet: first id wins (phone-set.cpp:51-56)
  std::unordered_map<std::string, std::vector<std::string>> seg_dict;
  if (seg_dict_path && seg_dict_path[0]) funasr_b200::LoadSegDict(seg_dict_path, &seg_dict);
  std::vector<int32_t> m, l;
  funasr_b200::PackHotwords(hotwords ? hotwords : "", token_id, seg_dict, &m, &l);
  if ((int)l.size() > cap_rows) return -1;
  memcpy(ids, m.data(), m.size() * sizeof(int32_t));
  memcpy(lengths, l.data(), l.size() * sizeof(int32_t));
  return (int)l.size();
}
int b200pf_host_sentence_stamps(const char* text, const char* stamp, char* out, int cap) {
  return CopyOut(pf::host::SentenceStamps(text ? text : "", stamp ? stamp : ""), out, cap);
}
// ---- punctuation (CTTransformer mirror) ----
int b200pf_host_punc_tokenize(const char* const* tokens, int n_tokens, const char* text, int32_t* ids, int cap) {
  funasr_b200::PuncTokenizer tk;
  tk.Open(std::vector<std::string>(tokens, tokens + n_tokens), std::vector<std::string>());
  std::vector<std::string> pieces;
  std::vector<int32_t> v;
  tk.Tokenize(text, &pieces, &v);
  if ((int)v.size() > cap) return -1;
  if (!v.empty()) memcpy(ids, v.data(), v.size() * sizeof(int32_t));
  return (int)v.size();
}
int b200pf_host_punc_add_scripted(const char* const* tokens, int n_tokens, const char* const* punc_list, int n_punc, const char* text,
                                  const char* lang, int seed, int every, char* out, int cap) {
  funasr_b200::PuncTokenizer tk;
  tk.Open(std::vector<std::string>(tokens, tokens + n_tokens), std::vector<std::string>(punc_list, punc_list + n_punc));
  // the scripted stand-in network of tests/test_punc.py::scripted_punc
  auto infer = [&](const std::vector<int32_t>& ids) {
    std::vector<int32_t> r(ids.size());
    for (size_t pos = 0; pos < ids.size(); ++pos) {
      uint32_t h = (uint32_t)((uint64_t)(uint32_t)ids[pos] * 2654435761ull + (uint64_t)pos * 40503ull + (uint64_t)seed * 97ull);
      h = (h >> 7) & 0xFFFF;
      if (every && h % (uint32_t)every == 0) r[pos] = h % 3 ? 3 : 4;
      else if (h % 11 == 1) r[pos] = 2;
      else if (h % 37 == 2) r[pos] = 0;
      else r[pos] = 1;
    }
    return r;
  };
  return CopyOut(funasr_b200::AddPuncWith(tk, text, lang ? lang : "zh-cn", infer), out, cap);
}
namespace {
std::vector<std::string> SplitCache(const char* s) {
  std::vector<std::string> out;
  std::string cur;
  for (const char* c = s ? s : ""; *c; ++c) {
    if (*c == '\x01') { out.push_back(cur); cur.clear(); } else cur += *c;
  }
  return out;
}
std::string JoinCache(const std::vector<std::string>& v) {
  std::string s;
  for (const std::string& w : v) s += w + '\x01';
  return s;
}
}  // namespace
int b200pf_host_punc_online_add_scripted(const char* const* tokens, int n_tokens, const char* const* punc_list, int n_punc, const char* text,
                                         const char* cache_in, int seed, int every, char* out, int cap, char* cache_out, int cache_cap) {
  funasr_b200::PuncTokenizer tk;
  tk.Open(std::vector<std::string>(tokens, tokens + n_tokens), std::vector<std::string>(punc_list, punc_list + n_punc));
  std::vector<std::string> cache = SplitCache(cache_in);
  auto infer = [&](const std::vector<int32_t>& ids, int vad_pos) {   // tests/test_punc.py::scripted_punc with seed + 13 * vad_pos
    std::vector<int32_t> r(ids.size());
    const int sd = seed + 13 * vad_pos;
    for (size_t pos = 0; pos < ids.size(); ++pos) {
      uint32_t h = (uint32_t)((uint64_t)(uint32_t)ids[pos] * 2654435761ull + (uint64_t)pos * 40503ull + (uint64_t)sd * 97ull);
      h = (h >> 7) & 0xFFFF;
      if (every && h % (uint32_t)every == 0) r[pos] = h % 3 ? 3 : 4;
      else if (h % 11 == 1) r[pos] = 2;
      else if (h % 37 == 2) r[pos] = 0;
      else r[pos] = 1;
    }
    return r;
  };
  const int n = CopyOut(funasr_b200::AddPuncOnlineWith(tk, text, &cache, infer), out, cap);
  if (CopyOut(JoinCache(cache), cache_out, cache_cap) < 0) return -1;
  return n;
}
void* b200pf_host_punc_online_create(const char* punc_dir, int device, int max_tokens) {
  std::unique_ptr<funasr_b200::CTTransformerOnlineB200> p(new funasr_b200::CTTransformerOnlineB200(device, max_tokens));
  std::string err;
  if (!p->Init(punc_dir, &err)) { fprintf(stderr, "b200pf_host_punc_online_create: %s\n", err.c_str()); return nullptr; }
  return p.release();
}
void b200pf_host_punc_online_destroy(void* h) { delete (funasr_b200::CTTransformerOnlineB200*)h; }
int b200pf_host_punc_online_add(void* h, const char* text, const char* cache_in, char* out, int cap, char* cache_out, int cache_cap) {
  if (!h) return -1;
  std::vector<std::string> cache = SplitCache(cache_in);
  const int n = CopyOut(((funasr_b200::CTTransformerOnlineB200*)h)->AddPunc(text, cache, "zh-cn"), out, cap);
  if (CopyOut(JoinCache(cache), cache_out, cache_cap) < 0) return -1;
  return n;
}
void* b200pf_host_punc_create(const char* punc_dir, int device, int max_tokens) {
  std::unique_ptr<funasr_b200::CTTransformerB200> p(new funasr_b200::CTTransformerB200(device, max_tokens));
  std::string err;
  if (!p->Init(punc_dir, &err)) { fprintf(stderr, "b200pf_host_punc_create: %s\n", err.c_str()); return nullptr; }
  return p.release();
}
long long b200pf_host_punc_rounds(void* h) { return h ? ((funasr_b200::CTTransformerB200*)h)->rounds() : 0; }
void b200pf_host_punc_destroy(void* h) { delete (funasr_b200::CTTransformerB200*)h; }
int b200pf_host_punc_add(void* h, const char* text, const char* lang, char* out, int cap) {
  if (!h) return -1;
  return CopyOut(((funasr_b200::PuncModel*)(funasr_b200::CTTransformerB200*)h)->AddPunc(text, std::string(lang ? lang : "zh-cn")).c_str(), out, cap);
}
int b200pf_host_punc_add_batch(void* h, const char* const* texts, int n, const char* lang, char* out, int cap, int* rounds) {
  if (!h) return -1;
  std::vector<std::string> in(texts, texts + n);
  const std::vector<std::string> r = ((funasr_b200::CTTransformerB200*)h)->AddPuncBatch(in, lang ? lang : "zh-cn", rounds);
  size_t used = 0;
  for (const std::string& s : r) {     // results back to back, each NUL-terminated
    if (used + s.size() + 1 > (size_t)cap) return -1;
    memcpy(out + used, s.c_str(), s.size() + 1);
    used += s.size() + 1;
  }
  return (int)used;
}
int b200pf_host_init_seg_dict(void* h_offline, const char* path) {
  funasr_b200::ParaformerB200* m = FunOfflineModelB200(h_offline);
  if (!m || !path) return -1;
  m->InitSegDict(path);
  return 0;
}
int b200pf_host_model_forward_hw(void* h_offline, const float* const* din, const int* len, int n, const float* hw, int n_hw, int dim,
                                 char* out, int cap) {
  funasr_b200::Model* m = FunOfflineModel(h_offline);
  if (!m) return -1;
  std::vector<float*> ptrs(n);
  for (int i = 0; i < n; ++i) ptrs[i] = const_cast<float*>(din[i]);
  std::vector<int> l(len, len + n);
  std::vector<std::vector<float>> emb;
  for (int j = 0; j < n_hw; ++j) emb.emplace_back(hw + (size_t)j * dim, hw + (size_t)(j + 1) * dim);
  std::vector<std::string> r = m->Forward(ptrs.data(), l.data(), true, emb, nullptr, n);
  std::string joined;
  for (int i = 0; i < n; ++i) { if (i) joined += "\n"; joined += r[i]; }
  return CopyOut(joined, out, cap);
}
// Times `iters` calls of Model::Forward(float**, int*, ...) from C++ (std::chrono around the virtual call, strings included): what a
// reference caller pays per call, without the ctypes marshalling of the Python test harness.  ms_out = mean milliseconds per call;
// returns the number of non-empty result strings of the last call.
int b200pf_host_model_forward_timed(void* h_offline, const float* const* din, const int* len, int n, int iters, double* ms_out) {
  funasr_b200::Model* m = FunOfflineModel(h_offline);
  if (!m || iters <= 0) return -1;
  std::vector<float*> ptrs(n);
  for (int i = 0; i < n; ++i) ptrs[i] = const_cast<float*>(din[i]);
  std::vector<int> l(len, len + n);
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  int nonempty = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < iters; ++it) {
    std::vector<std::string> r = m->Forward(ptrs.data(), l.data(), true, hw, nullptr, n);
    if (it == iters - 1)
      for (const auto& x : r) nonempty += !x.empty();
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (ms_out) *ms_out = std::chrono::duration<double, std::milli>(t1 - t0).count() / iters;
  return nonempty;
}
// Model::Forward over float segments (the plugin seam itself): returns the '\n'-joined result strings.
int b200pf_host_model_forward(void* h_offline, const float* const* din, const int* len, int n, char* out, int cap) {
  funasr_b200::Model* m = FunOfflineModel(h_offline);
  if (!m) return -1;
  std::vector<float*> ptrs(n);
  for (int i = 0; i < n; ++i) ptrs[i] = const_cast<float*>(din[i]);
  std::vector<int> l(len, len + n);
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  std::vector<std::string> r = m->Forward(ptrs.data(), l.data(), true, hw, nullptr, n);
  std::string joined;
  for (int i = 0; i < n; ++i) { if (i) joined += "\n"; joined += r[i]; }
  return CopyOut(joined, out, cap);
}


// FunASRInit -> FunASRInfer(file) or FunASRInferBuffer(buf) -> FunASRGetResult -> FunASRUninit (funasrruntime.h:60-78)
int b200pf_host_funasr_infer(const char* model_dir, int device, int max_rows, const char* wav_path, const char* buf, int n_bytes, char* out, int cap) {
  std::map<std::string, std::string> mp;
  mp["model-dir"] = model_dir;
  mp["device"] = std::to_string(device);
  mp["max-rows"] = std::to_string(max_rows);
  FUNASR_HANDLE h = FunASRInit(mp, 1);
  if (!h) return -1;
  FUNASR_RESULT r = wav_path ? FunASRInfer(h, wav_path, RASR_NONE, nullptr) : FunASRInferBuffer(h, buf, n_bytes, RASR_NONE, nullptr);
  int n = -2;
  if (r) { n = CopyOut(FunASRGetResult(r, 0), out, cap); FunASRFreeResult(r); }
  FunASRUninit(h);
  return n;
}

// pf::host::SegmentVad (E2EVadModel's job, e2e-vad.h): silence probabilities -> [start_ms, end_ms] pairs; returns the count
int b200pf_host_vad_segments(const float* sil_prob, int n_frames, int max_end_sil_ms, int max_seg_ms, float thres, int* out, int cap) {
  pf::host::VadOptions o;
  o.max_end_silence_ms = max_end_sil_ms; o.max_single_segment_ms = max_seg_ms; o.speech_noise_thres = thres;
  const std::vector<std::pair<int, int>> segs = pf::host::SegmentVad(sil_prob, n_frames, o);
  for (size_t i = 0; i < segs.size() && (int)i < cap; ++i) { out[2 * i] = segs[i].first; out[2 * i + 1] = segs[i].second; }
  return (int)segs.size();
}

// ---- MicroBatcher ------------------------------------------------------------
// Mock inner model for host-only tests: "n=<len>;b=<batch size>;x=<first sample>;hw=<rows of hw_emb>", after
// sleeping latency_us once per batch (a stand-in for the GPU forward).
void* b200pf_host_mb_create_mock(int max_wait_us, int max_batch, int max_rows, int latency_us) {
  funasr_b200::MicroBatcherOptions o;
  o.max_wait_us = max_wait_us; o.max_batch = max_batch; o.max_rows = max_rows;
  return new funasr_b200::MicroBatcher(
      [latency_us](float** din, int* len, int n, const std::vector<std::vector<float>>& hw) {
        if (latency_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(latency_us));
        std::vector<std::string> out(n);
        for (int i = 0; i < n; ++i) {
          if (i > 0 && len[i] < len[i - 1]) { out[i] = "unsorted"; continue; }
          out[i] = "n=" + std::to_string(len[i]) + ";b=" + std::to_string(n) + ";x=" + std::to_string(len[i] > 0 ? (int)din[i][0] : 0) +
                   ";hw=" + std::to_string(hw.size());
        }
        return out;
      },
      o);
}
void* b200pf_host_mb_create(void* h_offline, int max_wait_us, int max_batch, int max_rows) {
  funasr_b200::MicroBatcherOptions o;
  o.max_wait_us = max_wait_us; o.max_batch = max_batch; o.max_rows = max_rows;
  if (funasr_b200::MultiGpuParaformer* pool = FunOfflinePoolB200(h_offline)) {
    // connections -> micro-batcher -> per-GPU queues: one batch is split over all GPUs of the handle
    return new funasr_b200::MicroBatcher(
        [pool](float** din, int* len, int n, const std::vector<std::vector<float>>& hw) { return pool->Forward(din, len, true, hw, nullptr, n); }, o);
  }
  funasr_b200::ParaformerB200* m = FunOfflineModelB200(h_offline);
  if (!m) return nullptr;
  return new funasr_b200::MicroBatcher(m, o);
}
void b200pf_host_mb_destroy(void* mb) { delete (funasr_b200::MicroBatcher*)mb; }
// One connection's batch-1 call: Model::Forward(buff, len, true, hw_emb, dec_handle, 1) (funasrruntime.cpp:570-586).
int b200pf_host_mb_forward(void* mb, const float* pcm, int len, const float* hw, int n_hw, int dim, char* out, int cap) {
  std::vector<std::vector<float>> emb;
  for (int j = 0; j < n_hw; ++j) emb.emplace_back(hw + (size_t)j * dim, hw + (size_t)(j + 1) * dim);
  if (n_hw == 0) emb.assign(1, std::vector<float>(1, 0.0f));  // the reference's default argument {{0.0}}
  float* one[1] = {const_cast<float*>(pcm)};
  int l[1] = {len};
  std::vector<std::string> r = ((funasr_b200::MicroBatcher*)mb)->Forward(one, l, true, emb, nullptr, 1);
  return CopyOut(r.empty() ? std::string() : r[0], out, cap);
}
// segments, batches, closed_by_deadline, closed_by_size, max_batch_seen, mean wait us, max wait us
int b200pf_host_mb_stats(void* mb, double* out7) {
  const funasr_b200::MicroBatcherStats s = ((funasr_b200::MicroBatcher*)mb)->stats();
  out7[0] = (double)s.segments; out7[1] = (double)s.batches; out7[2] = (double)s.closed_by_deadline; out7[3] = (double)s.closed_by_size;
  out7[4] = (double)s.max_batch_seen; out7[5] = s.segments ? s.wait_us_sum / (double)s.segments : 0.0; out7[6] = s.wait_us_max;
  return 0;
}

}  // extern "C"
