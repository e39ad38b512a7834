// ORACLE-SIDE SHIM (test infrastructure): C entry points into the REFERENCE's own host text code, compiled in place from
// /root/reference by oracle/Makefile (target `ref`) into oracle/_ref/libfunasr_text_ref.so:
//     funasr::Vocab::Vector2StringV2 / Vector2String   onnxruntime/src/vocab.cpp:98-104,164-305
//     funasr::TimestampOnnx, funasr::PostProcess        onnxruntime/src/util.cpp:720-963
//     funasr::ParaformerOnline::GetPosEmb / CifSearch   onnxruntime/src/paraformer-online.cpp:240-345
//     funasr::E2EVadModel (scores -> speech segments)   onnxruntime/src/e2e-vad.h
// Nothing of the reference is copied: this file only calls it.  Two things are supplied here because the reference gets
// them from CMake: a stand-in <gflags/gflags.h> (oracle/stubs; the vendored gflags header is generated) and the few glog
// LogMessage symbols the sources reference through LOG(...) (the vendored glog library is not built).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

// ParaformerOnline::GetPosEmb / CifSearch (paraformer-online.cpp:240-345) are private members; they are reached by opening
// the access specifiers for the reference headers only (the standard headers above are already included and guarded).
#define private public
#define protected public
#include "precomp.h"
#undef private
#undef protected

namespace google {
static std::ostringstream g_sink;
LogMessageTime::LogMessageTime() : time_struct_(), timestamp_(0), usecs_(0), gmtoffset_(0) {}
LogMessage::LogMessage(const char*, int) : allocated_(nullptr), data_(nullptr) {}
LogMessage::LogMessage(const char*, int, int) : allocated_(nullptr), data_(nullptr) {}
LogMessage::~LogMessage() { g_sink.str(""); }
std::ostream& LogMessage::stream() { return g_sink; }
}  // namespace google

namespace {
int CopyOut(const std::string& s, char* out, int cap) {
  if (out && cap > 0) {
    const int n = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return (int)s.size();
}
}  // namespace

extern "C" {

void* ref_vocab_create(const char* tokens_json) { return new funasr::Vocab(tokens_json); }
void ref_vocab_destroy(void* v) { delete (funasr::Vocab*)v; }

// Paraformer::GreedySearch without stamps (paraformer.cpp:396-397): Vector2StringV2(ids, language)
int ref_vector2string_v2(void* v, const int* ids, int n, const char* lang, char* out, int cap) {
  std::vector<int> in(ids, ids + n);
  return CopyOut(((funasr::Vocab*)v)->Vector2StringV2(in, lang ? lang : ""), out, cap);
}

// Paraformer::GreedySearch with stamps (paraformer.cpp:398-407): Vector2String -> TimestampOnnx -> PostProcess
int ref_greedy_with_stamps(void* v, const int* ids, int n, const float* us_alphas, const float* us_peaks, int n_frames, char* out, int cap) {
  std::vector<int> in(ids, ids + n);
  std::vector<std::string> char_list;
  std::vector<std::vector<float>> timestamp_list;
  std::string res_str;
  ((funasr::Vocab*)v)->Vector2String(in, char_list);
  std::vector<std::string> raw_char(char_list);
  std::vector<float> al(us_alphas, us_alphas + n_frames), pk(us_peaks, us_peaks + n_frames);
  funasr::TimestampOnnx(al, pk, char_list, res_str, timestamp_list);
  return CopyOut(funasr::PostProcess(raw_char, timestamp_list), out, cap);
}

// funasr::TimestampSentence (util.cpp:569-637): punctuated text + "[[b,e],...]" -> the stamp_sents JSON array
int ref_timestamp_sentence(const char* text, const char* stamp, char* out, int cap) {
  std::string t(text), s(stamp);
  return CopyOut(funasr::TimestampSentence(t, s), out, cap);
}

// ---- ParaformerOnline::GetPosEmb / CifSearch, called on an object whose constructor (ORT sessions) is bypassed --------
// Zero-filled storage is a valid state for the std::vector members of libstdc++; the scalar members the two functions
// read are set explicitly.  chunk_size {0, T, 0} + is_last_chunk reproduces the offline predictor: no alpha is zeroed
// and the tail frame (alpha = tail_alphas, zero hidden) is appended (paraformer-online.cpp:281-300).
void* ref_online_create(float cif_threshold, float tail_alphas) {
  void* mem = calloc(1, sizeof(funasr::ParaformerOnline));
  funasr::ParaformerOnline* po = (funasr::ParaformerOnline*)mem;
  po->cif_threshold = cif_threshold;
  po->tail_alphas = tail_alphas;
  return po;
}
void ref_online_destroy(void* h) { free(h); }  // no destructor: nothing but vectors we clear below

// hidden [T][D], alphas [T] -> frames written to out [cap_tok][D]; returns the number of fired tokens
int ref_cif_search(void* h, const float* hidden, const float* alphas, int T, int D, float* out, int cap_tok) {
  funasr::ParaformerOnline* po = (funasr::ParaformerOnline*)h;
  po->chunk_size.clear();
  po->chunk_size.push_back(0); po->chunk_size.push_back(T); po->chunk_size.push_back(0);
  po->is_last_chunk = true;
  po->hidden_cache_.clear();
  po->alphas_cache_.clear();
  std::vector<std::vector<float>> hid(T, std::vector<float>(D));
  for (int t = 0; t < T; ++t) std::copy(hidden + (size_t)t * D, hidden + (size_t)(t + 1) * D, hid[t].begin());
  std::vector<float> al(alphas, alphas + T);
  std::vector<std::vector<float>> frames;
  po->CifSearch(hid, al, true, frames);
  const int n = (int)frames.size();
  for (int i = 0; i < n && i < cap_tok; ++i) std::copy(frames[i].begin(), frames[i].end(), out + (size_t)i * D);
  po->hidden_cache_.clear();
  po->alphas_cache_.clear();
  std::vector<int>().swap(po->chunk_size);
  return n;
}

// feats [T][D] += position encoding for positions 1..T (start_idx_cache_ = 0), in place
void ref_pos_emb(void* h, float* feats, int T, int D) {
  funasr::ParaformerOnline* po = (funasr::ParaformerOnline*)h;
  po->start_idx_cache_ = 0;
  std::vector<std::vector<float>> f(T, std::vector<float>(D));
  for (int t = 0; t < T; ++t) std::copy(feats + (size_t)t * D, feats + (size_t)(t + 1) * D, f[t].begin());
  po->GetPosEmb(f, T, D);
  for (int t = 0; t < T; ++t) std::copy(f[t].begin(), f[t].end(), feats + (size_t)t * D);
}

// ---- E2EVadModel (onnxruntime/src/e2e-vad.h, header-only) -------------------------------------------------------------
// sil_prob [T]: probability of pdf 0 per 10 ms frame (the only column the model reads, sil_pdf_ids = {0}).  The waveform only
// feeds the decibel gate, which the default options switch off (decibel_thres = snr_thres = -100), so silence of the right
// length is passed.  out receives [start_ms, end_ms] pairs; the return value is their count.
static std::vector<std::vector<float>> ScoreRows(const float* sil_prob, int a, int b) {
  std::vector<std::vector<float>> rows;
  for (int t = a; t < b; ++t) rows.push_back(std::vector<float>{sil_prob[t], 1.0f - sil_prob[t]});
  return rows;
}

// one call over the whole recording: is_final = true, online = false (FsmnVad::Infer's form, fsmn-vad.cpp:226-240)
int ref_e2e_vad_offline(const float* sil_prob, int T, int max_end_sil, int max_seg_ms, float thres, int* out, int cap) {
  funasr::E2EVadModel m;
  std::vector<float> wave((size_t)T * 160 + 240, 0.f);
  std::vector<std::vector<int>> segs = m(ScoreRows(sil_prob, 0, T), wave, true, false, max_end_sil, max_seg_ms, thres, 16000);
  int n = 0;
  for (auto& s : segs) { if (n < cap) { out[2 * n] = s[0]; out[2 * n + 1] = s[1]; } ++n; }
  return n;
}

// the way Audio::CutSplit drives it (audio.cpp:1172-1226): chunk_frames scores per call through ONE model object with
// online = true, is_final on the last chunk, then CutSplit's pairing of (start, -1) / (-1, end) into segments
int ref_e2e_vad_online(const float* sil_prob, int T, int chunk_frames, int max_end_sil, int max_seg_ms, float thres, int* out, int cap) {
  funasr::E2EVadModel m;
  std::vector<std::vector<int>> vad_segments;
  for (int a = 0; a < T; a += chunk_frames) {
    const int b = std::min(T, a + chunk_frames);
    std::vector<float> wave((size_t)(b - a) * 160 + 240, 0.f);
    std::vector<std::vector<int>> cut = m(ScoreRows(sil_prob, a, b), wave, b == T, true, max_end_sil, max_seg_ms, thres, 16000);
    vad_segments.insert(vad_segments.end(), cut.begin(), cut.end());
  }
  int n = 0, start_i = -1, end_i = -1;
  for (auto& seg : vad_segments) {
    if (seg.size() != 2) break;
    if (seg[0] != -1) start_i = seg[0];
    if (seg[1] != -1) end_i = seg[1];
    if (start_i != -1 && end_i != -1) {
      if (n < cap) { out[2 * n] = start_i; out[2 * n + 1] = end_i; }
      ++n;
      start_i = -1; end_i = -1;
    }
  }
  return n;
}

}  // extern "C"
