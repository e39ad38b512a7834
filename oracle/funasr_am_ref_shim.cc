// ORACLE-SIDE (test infrastructure; only tests/, smoke() and bench.py's cpu_baseline leg may use anything under oracle/).
//
// C entry points over the reference's OWN compiled acoustic-model and punctuation host classes -- funasr::Paraformer
// (onnxruntime/src/paraformer.cpp: InitAsr :21-52, LfrCmvn :378-418, Forward :420-582, CompileHotwordEmbedding :592-693) and
// funasr::CTTransformer (ct-transformer.cpp: AddPunc :40-157, Infer :164-203) -- built where the sources lie (oracle/Makefile,
// target `ref`).  Their onnxruntime sessions are served by oracle/fake_ort.cc: the network behind a session is whatever callback
// the test registered (the oracle's numpy/torch restatement of the graph), everything around it is the reference's code.
#include "precomp.h"

namespace google {
// LogMessage symbols the sources reference through LOG(...) (the vendored glog library is not built).
static std::ostringstream g_sink;
LogMessageTime::LogMessageTime() : time_struct_(), timestamp_(0), usecs_(0), gmtoffset_(0) {}
LogMessage::LogMessage(const char*, int) : allocated_(nullptr), data_(nullptr) {}
LogMessage::LogMessage(const char*, int, int) : allocated_(nullptr), data_(nullptr) {}
LogMessage::~LogMessage() { g_sink.str(""); }
std::ostream& LogMessage::stream() { return g_sink; }
}  // namespace google

static int CopyOut(const std::string& s, char* out, int cap) {
  if ((int)s.size() + 1 > cap) return -1;
  memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

extern "C" {
void* ref_am_create(const char* am_model, const char* am_cmvn, const char* am_config, const char* token_file) {
  funasr::Paraformer* p = new funasr::Paraformer();
  p->InitAsr(am_model, am_cmvn, am_config, token_file, 1);
  return p;
}
void ref_am_destroy(void* h) { delete (funasr::Paraformer*)h; }
void ref_am_init_hw(void* h, const char* hw_model) { ((funasr::Paraformer*)h)->InitHwCompiler(hw_model, 1); }
void ref_am_init_seg_dict(void* h, const char* path) { ((funasr::Paraformer*)h)->InitSegDict(path); }
// Paraformer::Forward(float** din, int* len, true, hw_emb, nullptr, 1): pcm in the reference's float form (int16 / 32768).
int ref_am_forward(void* h, const float* pcm, int n, const float* hw, int n_hw, int dim, char* out, int cap) {
  std::vector<std::vector<float>> emb;
  for (int j = 0; j < n_hw; ++j) emb.emplace_back(hw + (size_t)j * dim, hw + (size_t)(j + 1) * dim);
  if (emb.empty()) emb.push_back(std::vector<float>(1, 0.0f));
  float* din[1] = {const_cast<float*>(pcm)};
  int len[1] = {n};
  std::vector<std::string> r = ((funasr::Paraformer*)h)->Forward(din, len, true, emb, nullptr, 1);
  return CopyOut(r.empty() ? std::string() : r[0], out, cap);
}
int ref_am_compile_hotwords(void* h, const char* hotwords, float* out, int cap_rows, int dim) {
  std::string hw(hotwords);
  std::vector<std::vector<float>> emb = ((funasr::Paraformer*)h)->CompileHotwordEmbedding(hw);
  if ((int)emb.size() > cap_rows) return -1;
  for (size_t j = 0; j < emb.size(); ++j) {
    if ((int)emb[j].size() != dim) return -2;
    memcpy(out + j * dim, emb[j].data(), dim * sizeof(float));
  }
  return (int)emb.size();
}

void* ref_punc_create(const char* punc_model, const char* punc_config, const char* token_file) {
  funasr::CTTransformer* p = new funasr::CTTransformer();
  p->InitPunc(punc_model, punc_config, token_file, 1);
  return p;
}
void ref_punc_destroy(void* h) { delete (funasr::CTTransformer*)h; }
int ref_punc_add(void* h, const char* text, const char* lang, char* out, int cap) {
  return CopyOut(((funasr::CTTransformer*)h)->AddPunc(text, std::string(lang)), out, cap);
}

// funasr::CTTransformerOnline (ct-transformer-online.cpp): the realtime punctuation model the 2-pass server uses for both legs
// (funasrruntime.cpp:545,610).  The word cache travels as one string, words separated by '\x01'.
void* ref_punc_online_create(const char* punc_model, const char* punc_config, const char* token_file) {
  funasr::CTTransformerOnline* p = new funasr::CTTransformerOnline();
  p->InitPunc(punc_model, punc_config, token_file, 1);
  return p;
}
void ref_punc_online_destroy(void* h) { delete (funasr::CTTransformerOnline*)h; }
int ref_punc_online_add(void* h, const char* text, const char* cache_in, const char* lang, char* out, int cap, char* cache_out, int cache_cap) {
  std::vector<std::string> cache;
  std::string cur;
  for (const char* c = cache_in; *c; ++c) {
    if (*c == '\x01') { cache.push_back(cur); cur.clear(); } else cur += *c;
  }
  const int n = CopyOut(((funasr::CTTransformerOnline*)h)->AddPunc(text, cache, std::string(lang)), out, cap);
  std::string joined;
  for (const std::string& w : cache) joined += w + '\x01';
  if (CopyOut(joined, cache_out, cache_cap) < 0) return -1;
  return n;
}

// funasr::FsmnVad + funasr::FsmnVadOnline (fsmn-vad.cpp, fsmn-vad-online.cpp): the VAD cut of the offline path exactly as
// Audio::CutSplit drives it (audio.cpp:1172-1226) -- the recording in 1 s pieces through FsmnVadOnline::Infer (the reference's own
// online fbank / LFR caches, FSMN caches through the session, E2E scorer), start / end frames paired into segments (ms).
void* ref_vad_create(const char* vad_model, const char* vad_cmvn, const char* vad_config) {
  funasr::FsmnVad* v = new funasr::FsmnVad();
  v->InitVad(vad_model, vad_cmvn, vad_config, 1);
  return v;
}
void ref_vad_destroy(void* h) { delete (funasr::FsmnVad*)h; }
int ref_vad_cutsplit(void* h, const float* pcm, int speech_len, int vad_tail_sil, int vad_max_len, int* seg_ms, int cap) {
  funasr::FsmnVad* vad = (funasr::FsmnVad*)h;
  vad->SetConfig(vad_tail_sil, vad_max_len);                       // FunOfflineInferBuffer, funasrruntime.cpp:219-220
  std::unique_ptr<funasr::VadModel> online(new funasr::FsmnVadOnline(vad));
  const int dest_sample_rate = 16000;
  int step = dest_sample_rate * 1;
  bool is_final = false;
  std::vector<std::vector<int>> vad_segments;
  for (int sample_offset = 0; sample_offset < speech_len; sample_offset += std::min(step, speech_len - sample_offset)) {
    if (sample_offset + step >= speech_len - 1) {
      step = speech_len - sample_offset;
      is_final = true;
    } else {
      is_final = false;
    }
    std::vector<float> pcm_data(pcm + sample_offset, pcm + sample_offset + step);
    std::vector<std::vector<int>> cut = online->Infer(pcm_data, is_final);
    vad_segments.insert(vad_segments.end(), cut.begin(), cut.end());
  }
  int start = -1, end = -1, n = 0;
  for (const std::vector<int>& sg : vad_segments) {
    if (sg.size() != 2) break;
    if (sg[0] != -1) start = sg[0];
    if (sg[1] != -1) end = sg[1];
    if (start != -1 && end != -1) {
      if (n >= cap) return -1;
      seg_ms[2 * n] = start;
      seg_ms[2 * n + 1] = end;
      ++n;
      start = -1;
      end = -1;
    }
  }
  return n;
}

// The reference's exported offline API itself (funasrruntime.cpp: FunOfflineInit :36-40, FunOfflineInferBuffer :208-340) on an
// OfflineStream built from real model directories (offline-stream.cpp): LoadPcmwav, CutSplit, length sort, FetchDynamic,
// Paraformer::Forward per segment, stitching, punctuation, TimestampSentence -- with every session served by fake_ort.cc.
void* ref_offline_init(const char* const* keys, const char* const* values, int n) {
  std::map<std::string, std::string> mp;
  for (int i = 0; i < n; ++i) mp[keys[i]] = values[i];
  return FunOfflineInit(mp, 1, false, 1);
}
void ref_offline_uninit(void* h) { FunOfflineUninit(h); }
int ref_offline_infer_buffer(void* h, const char* buf, int n_bytes, int vad_tail_sil, int vad_max_len, char* text, int text_cap, char* stamp,
                             int stamp_cap, char* sents, int sents_cap) {
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  FUNASR_RESULT r = FunOfflineInferBuffer(h, buf, n_bytes, RASR_NONE, nullptr, hw, 16000, "pcm", true, vad_tail_sil, vad_max_len);
  if (!r) return -1;
  const int n = CopyOut(FunASRGetResult(r, 0), text, text_cap);
  if (CopyOut(FunASRGetStamp(r), stamp, stamp_cap) < 0 || CopyOut(FunASRGetStampSents(r), sents, sents_cap) < 0) return -2;
  FunASRFreeResult(r);
  return n;
}
}  // extern "C"
