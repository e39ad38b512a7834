// ORACLE-SIDE SHIM (test infrastructure): C entry points into the REFERENCE's own host text code, compiled in place from
// /root/reference by oracle/Makefile (target `ref`) into oracle/_ref/libfunasr_text_ref.so:
//     funasr::Vocab::Vector2StringV2 / Vector2String   onnxruntime/src/vocab.cpp:98-104,164-305
//     funasr::TimestampOnnx, funasr::PostProcess        onnxruntime/src/util.cpp:720-963
// Nothing of the reference is copied: this file only calls it.  Two things are supplied here because the reference gets
// them from CMake: a stand-in <gflags/gflags.h> (oracle/stubs; the vendored gflags header is generated) and the few glog
// LogMessage symbols the sources reference through LOG(...) (the vendored glog library is not built).
#include <sstream>
#include <string>
#include <vector>

#include "precomp.h"

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

}  // extern "C"
