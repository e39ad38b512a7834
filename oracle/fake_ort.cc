// ORACLE-SIDE (test infrastructure; only tests/, smoke() and bench.py's cpu_baseline leg may use anything under oracle/).
//
// A stand-in for libonnxruntime's C API (ORT 1.14, the version the reference vendors under
// websocket/onnxruntime-linux-x64-1.14.0/include), just wide enough for the reference's own Paraformer / CTTransformer classes
// (onnxruntime/src/paraformer.cpp, ct-transformer.cpp) to run UNMODIFIED with their neural network replaced by a callback:
// OrtApi::Run hands the input tensors to a function registered for the session's model path and returns what it produces.
// Everything the reference does around the session -- fbank, LFR + CMVN, greedy search, timestamps, detokenisation, hotword id
// packing, the punctuation mini-sentence cache -- is then the reference's own compiled code, which is what the oracle's restatement
// of those steps is pinned against (oracle/am_ref.py).  onnxruntime itself is not in this image; no ONNX graph is evaluated here.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "onnxruntime_c_api.h"

extern "C" {
struct fake_ort_tensor {
  int32_t type;       // ONNXTensorElementDataType
  int32_t ndim;
  int64_t shape[8];
  void* data;
};
// cb(user, run context, n_in, inputs): produces the outputs with fake_ort_set_output(ctx, index, ...); returns 0 on success.
typedef int (*fake_ort_run_fn)(void* user, void* ctx, int n_in, const fake_ort_tensor* in);
}

struct OrtEnv { int unused; };
struct OrtStatus { std::string msg; };
struct OrtSessionOptions { int unused; };
struct OrtMemoryInfo { int unused; };
struct OrtTensorTypeAndShapeInfo {
  ONNXTensorElementDataType type;
  std::vector<int64_t> shape;
};
struct OrtValue {
  ONNXTensorElementDataType type;
  std::vector<int64_t> shape;
  void* data;
  std::vector<char> owned;
};
struct OrtSession {
  std::string path;
  int n_in, n_out;
  fake_ort_run_fn fn;
  void* user;
  std::vector<std::string> in_names, out_names;
};

namespace {
struct Registration {
  std::string suffix;
  int n_in, n_out;
  fake_ort_run_fn fn;
  void* user;
};
std::mutex g_mu;
std::vector<Registration> g_reg;
struct RunCtx {
  std::vector<OrtValue*> out;
};

size_t ElemSize(ONNXTensorElementDataType t) {
  switch (t) {
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_FLOAT: return 4;
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_INT32: return 4;
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_INT64: return 8;
    case ONNX_TENSOR_ELEMENT_DATA_TYPE_DOUBLE: return 8;
    default: return 0;
  }
}
OrtStatus* Err(const char* m) { OrtStatus* s = new OrtStatus; s->msg = m; return s; }

template <size_t I>
[[noreturn]] void Unsupported() {
  fprintf(stderr, "fake_ort: the reference called OrtApi slot %zu (member order of onnxruntime_c_api.h), which this stand-in does not provide\n", I);
  abort();
}
template <size_t... I>
void FillTraps(void** slots, std::index_sequence<I...>) {
  void* t[] = {reinterpret_cast<void*>(&Unsupported<I>)...};
  for (size_t i = 0; i < sizeof...(I); ++i) slots[i] = t[i];
}

OrtStatus* ORT_API_CALL CreateEnv(OrtLoggingLevel, const char*, OrtEnv** out) { *out = new OrtEnv; return nullptr; }
OrtStatus* ORT_API_CALL SetLanguageProjection(const OrtEnv*, OrtLanguageProjection) { return nullptr; }
OrtStatus* ORT_API_CALL CreateSessionOptions(OrtSessionOptions** out) { *out = new OrtSessionOptions; return nullptr; }
OrtStatus* ORT_API_CALL SetIntraOpNumThreads(OrtSessionOptions*, int) { return nullptr; }
OrtStatus* ORT_API_CALL SetInterOpNumThreads(OrtSessionOptions*, int) { return nullptr; }
OrtStatus* ORT_API_CALL SetSessionGraphOptimizationLevel(OrtSessionOptions*, GraphOptimizationLevel) { return nullptr; }
OrtStatus* ORT_API_CALL DisableCpuMemArena(OrtSessionOptions*) { return nullptr; }
OrtStatus* ORT_API_CALL CreateSession(const OrtEnv*, const ORTCHAR_T* path, const OrtSessionOptions*, OrtSession** out) {
  std::lock_guard<std::mutex> lk(g_mu);
  const std::string p(path);
  for (auto it = g_reg.rbegin(); it != g_reg.rend(); ++it) {
    if (p.size() >= it->suffix.size() && p.compare(p.size() - it->suffix.size(), it->suffix.size(), it->suffix) == 0) {
      OrtSession* s = new OrtSession;
      s->path = p; s->n_in = it->n_in; s->n_out = it->n_out; s->fn = it->fn; s->user = it->user;
      for (int i = 0; i < s->n_in; ++i) s->in_names.push_back("in" + std::to_string(i));
      for (int i = 0; i < s->n_out; ++i) s->out_names.push_back("out" + std::to_string(i));
      *out = s;
      return nullptr;
    }
  }
  return Err(("fake_ort: no network registered for " + p).c_str());
}
OrtStatus* ORT_API_CALL SessionGetInputCount(const OrtSession* s, size_t* out) { *out = (size_t)s->n_in; return nullptr; }
OrtStatus* ORT_API_CALL SessionGetOutputCount(const OrtSession* s, size_t* out) { *out = (size_t)s->n_out; return nullptr; }
OrtStatus* ORT_API_CALL SessionGetInputName(const OrtSession* s, size_t i, OrtAllocator* a, char** v) {
  const std::string& n = s->in_names.at(i);
  *v = (char*)a->Alloc(a, n.size() + 1);
  memcpy(*v, n.c_str(), n.size() + 1);
  return nullptr;
}
OrtStatus* ORT_API_CALL SessionGetOutputName(const OrtSession* s, size_t i, OrtAllocator* a, char** v) {
  const std::string& n = s->out_names.at(i);
  *v = (char*)a->Alloc(a, n.size() + 1);
  memcpy(*v, n.c_str(), n.size() + 1);
  return nullptr;
}
OrtMemoryInfo g_cpu_info;
void* ORT_API_CALL AllocFn(OrtAllocator*, size_t n) { return malloc(n); }
void ORT_API_CALL FreeFn(OrtAllocator*, void* p) { free(p); }
const OrtMemoryInfo* ORT_API_CALL InfoFn(const OrtAllocator*) { return &g_cpu_info; }
OrtAllocator g_alloc = {ORT_API_VERSION, AllocFn, FreeFn, InfoFn};
OrtStatus* ORT_API_CALL GetAllocatorWithDefaultOptions(OrtAllocator** out) { *out = &g_alloc; return nullptr; }
OrtStatus* ORT_API_CALL AllocatorFree(OrtAllocator* a, void* p) { a->Free(a, p); return nullptr; }
OrtStatus* ORT_API_CALL AllocatorAlloc(OrtAllocator* a, size_t n, void** out) { *out = a->Alloc(a, n); return nullptr; }
OrtStatus* ORT_API_CALL CreateCpuMemoryInfo(OrtAllocatorType, OrtMemType, OrtMemoryInfo** out) { *out = new OrtMemoryInfo; return nullptr; }
OrtStatus* ORT_API_CALL CreateTensorWithDataAsOrtValue(const OrtMemoryInfo*, void* data, size_t, const int64_t* shape, size_t nd,
                                                       ONNXTensorElementDataType type, OrtValue** out) {
  OrtValue* v = new OrtValue;
  v->type = type;
  v->shape.assign(shape, shape + nd);
  v->data = data;
  *out = v;
  return nullptr;
}
OrtStatus* ORT_API_CALL GetTensorMutableData(OrtValue* v, void** out) { *out = v->data; return nullptr; }
OrtStatus* ORT_API_CALL GetTensorTypeAndShape(const OrtValue* v, OrtTensorTypeAndShapeInfo** out) {
  OrtTensorTypeAndShapeInfo* i = new OrtTensorTypeAndShapeInfo;
  i->type = v->type;
  i->shape = v->shape;
  *out = i;
  return nullptr;
}
OrtStatus* ORT_API_CALL GetTensorElementType(const OrtTensorTypeAndShapeInfo* i, ONNXTensorElementDataType* out) { *out = i->type; return nullptr; }
OrtStatus* ORT_API_CALL GetDimensionsCount(const OrtTensorTypeAndShapeInfo* i, size_t* out) { *out = i->shape.size(); return nullptr; }
OrtStatus* ORT_API_CALL GetDimensions(const OrtTensorTypeAndShapeInfo* i, int64_t* d, size_t n) {
  for (size_t k = 0; k < n && k < i->shape.size(); ++k) d[k] = i->shape[k];
  return nullptr;
}
OrtStatus* ORT_API_CALL GetTensorShapeElementCount(const OrtTensorTypeAndShapeInfo* i, size_t* out) {
  size_t n = 1;
  for (int64_t d : i->shape) n *= (size_t)d;
  *out = n;
  return nullptr;
}
OrtStatus* ORT_API_CALL IsTensor(const OrtValue*, int* out) { *out = 1; return nullptr; }
OrtStatus* ORT_API_CALL Run(OrtSession* s, const OrtRunOptions*, const char* const*, const OrtValue* const* in, size_t n_in,
                            const char* const*, size_t n_out, OrtValue** out) {
  std::vector<fake_ort_tensor> t(n_in);
  for (size_t i = 0; i < n_in; ++i) {
    t[i].type = (int32_t)in[i]->type;
    t[i].ndim = (int32_t)in[i]->shape.size();
    for (int k = 0; k < 8; ++k) t[i].shape[k] = k < t[i].ndim ? in[i]->shape[k] : 0;
    t[i].data = in[i]->data;
  }
  RunCtx ctx;
  ctx.out.assign(n_out, nullptr);
  const int rc = s->fn(s->user, &ctx, (int)n_in, t.data());
  bool ok = rc == 0;
  for (size_t i = 0; i < n_out; ++i) ok = ok && ctx.out[i] != nullptr;
  if (!ok) {
    for (OrtValue* v : ctx.out) delete v;
    return Err("fake_ort: the registered network failed or left an output unset");
  }
  for (size_t i = 0; i < n_out; ++i) out[i] = ctx.out[i];
  return nullptr;
}
OrtErrorCode ORT_API_CALL GetErrorCode(const OrtStatus*) { return ORT_FAIL; }
const char* ORT_API_CALL GetErrorMessage(const OrtStatus* s) { return s->msg.c_str(); }
void ORT_API_CALL ReleaseEnv(OrtEnv* p) { delete p; }
void ORT_API_CALL ReleaseStatus(OrtStatus* p) { delete p; }
void ORT_API_CALL ReleaseMemoryInfo(OrtMemoryInfo* p) { delete p; }
void ORT_API_CALL ReleaseSession(OrtSession* p) { delete p; }
void ORT_API_CALL ReleaseValue(OrtValue* p) { delete p; }
void ORT_API_CALL ReleaseRunOptions(OrtRunOptions*) {}
void ORT_API_CALL ReleaseTensorTypeAndShapeInfo(OrtTensorTypeAndShapeInfo* p) { delete p; }
void ORT_API_CALL ReleaseSessionOptions(OrtSessionOptions* p) { delete p; }

alignas(OrtApi) unsigned char g_api_mem[sizeof(OrtApi)];   // OrtApi has no default constructor
OrtApi& g_api = *reinterpret_cast<OrtApi*>(g_api_mem);
bool g_api_ready = false;
const OrtApi* ORT_API_CALL GetApi(uint32_t) {
  if (!g_api_ready) {
    // every member of OrtApi is a function pointer: the ones the reference does not reach trap loudly
    void** slots = reinterpret_cast<void**>(&g_api);
    FillTraps(slots, std::make_index_sequence<sizeof(OrtApi) / sizeof(void*)>());
#define SET(name) g_api.name = name
    SET(CreateEnv); SET(SetLanguageProjection); SET(CreateSessionOptions); SET(SetIntraOpNumThreads); SET(SetInterOpNumThreads);
    SET(SetSessionGraphOptimizationLevel); SET(DisableCpuMemArena); SET(CreateSession); SET(SessionGetInputCount);
    SET(SessionGetOutputCount); SET(SessionGetInputName); SET(SessionGetOutputName); SET(GetAllocatorWithDefaultOptions);
    SET(AllocatorFree); SET(AllocatorAlloc); SET(CreateCpuMemoryInfo); SET(CreateTensorWithDataAsOrtValue); SET(GetTensorMutableData);
    SET(GetTensorTypeAndShape); SET(GetTensorElementType); SET(GetDimensionsCount); SET(GetDimensions);
    SET(GetTensorShapeElementCount); SET(IsTensor); SET(Run); SET(GetErrorCode); SET(GetErrorMessage); SET(ReleaseEnv);
    SET(ReleaseStatus); SET(ReleaseMemoryInfo); SET(ReleaseSession); SET(ReleaseValue); SET(ReleaseRunOptions);
    SET(ReleaseTensorTypeAndShapeInfo); SET(ReleaseSessionOptions);
#undef SET
    g_api_ready = true;
  }
  return &g_api;
}
const char* ORT_API_CALL GetVersionString() { return "1.14.0-standin"; }
const OrtApiBase g_base = {GetApi, GetVersionString};
}  // namespace

extern "C" {
const OrtApiBase* ORT_API_CALL OrtGetApiBase(void) NO_EXCEPTION { return &g_base; }

// Sessions created for a model path ending in `suffix` run `fn`; later registrations win.
void fake_ort_register(const char* suffix, int n_in, int n_out, fake_ort_run_fn fn, void* user) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_reg.push_back(Registration{suffix, n_in, n_out, fn, user});
}
// Called from inside a run callback: output `index` = a copy of `data` (shape[ndim], element type `type`).
int fake_ort_set_output(void* ctx_, int index, int type, int ndim, const int64_t* shape, const void* data) {
  RunCtx* ctx = (RunCtx*)ctx_;
  if (index < 0 || index >= (int)ctx->out.size()) return -1;
  const size_t es = ElemSize((ONNXTensorElementDataType)type);
  if (es == 0) return -2;
  size_t n = 1;
  for (int k = 0; k < ndim; ++k) n *= (size_t)shape[k];
  OrtValue* v = new OrtValue;
  v->type = (ONNXTensorElementDataType)type;
  v->shape.assign(shape, shape + ndim);
  v->owned.resize(n * es ? n * es : 1);
  if (n) memcpy(v->owned.data(), data, n * es);
  v->data = v->owned.data();
  delete ctx->out[index];
  ctx->out[index] = v;
  return 0;
}
}
