// Readers for the model directory the reference's FunOfflineInit consumes
// (onnxruntime/include/com-define.h:52-88, onnxruntime/src/offline-stream.cpp:60-87):
//   am.mvn       Paraformer::LoadCmvn            onnxruntime/src/paraformer.cpp:325-360
//   config.yaml  Paraformer::LoadConfigFromYaml  onnxruntime/src/paraformer.cpp:178-200 (frontend_conf.fs, lang)
//   tokens.json  Vocab::LoadVocabFromJson        onnxruntime/src/vocab.cpp:46-63
// plus this implementation's flat fp32 tensor file model.b200pf (format: asr-2pass_b200/modelfile.py).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace pf {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
};

struct WeightFile {
  std::map<std::string, double> cfg;
  std::map<std::string, HostTensor> tensors;
};

// All return false and fill `err` on failure.
bool read_weight_file(const std::string& path, WeightFile* out, std::string* err);
bool read_am_mvn(const std::string& path, std::vector<float>* means, std::vector<float>* vars, std::string* err);
bool read_tokens_json(const std::string& path, std::vector<std::string>* tokens, std::string* err);
// fs defaults to 16000 and lang to "zh-cn" when the keys are absent (paraformer.h:104,122).
bool read_config_yaml(const std::string& path, int* fs, std::string* lang, std::string* err);

}  // namespace pf
