// MultiGpuParaformer — one funasr::Model handle over all B200s of a box.
//
// Segments are independent units (ParaformerTorch::Forward treats batch items independently after the model call,
// onnxruntime/src/paraformer-torch.cpp:431-467), so the path shards with no exchange step: weights are replicated,
// every GPU has its own engine, stream and worker thread (an independent per-GPU queue), and one Forward() call is
// split over the GPUs by longest-processing-time-first on the per-segment FLOP estimate c1*T + c2*T^2 (the attention
// term).  Results come back in the caller's order, like the reference's index_vector un-permute
// (onnxruntime/src/funasrruntime.cpp:270-279).  There is no collective and no inter-GPU traffic.
//
// The reference's servers are single-process (one FunOfflineInit handle shared by decoder-thread-num threads,
// websocket/bin/websocket-server.cpp:387-403), which is why this lives behind the same seam instead of in a launcher.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "paraformer_b200.h"

namespace funasr_b200 {

// Longest-processing-time-first assignment of segments (by sample count) to n_dev queues.  assign[i] = queue of
// segment i.  Pure host arithmetic (also exported for tests).
void PartitionSegments(const int* len, int n, int n_dev, std::vector<int>* assign);
double SegmentCost(int n_samples);  // relative FLOP estimate of one segment

class MultiGpuParaformer : public Model {
 public:
  MultiGpuParaformer(const std::vector<int>& devices, int max_rows = 0, int max_segments = 0);
  ~MultiGpuParaformer() override;

  bool Init(const std::string& model_dir, std::string* err);
  void InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config, const std::string& token_file,
               int thread_num) override;
  std::vector<std::string> Forward(float** din, int* len, bool input_finished, const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                                   void* wfst_decoder = nullptr, int batch_in = 1) override;
  std::string Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                      void* wfst_decoder = nullptr) override;
  // int16 segments by pointer (the stand-alone shim's VAD cut points): same sharding, each GPU's worker copies its own share
  std::vector<std::string> ForwardSegments16(const int16_t* const* seg, const int64_t* len, int n_seg,
                                             const std::vector<std::vector<float>>& hw_emb = {{0.0}});
  void StartUtterance() override {}
  void EndUtterance() override {}
  void Reset() override {}
  std::string Rescoring() override { return ""; }
  void InitHwCompiler(const std::string& hw_model, int thread_num) override;
  void InitSegDict(const std::string& seg_dict_model) override;
  std::vector<std::vector<float>> CompileHotwordEmbedding(std::string& hotwords) override;  // on the first GPU
  std::string GetLang() override;
  int GetAsrSampleRate() override;
  void SetBatchSize(int batch_size) override;
  int GetBatchSize() override;

  int n_devices() const { return (int)models_.size(); }
  ParaformerB200* model(int i) { return models_[i].get(); }
  // segments decoded per device so far (diagnostics / tests)
  std::vector<long long> segments_per_device();

 private:
  struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> jobs;
    bool stop = false;
    long long segments = 0;
  };
  void Post(int dev, std::function<void()> job);
  void AssembleText(const std::vector<const SegmentRaw*>& by_index, std::vector<std::string>* results);
  static void Loop(Worker* w);

  std::vector<int> devices_;
  int max_rows_, max_segments_;
  std::vector<std::unique_ptr<ParaformerB200>> models_;
  std::vector<std::unique_ptr<Worker>> workers_;
  std::mutex text_mu_;   // the one detokeniser (models_[0]) that turns every call's token ids into text, in the caller's order
};

}  // namespace funasr_b200
