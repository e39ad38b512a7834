// MicroBatcher — cross-connection segment micro-batching for the 2-pass offline leg.
//
// The reference's 2-pass server decodes every closed VAD segment with a batch-1 call from the connection's own
// strand: FunTpassInferBuffer -> `asr_handle->Forward(buff, len, true, hw_emb, dec_handle, 1)`
// (onnxruntime/src/funasrruntime.cpp:570-586; strands: websocket/bin/websocket-server-2pass.cpp:270,536-552).
// On a CPU that is the right shape (one ORT session run per decoder thread).  On a B200 a batch-1 forward is ~600
// dependent kernel launches of a few microseconds each, i.e. latency bound: the GPU is filled only when segments
// of MANY connections travel together.
//
// MicroBatcher sits in the same seam (it implements funasr::Model's offline virtuals, so it can be what
// TpassStream::asr_handle / OfflineStream::asr_handle points to) in front of ONE batched model:
//   * every Forward() call enqueues its segments and blocks on their results;
//   * a dispatcher thread closes a batch when the OLDEST waiting segment has waited `max_wait_us`, or earlier when
//     `max_batch` segments / `max_rows` packed LFR rows are waiting;
//   * a batch holds only segments with the same hotword matrix (hw_emb is per connection, paraformer.cpp:515-531),
//     is sorted ascending by length like Audio::CutSplit does (audio.cpp:1233-1238), runs as ONE batched Forward of
//     the inner model, and the results are handed back to their callers in their original order.
// Results are identical to direct calls: the B200 engine is batch invariant (a segment's tokens do not depend on what
// it is batched with; tests/test_gpu_parity.py::test_batch_invariance_and_input_formats).
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "paraformer_b200.h"

namespace funasr_b200 {

struct MicroBatcherOptions {
  int max_wait_us = 5000;    // deadline for the oldest waiting segment.  5 ms: a batch-1 forward itself takes 4 ms, and the sweep in
                             // profiles/r02_microbatch_sweep.txt gives 41.8 k RTFx / 3.9 ms mean wait at 5 ms against 22.2 k / 18.7 ms at 20 ms
  int max_batch = 256;       // segments per batch
  int max_rows = 32768;      // packed LFR rows (T + 1 per segment) per batch; must not exceed the engine's max_rows
};

struct MicroBatcherStats {
  int64_t segments = 0;        // segments decoded
  int64_t batches = 0;         // inner Forward calls
  int64_t closed_by_deadline = 0, closed_by_size = 0;
  int64_t max_batch_seen = 0;
  double wait_us_sum = 0.0;    // queueing delay (enqueue -> batch start), summed over segments
  double wait_us_max = 0.0;
};

class MicroBatcher : public Model {
 public:
  // The batched call the dispatcher makes: same meaning as Model::Forward(float**, int*, true, hw_emb, nullptr, n).
  typedef std::function<std::vector<std::string>(float** din, int* len, int n, const std::vector<std::vector<float>>& hw_emb)> BatchFn;

  MicroBatcher(BatchFn fn, const MicroBatcherOptions& opt);
  // Wraps a ParaformerB200 (not owned); the other Model virtuals are forwarded to it.
  MicroBatcher(ParaformerB200* inner, const MicroBatcherOptions& opt);
  ~MicroBatcher() override;

  // Blocks until every one of the call's segments has been decoded.  Safe to call from any number of threads.
  std::vector<std::string> Forward(float** din, int* len, bool input_finished, const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                                   void* wfst_decoder = nullptr, int batch_in = 1) override;
  std::string Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb = {{0.0}},
                      void* wfst_decoder = nullptr) override;

  void StartUtterance() override {}
  void EndUtterance() override {}
  void Reset() override {}
  std::string Rescoring() override { return ""; }
  void InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config, const std::string& token_file,
               int thread_num) override;
  void InitHwCompiler(const std::string& hw_model, int thread_num) override { if (inner_) inner_->InitHwCompiler(hw_model, thread_num); }
  void InitSegDict(const std::string& seg_dict_model) override { if (inner_) inner_->InitSegDict(seg_dict_model); }
  std::vector<std::vector<float>> CompileHotwordEmbedding(std::string& hotwords) override;
  std::string GetLang() override { return inner_ ? inner_->GetLang() : std::string("zh-cn"); }
  int GetAsrSampleRate() override { return inner_ ? inner_->GetAsrSampleRate() : 16000; }
  void SetBatchSize(int batch_size) override { if (inner_) inner_->SetBatchSize(batch_size); }
  int GetBatchSize() override { return inner_ ? inner_->GetBatchSize() : opt_.max_batch; }

  MicroBatcherStats stats();

 private:
  struct Call;  // one Forward() invocation: its segments complete together
  struct Item {
    float* data;
    int len;
    int rows;
    const std::vector<std::vector<float>>* hw;  // the caller blocks until its segments are done, so its matrix outlives the batch
    Call* call;
    int index;  // position inside the call
    std::chrono::steady_clock::time_point t_enq;
  };
  void Run();
  static bool SameHotwords(const std::vector<std::vector<float>>& a, const std::vector<std::vector<float>>& b);

  BatchFn fn_;
  ParaformerB200* inner_ = nullptr;
  MicroBatcherOptions opt_;
  std::mutex mu_;
  std::condition_variable cv_work_;
  std::deque<Item> queue_;
  int64_t queued_rows_ = 0;
  bool stop_ = false;
  MicroBatcherStats stats_;
  std::thread worker_;
};

}  // namespace funasr_b200
