#include "micro_batcher.h"

#include <algorithm>
#include <cstring>

namespace funasr_b200 {

struct MicroBatcher::Call {
  std::mutex m;
  std::condition_variable cv;
  int remaining = 0;
  std::vector<std::string> results;
};

MicroBatcher::MicroBatcher(BatchFn fn, const MicroBatcherOptions& opt) : fn_(std::move(fn)), opt_(opt) {
  if (opt_.max_batch < 1) opt_.max_batch = 1;
  if (opt_.max_rows < 1) opt_.max_rows = 1;
  if (opt_.max_wait_us < 0) opt_.max_wait_us = 0;
  worker_ = std::thread(&MicroBatcher::Run, this);
}

MicroBatcher::MicroBatcher(ParaformerB200* inner, const MicroBatcherOptions& opt)
    : MicroBatcher(
          [inner](float** din, int* len, int n, const std::vector<std::vector<float>>& hw) {
            return inner->Forward(din, len, true, hw, nullptr, n);
          },
          opt) {
  inner_ = inner;
}

MicroBatcher::~MicroBatcher() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  if (worker_.joinable()) worker_.join();  // drains what is still queued
}

void MicroBatcher::InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config,
                           const std::string& token_file, int thread_num) {
  if (inner_) inner_->InitAsr(am_model, am_cmvn, am_config, token_file, thread_num);
}

std::vector<std::vector<float>> MicroBatcher::CompileHotwordEmbedding(std::string& hotwords) {
  if (inner_) return inner_->CompileHotwordEmbedding(hotwords);
  return std::vector<std::vector<float>>(1, std::vector<float>(512, 0.0f));
}

bool MicroBatcher::SameHotwords(const std::vector<std::vector<float>>& a, const std::vector<std::vector<float>>& b) {
  if (&a == &b) return true;
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i) {
    if (a[i].size() != b[i].size()) return false;
    if (!a[i].empty() && memcmp(a[i].data(), b[i].data(), a[i].size() * sizeof(float)) != 0) return false;
  }
  return true;
}

std::vector<std::string> MicroBatcher::Forward(float** din, int* len, bool input_finished, const std::vector<std::vector<float>>& hw_emb,
                                               void* wfst_decoder, int batch_in) {
  (void)input_finished; (void)wfst_decoder;
  if (batch_in <= 0) return std::vector<std::string>();
  Call call;
  call.remaining = batch_in;
  call.results.resize(batch_in);
  const auto now = std::chrono::steady_clock::now();
  {
    std::lock_guard<std::mutex> lk(mu_);
    if (stop_) return call.results;
    for (int i = 0; i < batch_in; ++i) {
      Item it;
      it.data = din[i];
      it.len = len[i];
      const int T = b200pf_num_lfr_frames(len[i]);
      it.rows = T > 0 ? T + 1 : 0;
      it.hw = &hw_emb;  // the caller blocks below, so its matrix outlives the batch
      it.call = &call;
      it.index = i;
      it.t_enq = now;
      queue_.push_back(it);
      queued_rows_ += it.rows;
    }
  }
  cv_work_.notify_all();
  std::unique_lock<std::mutex> lk(call.m);
  call.cv.wait(lk, [&] { return call.remaining == 0; });
  return call.results;
}

std::string MicroBatcher::Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb, void* wfst_decoder) {
  float* one[1] = {din};
  int l[1] = {len};
  std::vector<std::string> r = Forward(one, l, input_finished, hw_emb, wfst_decoder, 1);
  return r.empty() ? std::string() : r[0];
}

void MicroBatcher::Run() {
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_work_.wait(lk, [&] { return stop_ || !queue_.empty(); });
    if (queue_.empty()) {
      if (stop_) return;
      continue;
    }
    // close the batch at the oldest segment's deadline, or as soon as enough work is waiting
    const auto deadline = queue_.front().t_enq + std::chrono::microseconds(opt_.max_wait_us);
    bool by_size = false;
    while (!stop_) {
      if ((int)queue_.size() >= opt_.max_batch || queued_rows_ >= opt_.max_rows) { by_size = true; break; }
      if (cv_work_.wait_until(lk, deadline) == std::cv_status::timeout) break;
      if (std::chrono::steady_clock::now() >= deadline) break;
    }
    // oldest first; only segments that share the oldest one's hotword matrix travel together
    std::vector<Item> batch;
    const std::vector<std::vector<float>>* key = queue_.front().hw;
    int64_t rows = 0;
    for (auto it = queue_.begin(); it != queue_.end();) {
      const bool fits = batch.empty() || ((int)batch.size() < opt_.max_batch && rows + it->rows <= opt_.max_rows);
      if (fits && SameHotwords(*it->hw, *key)) {
        batch.push_back(*it);
        rows += it->rows;
        queued_rows_ -= it->rows;
        it = queue_.erase(it);
      } else {
        ++it;
      }
    }
    const auto t_start = std::chrono::steady_clock::now();
    lk.unlock();

    // ascending length, like the reference sorts VAD segments before batching (audio.cpp:1233-1238)
    std::stable_sort(batch.begin(), batch.end(), [](const Item& a, const Item& b) { return a.len < b.len; });
    std::vector<float*> ptrs(batch.size());
    std::vector<int> lens(batch.size());
    for (size_t i = 0; i < batch.size(); ++i) { ptrs[i] = batch[i].data; lens[i] = batch[i].len; }
    std::vector<std::string> out;
    try {
      out = fn_(ptrs.data(), lens.data(), (int)batch.size(), *key);
    } catch (...) {
      out.clear();  // like the reference: a failing inference yields "" for its segments (paraformer.cpp:582-587)
    }
    out.resize(batch.size());
    double wait_sum = 0.0, wait_max = 0.0;
    for (size_t i = 0; i < batch.size(); ++i) {
      const double w = std::chrono::duration<double, std::micro>(t_start - batch[i].t_enq).count();
      wait_sum += w;
      wait_max = std::max(wait_max, w);
      Call* c = batch[i].call;
      bool done;
      {
        std::lock_guard<std::mutex> cl(c->m);
        c->results[batch[i].index] = std::move(out[i]);
        done = --c->remaining == 0;
        // notify while holding the lock: the Call lives on the waiter's stack and may be destroyed as soon as
        // the waiter observes remaining == 0
        if (done) c->cv.notify_all();
      }
    }
    lk.lock();
    stats_.segments += (int64_t)batch.size();
    stats_.batches += 1;
    if (by_size) stats_.closed_by_size += 1; else stats_.closed_by_deadline += 1;
    stats_.max_batch_seen = std::max<int64_t>(stats_.max_batch_seen, (int64_t)batch.size());
    stats_.wait_us_sum += wait_sum;
    stats_.wait_us_max = std::max(stats_.wait_us_max, wait_max);
  }
}

MicroBatcherStats MicroBatcher::stats() {
  std::lock_guard<std::mutex> lk(mu_);
  return stats_;
}

}  // namespace funasr_b200
