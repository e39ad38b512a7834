#include "multi_gpu.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace funasr_b200 {

double SegmentCost(int n_samples) {
  // c1*T + c2*T^2 per segment: T LFR frames, ~335 MFLOP per frame of dense work plus the
  // attention term 50 * 4 * T^2 * 512 (+ the decoder's cross attention at L ~ T/2)
  const double T = (double)b200pf_num_lfr_frames(n_samples);
  return T <= 0 ? 0.0 : 335e6 * T + (50.0 * 4.0 * 512.0 + 16.0 * 2.0 * 512.0) * T * T;
}

void PartitionSegments(const int* len, int n, int n_dev, std::vector<int>* assign) {
  assign->assign(n, 0);
  if (n_dev <= 1 || n <= 0) return;
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] > len[b]; });  // longest first
  std::vector<double> load(n_dev, 0.0);
  for (int i : order) {
    const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    (*assign)[i] = d;
    load[d] += SegmentCost(len[i]) + 1.0;  // +1: zero-cost (too short) segments are spread as well
  }
}

MultiGpuParaformer::MultiGpuParaformer(const std::vector<int>& devices, int max_rows, int max_segments)
    : devices_(devices), max_rows_(max_rows), max_segments_(max_segments) {
  if (devices_.empty()) devices_.push_back(0);
}

MultiGpuParaformer::~MultiGpuParaformer() {
  for (auto& w : workers_) {
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->stop = true;
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
  }
}

void MultiGpuParaformer::Loop(Worker* w) {
  std::unique_lock<std::mutex> lk(w->mu);
  for (;;) {
    w->cv.wait(lk, [&] { return w->stop || !w->jobs.empty(); });
    if (w->jobs.empty()) return;  // stop requested and drained
    std::function<void()> job = std::move(w->jobs.front());
    w->jobs.pop_front();
    lk.unlock();
    job();
    lk.lock();
  }
}

void MultiGpuParaformer::Post(int dev, std::function<void()> job) {
  Worker* w = workers_[dev].get();
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->jobs.push_back(std::move(job));
  }
  w->cv.notify_one();
}

bool MultiGpuParaformer::Init(const std::string& model_dir, std::string* err) {
  models_.clear();
  for (int d : devices_) {
    std::unique_ptr<ParaformerB200> m(new ParaformerB200(d, max_rows_, max_segments_));
    if (!m->Init(model_dir, err)) return false;   // weights are replicated: one resident copy per GPU
    models_.push_back(std::move(m));
  }
  for (size_t i = 0; i < models_.size(); ++i) {
    workers_.emplace_back(new Worker);
    workers_.back()->th = std::thread(&MultiGpuParaformer::Loop, workers_.back().get());
  }
  return true;
}

void MultiGpuParaformer::InitAsr(const std::string& am_model, const std::string& am_cmvn, const std::string& am_config,
                                 const std::string& token_file, int thread_num) {
  (void)am_cmvn; (void)am_config; (void)token_file; (void)thread_num;
  const size_t p = am_model.find_last_of('/');
  std::string err;
  if (!Init(p == std::string::npos ? std::string(".") : am_model.substr(0, p), &err)) {
    fprintf(stderr, "Error when load am b200pf model: %s\n", err.c_str());
    exit(-1);  // like the reference (paraformer.cpp:43-46)
  }
}

// The serial pass over a call's segments in the caller's order: the workers prepared each segment's text for both incoming
// detokeniser states, so this only follows the state chain (what a one-GPU handle's detokeniser would have gone through).
void MultiGpuParaformer::AssembleText(const std::vector<const SegmentRaw*>& by_index, std::vector<std::string>* results) {
  pf::host::Detokenizer* vocab = models_[0]->vocab();
  bool st = vocab->ended_on_english_word();
  for (size_t i = 0; i < by_index.size(); ++i) {
    const SegmentRaw* r = by_index[i];
    if (!r || !r->has_features) continue;   // "" and the state is left alone, as ParaformerB200::Decode does
    if (r->has_text) {
      (*results)[i] = r->text[st ? 1 : 0];
      st = r->ended[st ? 1 : 0];
    } else {
      vocab->set_ended_on_english_word(st);
      (*results)[i] = models_[0]->TextOf(*r);
      st = vocab->ended_on_english_word();
    }
  }
  vocab->set_ended_on_english_word(st);
}

std::vector<std::string> MultiGpuParaformer::Forward(float** din, int* len, bool input_finished,
                                                     const std::vector<std::vector<float>>& hw_emb, void* wfst_decoder, int batch_in) {
  std::vector<std::string> results(batch_in > 0 ? batch_in : 0);
  if (batch_in <= 0 || models_.empty()) return results;
  const int nd = (int)models_.size();
  if (nd == 1) return models_[0]->Forward(din, len, input_finished, hw_emb, wfst_decoder, batch_in);
  std::vector<int> assign;
  PartitionSegments(len, batch_in, nd, &assign);
  // Greedy path: the workers return token ids (SegmentRaw) and the text is built HERE, in the caller's order, through the first
  // model's detokeniser -- so that sharding changes no character of what a one-GPU handle returns (the reference's Vocab carries
  // a leading-space decision from one call to the next).  With an LM decoder handle the workers decode themselves.
  const bool central_text = wfst_decoder == nullptr;
  struct Share { std::vector<float*> ptr; std::vector<int> len, idx; std::vector<std::string> out; std::vector<SegmentRaw> raw; };
  std::vector<Share> share(nd);
  for (int i = 0; i < batch_in; ++i) {
    Share& s = share[assign[i]];
    s.ptr.push_back(din[i]); s.len.push_back(len[i]); s.idx.push_back(i);
  }
  std::mutex mu;
  std::condition_variable cv;
  int pending = 0;
  for (int d = 0; d < nd; ++d) if (!share[d].idx.empty()) ++pending;
  for (int d = 0; d < nd; ++d) {
    if (share[d].idx.empty()) continue;
    Share* s = &share[d];
    ParaformerB200* m = models_[d].get();
    Worker* w = workers_[d].get();
    Post(d, [=, &hw_emb, &mu, &cv, &pending]() {
      if (central_text) s->raw = m->ForwardRaw(s->ptr.data(), s->len.data(), (int)s->idx.size(), hw_emb);
      else s->out = m->Forward(s->ptr.data(), s->len.data(), input_finished, hw_emb, wfst_decoder, (int)s->idx.size());
      {
        std::lock_guard<std::mutex> wl(w->mu);
        w->segments += (long long)s->idx.size();
      }
      std::lock_guard<std::mutex> lk(mu);
      if (--pending == 0) cv.notify_all();
    });
  }
  {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return pending == 0; });
  }
  if (central_text) {
    std::vector<const SegmentRaw*> by_index(batch_in, nullptr);
    for (int d = 0; d < nd; ++d)
      for (size_t k = 0; k < share[d].idx.size() && k < share[d].raw.size(); ++k) by_index[share[d].idx[k]] = &share[d].raw[k];
    std::lock_guard<std::mutex> lk(text_mu_);
    AssembleText(by_index, &results);
    return results;
  }
  for (int d = 0; d < nd; ++d)
    for (size_t k = 0; k < share[d].idx.size(); ++k)
      if (k < share[d].out.size()) results[share[d].idx[k]] = std::move(share[d].out[k]);
  return results;
}

std::vector<std::string> MultiGpuParaformer::ForwardSegments16(const int16_t* const* seg, const int64_t* len, int n_seg,
                                                               const std::vector<std::vector<float>>& hw_emb) {
  std::vector<std::string> results(n_seg > 0 ? n_seg : 0);
  if (n_seg <= 0 || models_.empty()) return results;
  const int nd = (int)models_.size();
  if (nd == 1) return models_[0]->ForwardSegments16(seg, len, n_seg, hw_emb);
  static const bool trace = getenv("B200PF_HOST_TRACE") != nullptr;   // per-phase wall times of every call on stderr
  using clk = std::chrono::steady_clock;
  const clk::time_point t0 = clk::now();
  std::vector<double> worker_ms(nd, 0.0);
  std::vector<int> len32(n_seg), assign;
  for (int i = 0; i < n_seg; ++i) len32[i] = (int)len[i];
  PartitionSegments(len32.data(), n_seg, nd, &assign);
  struct Share { std::vector<const int16_t*> ptr; std::vector<int64_t> len; std::vector<int> idx; std::vector<SegmentRaw> raw; };
  std::vector<Share> share(nd);
  for (int i = 0; i < n_seg; ++i) {
    Share& s = share[assign[i]];
    s.ptr.push_back(seg[i]); s.len.push_back(len[i]); s.idx.push_back(i);
  }
  std::mutex mu;
  std::condition_variable cv;
  int pending = 0;
  for (int d = 0; d < nd; ++d) if (!share[d].idx.empty()) ++pending;
  for (int d = 0; d < nd; ++d) {
    if (share[d].idx.empty()) continue;
    Share* s = &share[d];
    ParaformerB200* m = models_[d].get();
    Worker* w = workers_[d].get();
    Post(d, [=, &hw_emb, &mu, &cv, &pending, &worker_ms]() {
      const clk::time_point w0 = clk::now();
      s->raw = m->ForwardSegments16Raw(s->ptr.data(), s->len.data(), (int)s->idx.size(), hw_emb);
      worker_ms[d] = std::chrono::duration<double, std::milli>(clk::now() - w0).count();
      {
        std::lock_guard<std::mutex> wl(w->mu);
        w->segments += (long long)s->idx.size();
      }
      std::lock_guard<std::mutex> lk(mu);
      if (--pending == 0) cv.notify_all();
    });
  }
  {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return pending == 0; });
  }
  const clk::time_point t1 = clk::now();
  {   // text in the caller's order through one detokeniser (see Forward above)
    std::vector<const SegmentRaw*> by_index(n_seg, nullptr);
    for (int d = 0; d < nd; ++d)
      for (size_t k = 0; k < share[d].idx.size() && k < share[d].raw.size(); ++k) by_index[share[d].idx[k]] = &share[d].raw[k];
    std::lock_guard<std::mutex> lk(text_mu_);
    AssembleText(by_index, &results);
  }
  if (trace) {
    const clk::time_point t2 = clk::now();
    fprintf(stderr, "[b200pf pool] %d segments on %d GPUs: dispatch + workers %.2f ms (per GPU:", n_seg, nd,
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    for (int d = 0; d < nd; ++d) fprintf(stderr, " %.2f", worker_ms[d]);
    fprintf(stderr, "), text %.2f ms\n", std::chrono::duration<double, std::milli>(t2 - t1).count());
  }
  return results;
}

std::string MultiGpuParaformer::Forward(float* din, int len, bool input_finished, const std::vector<std::vector<float>>& hw_emb,
                                        void* wfst_decoder) {
  float* one[1] = {din};
  int l[1] = {len};
  std::vector<std::string> r = Forward(one, l, input_finished, hw_emb, wfst_decoder, 1);
  return r.empty() ? std::string() : r[0];
}

void MultiGpuParaformer::InitHwCompiler(const std::string& hw_model, int thread_num) {
  for (auto& m : models_) m->InitHwCompiler(hw_model, thread_num);
}
void MultiGpuParaformer::InitSegDict(const std::string& seg_dict_model) {
  for (auto& m : models_) m->InitSegDict(seg_dict_model);
}
std::vector<std::vector<float>> MultiGpuParaformer::CompileHotwordEmbedding(std::string& hotwords) {
  if (models_.empty()) return std::vector<std::vector<float>>();
  return models_[0]->CompileHotwordEmbedding(hotwords);
}
std::string MultiGpuParaformer::GetLang() { return models_.empty() ? std::string("zh-cn") : models_[0]->GetLang(); }
int MultiGpuParaformer::GetAsrSampleRate() { return models_.empty() ? 16000 : models_[0]->GetAsrSampleRate(); }
void MultiGpuParaformer::SetBatchSize(int batch_size) { for (auto& m : models_) m->SetBatchSize(batch_size); }
int MultiGpuParaformer::GetBatchSize() { return models_.empty() ? 1 : models_[0]->GetBatchSize(); }

std::vector<long long> MultiGpuParaformer::segments_per_device() {
  std::vector<long long> out;
  for (auto& w : workers_) {
    std::lock_guard<std::mutex> lk(w->mu);
    out.push_back(w->segments);
  }
  return out;
}

}  // namespace funasr_b200
