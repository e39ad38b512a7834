#include "funasrruntime_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <numeric>

#include "micro_batcher.h"
#include "multi_gpu.h"
#include "paraformer_b200.h"
#include "punc_b200.h"
#include "vad_segmenter.h"

namespace {

struct RecogResult {  // FUNASR_RECOG_RESULT, onnxruntime/src/commonfunc.h:8-15
  std::string msg, stamp, stamp_sents, tpass_msg;
  float snippet_time = 0.f;
};

struct OfflineHandle {  // stands where funasr::OfflineStream does; owns only the acoustic model
  std::unique_ptr<funasr_b200::ParaformerB200> asr;        // single GPU
  std::unique_ptr<funasr_b200::MultiGpuParaformer> pool;   // "devices" = "0,1,...": one engine per GPU, independent queues
  // "micro-batch-us" = deadline: concurrent FunOfflineInfer* calls (the servers' decoder-thread-num threads, one request each)
  // are merged into batched forwards instead of queueing up as latency-bound small ones
  std::unique_ptr<funasr_b200::MicroBatcher> batcher;
  // "vad-dir": FSMN-VAD scores on the GPU + the E2E state machine on the host cut a recording into speech segments the way
  // Audio::CutSplit does (audio.cpp:1172-1226); without it FunOfflineInferBuffer falls back to hard cuts at vad_max_len
  b200pf_vad* vad = nullptr;
  std::mutex vad_mu;   // one VAD workspace per handle
  float vad_thres = 0.6f;
  // "punc-dir": CT-Transformer punctuation on the GPU after stitching (the UsePunc() branch, funasrruntime.cpp:317-320)
  std::unique_ptr<funasr_b200::CTTransformerB200> punc;
  ~OfflineHandle() { if (vad) b200pf_vad_destroy(vad); }
  funasr_b200::Model* model() { return pool ? (funasr_b200::Model*)pool.get() : (funasr_b200::Model*)asr.get(); }
  funasr_b200::ParaformerB200* first() { return pool ? pool->model(0) : asr.get(); }
};

int ToInt(const std::map<std::string, std::string>& m, const char* key, int dflt) {
  auto it = m.find(key);
  return it == m.end() ? dflt : atoi(it->second.c_str());
}

// Audio::FetchDynamic (audio.cpp:1052-1108) over an ascending-length queue: <= batch_size items,
// max_len * (n + 1) <= 300 s, a >= 60 s item travels alone.
std::vector<std::vector<int>> FormBatches(const std::vector<long long>& len_sorted, int batch_size) {
  const long long max_acc = 300LL * 1000 * 16, max_sent = 60LL * 1000 * 16;
  std::vector<std::vector<int>> out;
  size_t head = 0;
  while (head < len_sorted.size()) {
    std::vector<int> cur;
    long long max_len = 0;
    const size_t room = std::min<size_t>((size_t)std::max(batch_size, 1), len_sorted.size() - head);
    for (size_t k = 0; k < room; ++k) {
      const long long L = len_sorted[head];
      if (L >= max_sent) {
        if (cur.empty()) { cur.push_back((int)head); ++head; }
        break;
      }
      max_len = std::max(max_len, L);
      if (max_len * (long long)(cur.size() + 1) > max_acc) break;
      cur.push_back((int)head);
      ++head;
    }
    if (cur.empty()) break;
    out.push_back(cur);
  }
  return out;
}

// What FunOfflineInferBuffer does after stitching (funasrruntime.cpp:317-335): punctuation, then the per-sentence stamps.  (ITN
// and its TimestampSmooth stay on the reference host path.)
void FinishResult(OfflineHandle* h, RecogResult* res) {
  if (h->punc) res->msg = h->punc->AddPunc(res->msg.c_str(), h->model()->GetLang());
  if (!res->stamp.empty()) res->stamp_sents = pf::host::SentenceStamps(res->msg, res->stamp);
}

RecogResult* RunSegments(OfflineHandle* h, const short* pcm, long long n_samples, std::vector<long long> seg_b,
                         std::vector<long long> seg_e, const std::vector<std::vector<float>>& hw_emb) {
  funasr_b200::Model* asr = h->model();
  RecogResult* res = new RecogResult;
  res->snippet_time = (float)n_samples / asr->GetAsrSampleRate();
  if (res->snippet_time == 0) return res;
  const int n = (int)seg_b.size();
  if (n == 0) return res;   // a recording the VAD found no speech in
  // ascending length order + permutation (Audio::CutSplit, audio.cpp:1228-1238); std::sort like the reference
  std::vector<int> index(n);
  std::iota(index.begin(), index.end(), 0);
  std::sort(index.begin(), index.end(), [&](int a, int b) { return seg_e[a] - seg_b[a] < seg_e[b] - seg_b[b]; });
  std::vector<long long> len_sorted(n);
  for (int k = 0; k < n; ++k) len_sorted[k] = seg_e[index[k]] - seg_b[index[k]];
  std::vector<std::string> msgs(n);
  std::vector<float> starts(n);
  // The reference forms batches with Audio::FetchDynamic's padded-audio rule (max_len * (n + 1) <= 300 s, FormBatches above),
  // which exists because its GPU path PADS every item to the longest.  The B200 engine packs variable-length rows and a
  // segment's result does not depend on its batch, so the only cap that matters is the engine's own capacity:
  // ParaformerB200 splits by max_rows / max_segments itself.  B200PF_FETCH_DYNAMIC=1 restores the reference's rule.
  static const bool fetch_dynamic = getenv("B200PF_FETCH_DYNAMIC") != nullptr;
  std::vector<std::vector<int>> batches;
  if (fetch_dynamic) {
    batches = FormBatches(len_sorted, asr->GetBatchSize());
  } else {
    batches.emplace_back(n);
    std::iota(batches[0].begin(), batches[0].end(), 0);
  }
  for (const auto& batch : batches) {
    // segments are handed over by pointer (they may overlap or be out of order in the source buffer): no host-side gather
    std::vector<const int16_t*> ptrs(batch.size());
    std::vector<int64_t> lens(batch.size());
    for (size_t k = 0; k < batch.size(); ++k) {
      const int s = index[batch[k]];
      ptrs[k] = (const int16_t*)pcm + seg_b[s];
      lens[k] = seg_e[s] - seg_b[s];
    }
    std::vector<std::string> out;
    if (h->batcher) {
      // through the micro-batcher: it speaks the reference's float form (float = int16 / 32768, Audio::LoadPcmwav)
      std::vector<std::vector<float>> fl(batch.size());
      std::vector<float*> fp(batch.size());
      std::vector<int> li(batch.size());
      for (size_t k = 0; k < batch.size(); ++k) {
        fl[k].resize((size_t)lens[k]);
        for (int64_t i = 0; i < lens[k]; ++i) fl[k][i] = (float)ptrs[k][i] / 32768.0f;
        fp[k] = fl[k].data();
        li[k] = (int)lens[k];
      }
      out = h->batcher->Forward(fp.data(), li.data(), true, hw_emb, nullptr, (int)batch.size());
    } else {
      out = h->pool ? h->pool->ForwardSegments16(ptrs.data(), lens.data(), (int)batch.size(), hw_emb)
                    : h->asr->ForwardSegments16(ptrs.data(), lens.data(), (int)batch.size(), hw_emb);
    }
    for (size_t k = 0; k < batch.size(); ++k) {
      const int s = index[batch[k]];
      msgs[s] = out[k];
      starts[s] = (float)seg_b[s] / asr->GetAsrSampleRate();
    }
  }
  pf::host::StitchSegments(msgs, starts, asr->GetLang(), &res->msg, &res->stamp);
  FinishResult(h, res);
  return res;
}

}  // namespace

namespace {
// model_conf.speech_noise_thres of the VAD's config.yaml (fsmn-vad.cpp:38); a flat "key: value" scan is all this needs
float ReadVadThreshold(const std::string& vad_dir, float dflt) {
  std::ifstream f(vad_dir + "/config.yaml");
  std::string line;
  while (std::getline(f, line)) {
    const size_t k = line.find("speech_noise_thres:");
    if (k == std::string::npos || line.find('#') < k) continue;
    const float v = (float)atof(line.c_str() + k + strlen("speech_noise_thres:"));
    if (v > 0.f) return v;
  }
  return dflt;
}

bool InitPunc(OfflineHandle* h, const std::map<std::string, std::string>& model_path, int device) {
  auto pd = model_path.find("punc-dir");
  if (pd == model_path.end() || pd->second.empty()) return true;
  h->punc.reset(new funasr_b200::CTTransformerB200(device, ToInt(model_path, "punc-max-tokens", 0)));
  std::string err;
  if (!h->punc->Init(pd->second, &err)) {
    fprintf(stderr, "FunOfflineInit: punc-dir %s: %s\n", pd->second.c_str(), err.c_str());
    return false;
  }
  return true;
}

bool InitVad(OfflineHandle* h, const std::map<std::string, std::string>& model_path, int device) {
  auto vd = model_path.find("vad-dir");
  if (vd == model_path.end() || vd->second.empty()) return true;
  const int rc = b200pf_vad_create(vd->second.c_str(), device, ToInt(model_path, "vad-max-frames", 0), &h->vad);
  if (rc != 0) {
    fprintf(stderr, "FunOfflineInit: vad-dir %s: %s\n", vd->second.c_str(), b200pf_last_error());
    return false;
  }
  h->vad_thres = ReadVadThreshold(vd->second, 0.6f);
  auto th = model_path.find("vad-speech-noise-thres");
  if (th != model_path.end() && atof(th->second.c_str()) > 0) h->vad_thres = (float)atof(th->second.c_str());
  return true;
}
}  // namespace

FUNASR_HANDLE FunOfflineInit(std::map<std::string, std::string>& model_path, int thread_num, bool use_gpu, int batch_size) {
  (void)thread_num; (void)use_gpu;
  auto it = model_path.find("model-dir");
  if (it == model_path.end()) { fprintf(stderr, "FunOfflineInit: model-dir missing\n"); return nullptr; }
  std::unique_ptr<OfflineHandle> h(new OfflineHandle);
  std::string err;
  auto dv = model_path.find("devices");   // extra key: "0,1,2,3" -> one engine per listed GPU behind this one handle
  std::vector<int> devices;
  if (dv != model_path.end()) {
    size_t p = 0;
    while (p < dv->second.size()) {
      size_t q = dv->second.find(',', p);
      if (q == std::string::npos) q = dv->second.size();
      if (q > p) devices.push_back(atoi(dv->second.substr(p, q - p).c_str()));
      p = q + 1;
    }
  }
  if (devices.size() > 1) {
    h->pool.reset(new funasr_b200::MultiGpuParaformer(devices, ToInt(model_path, "max-rows", 0), ToInt(model_path, "max-segments", 0)));
    if (!h->pool->Init(it->second, &err)) {
      fprintf(stderr, "FunOfflineInit: %s\n", err.c_str());
      return nullptr;
    }
    h->pool->SetBatchSize(batch_size);
    if (!InitVad(h.get(), model_path, devices[0]) || !InitPunc(h.get(), model_path, devices[0])) return nullptr;
    if (ToInt(model_path, "micro-batch-us", 0) > 0) {
      funasr_b200::MicroBatcherOptions o;
      o.max_wait_us = ToInt(model_path, "micro-batch-us", 0);
      funasr_b200::MultiGpuParaformer* pool = h->pool.get();
      h->batcher.reset(new funasr_b200::MicroBatcher(
          [pool](float** din, int* len, int n, const std::vector<std::vector<float>>& hw) { return pool->Forward(din, len, true, hw, nullptr, n); }, o));
    }
    return h.release();
  }
  h->asr.reset(new funasr_b200::ParaformerB200(devices.size() == 1 ? devices[0] : ToInt(model_path, "device", 0),
                                               ToInt(model_path, "max-rows", 0), ToInt(model_path, "max-segments", 0)));
  if (!h->asr->Init(it->second, &err)) {
    fprintf(stderr, "FunOfflineInit: %s\n", err.c_str());
    return nullptr;
  }
  h->asr->SetBatchSize(batch_size);
  const int dev0 = devices.size() == 1 ? devices[0] : ToInt(model_path, "device", 0);
  if (!InitVad(h.get(), model_path, dev0) || !InitPunc(h.get(), model_path, dev0)) return nullptr;
  if (ToInt(model_path, "micro-batch-us", 0) > 0) {
    funasr_b200::MicroBatcherOptions o;
    o.max_wait_us = ToInt(model_path, "micro-batch-us", 0);
    h->batcher.reset(new funasr_b200::MicroBatcher(h->asr.get(), o));
  }
  return h.release();
}

void FunOfflineReset(FUNASR_HANDLE, FUNASR_DEC_HANDLE) {}

funasr_b200::ParaformerB200* FunOfflineModelB200(FUNASR_HANDLE handle) { return handle ? ((OfflineHandle*)handle)->first() : nullptr; }
funasr_b200::Model* FunOfflineModel(FUNASR_HANDLE handle) { return handle ? ((OfflineHandle*)handle)->model() : nullptr; }
funasr_b200::MultiGpuParaformer* FunOfflinePoolB200(FUNASR_HANDLE handle) { return handle ? ((OfflineHandle*)handle)->pool.get() : nullptr; }

FUNASR_RESULT FunOfflineInferSegmentsB200(FUNASR_HANDLE handle, const short* pcm, long long n_samples, const long long* seg_begin,
                                          const long long* seg_end, int n_seg, const std::vector<std::vector<float>>& hw_emb) {
  OfflineHandle* h = (OfflineHandle*)handle;
  if (!h || (!pcm && n_samples > 0)) return nullptr;
  std::vector<long long> b(seg_begin, seg_begin + n_seg), e(seg_end, seg_end + n_seg);
  for (int i = 0; i < n_seg; ++i)
    if (b[i] < 0 || e[i] < b[i] || e[i] > n_samples) return nullptr;
  return RunSegments(h, pcm, n_samples, b, e, hw_emb);
}

namespace {
bool VadCut(OfflineHandle* h, const short* pcm, long long n, int vad_tail_sil, int vad_max_len, std::vector<std::pair<int, int>>* segs) {
  const long long n_frames = n >= 400 ? 1 + (n - 400) / 160 : 0;
  std::vector<float> sil((size_t)std::max<long long>(1, n_frames));
  int64_t off[2] = {0, (int64_t)n};
  int32_t frame_off[2] = {0, 0};
  {
    std::lock_guard<std::mutex> lk(h->vad_mu);
    if (b200pf_vad_scores_s16(h->vad, (const int16_t*)pcm, off, 1, sil.data(), (int64_t)sil.size(), frame_off, nullptr, nullptr) != 0) {
      fprintf(stderr, "FunOfflineInferBuffer: VAD: %s\n", b200pf_last_error());
      return false;
    }
  }
  pf::host::VadOptions vo;
  vo.max_end_silence_ms = vad_tail_sil;
  vo.max_single_segment_ms = vad_max_len;
  vo.speech_noise_thres = h->vad_thres;
  *segs = pf::host::SegmentVad(sil.data(), frame_off[1] - frame_off[0], vo);
  return true;
}
}  // namespace

// The VAD cut alone (the job of Audio::CutSplit, audio.cpp:1172-1226): [start_ms, end_ms) pairs for one recording.
int FunOfflineVadSegmentsB200(FUNASR_HANDLE handle, const short* pcm, long long n_samples, int vad_tail_sil, int vad_max_len,
                              std::vector<std::pair<int, int>>* out) {
  OfflineHandle* h = (OfflineHandle*)handle;
  if (!h || !h->vad || !out) return -1;
  return VadCut(h, pcm, n_samples, vad_tail_sil, vad_max_len, out) ? 0 : -2;
}

FUNASR_RESULT FunOfflineInferBuffer(FUNASR_HANDLE handle, const char* sz_buf, int n_len, FUNASR_MODE, QM_CALLBACK,
                                    const std::vector<std::vector<float>>& hw_emb, int sampling_rate, std::string wav_format, bool,
                                    int vad_tail_sil, int vad_max_len, FUNASR_DEC_HANDLE, std::string, bool) {
  OfflineHandle* h = (OfflineHandle*)handle;
  if (!h) return nullptr;  // funasrruntime.cpp:216-217
  if (!(wav_format == "pcm" || wav_format == "PCM")) {
    fprintf(stderr, "FunOfflineInferBuffer: only raw PCM is decoded here (ffmpeg stays on the reference host path)\n");
    return nullptr;
  }
  if (sampling_rate != h->model()->GetAsrSampleRate()) {
    fprintf(stderr, "FunOfflineInferBuffer: resampling stays on the reference host path\n");
    return nullptr;
  }
  const long long n = n_len / 2;  // Audio::LoadPcmwav: little-endian int16 (audio.cpp:795-805)
  std::vector<short> pcm((size_t)n);
  const unsigned char* bytes = (const unsigned char*)sz_buf;
  for (long long i = 0; i < n; ++i) pcm[i] = (short)((bytes[2 * i + 1] << 8) | bytes[2 * i]);
  std::vector<long long> b, e;
  if (h->vad) {
    // UseVad() branch of the reference (funasrruntime.cpp:243-245): scores for the whole recording in one GPU pass, then the
    // E2E state machine; segment bounds are milliseconds * (sample_rate / 1000) like Audio::CutSplit (audio.cpp:1214-1215)
    if (n == 0) return RunSegments(h, pcm.data(), 0, b, e, hw_emb);
    std::vector<std::pair<int, int>> segs;
    if (!VadCut(h, pcm.data(), n, vad_tail_sil, vad_max_len, &segs)) return nullptr;
    const long long per_ms = h->model()->GetAsrSampleRate() / 1000;
    for (const auto& sg : segs) {
      const long long sb = std::min<long long>(n, sg.first * per_ms), se = std::min<long long>(n, sg.second * per_ms);
      if (se > sb) { b.push_back(sb); e.push_back(se); }
    }
    return RunSegments(h, pcm.data(), n, b, e, hw_emb);
  }
  // no VAD model: one segment, hard-cut at vad_max_len so that a segment always fits the engine
  const long long cut = std::max(1, vad_max_len) * 16LL;
  for (long long s = 0; s < n; s += cut) { b.push_back(s); e.push_back(std::min(n, s + cut)); }
  if (n == 0) { b.push_back(0); e.push_back(0); }
  return RunSegments(h, pcm.data(), n, b, e, hw_emb);
}

namespace {
// Reads a .wav (16-bit mono PCM) or raw .pcm file; returns false when it cannot be used.  *payload points into *data.
bool LoadAudioFile(const char* path, int default_rate, std::vector<char>* data, const char** payload, size_t* n_bytes, int* rate);
}  // namespace

FUNASR_HANDLE FunASRInit(std::map<std::string, std::string>& model_path, int thread_num, ASR_TYPE type) {
  (void)thread_num;
  if (type != ASR_OFFLINE) { fprintf(stderr, "FunASRInit: only the offline Paraformer is implemented here\n"); return nullptr; }
  auto it = model_path.find("model-dir");
  if (it == model_path.end()) { fprintf(stderr, "FunASRInit: model-dir missing\n"); return nullptr; }
  std::unique_ptr<funasr_b200::ParaformerB200> m(new funasr_b200::ParaformerB200(ToInt(model_path, "device", 0), ToInt(model_path, "max-rows", 0),
                                                                                  ToInt(model_path, "max-segments", 0)));
  std::string err;
  if (!m->Init(it->second, &err)) { fprintf(stderr, "FunASRInit: %s\n", err.c_str()); return nullptr; }
  return (funasr_b200::Model*)m.release();   // like the reference, the handle IS the model (funasrruntime.cpp:13-17)
}
void FunASRReset(FUNASR_HANDLE, FUNASR_DEC_HANDLE) {}
void FunASRUninit(FUNASR_HANDLE handle) { delete (funasr_b200::Model*)handle; }

FUNASR_RESULT FunASRInferBuffer(FUNASR_HANDLE handle, const char* sz_buf, int n_len, FUNASR_MODE, QM_CALLBACK, bool input_finished,
                                int sampling_rate, std::string wav_format) {
  funasr_b200::Model* m = (funasr_b200::Model*)handle;
  if (!m) return nullptr;  // funasrruntime.cpp:59-61
  if (!(wav_format == "pcm" || wav_format == "PCM") || sampling_rate != m->GetAsrSampleRate()) {
    fprintf(stderr, "FunASRInferBuffer: only raw PCM at the model's rate is decoded here (ffmpeg / resampling stay on the reference host path)\n");
    return nullptr;
  }
  const int n = n_len / 2;
  RecogResult* res = new RecogResult;
  res->snippet_time = (float)n / m->GetAsrSampleRate();
  if (res->snippet_time == 0) return res;
  std::vector<float> pcm((size_t)n);
  const unsigned char* bytes = (const unsigned char*)sz_buf;
  for (int i = 0; i < n; ++i) pcm[i] = (float)(short)((bytes[2 * i + 1] << 8) | bytes[2 * i]) / 32768.0f;  // Audio::LoadPcmwav
  res->msg += m->Forward(pcm.data(), n, input_finished);   // Audio::Fetch yields the whole audio once (no VAD on this API)
  if (funasr_b200::ParaformerB200::last_failed_segments() > 0) {
    // the recording did not fit the engine (this API decodes it as ONE segment: max-rows LFR frames, 60 ms each): a real error,
    // not an empty transcript.  FunOfflineInferBuffer cuts long recordings (VAD or vad_max_len) and has no such limit.
    fprintf(stderr, "FunASRInferBuffer: %.1f s of audio exceed the engine's capacity; raise \"max-rows\" or use FunOfflineInferBuffer\n", res->snippet_time);
    delete res;
    return nullptr;
  }
  return res;
}

FUNASR_RESULT FunASRInfer(FUNASR_HANDLE handle, const char* sz_filename, FUNASR_MODE mode, QM_CALLBACK cb, int sampling_rate) {
  if (!handle || !sz_filename) return nullptr;
  std::vector<char> data;
  const char* payload = nullptr;
  size_t n_bytes = 0;
  int rate = sampling_rate;
  if (!LoadAudioFile(sz_filename, sampling_rate, &data, &payload, &n_bytes, &rate)) return nullptr;
  return FunASRInferBuffer(handle, payload, (int)n_bytes, mode, cb, true, rate, "pcm");
}

FUNASR_RESULT FunOfflineInfer(FUNASR_HANDLE handle, const char* sz_filename, FUNASR_MODE mode, QM_CALLBACK cb,
                              const std::vector<std::vector<float>>& hw_emb, int sampling_rate, bool itn, int vad_tail_sil,
                              int vad_max_len, FUNASR_DEC_HANDLE dec_handle) {
  if (!handle || !sz_filename) return nullptr;
  std::vector<char> data;
  const char* payload = nullptr;
  size_t n_bytes = 0;
  int rate = sampling_rate;
  if (!LoadAudioFile(sz_filename, sampling_rate, &data, &payload, &n_bytes, &rate)) return nullptr;
  return FunOfflineInferBuffer(handle, payload, (int)n_bytes, mode, cb, hw_emb, rate, "pcm", itn, vad_tail_sil, vad_max_len, dec_handle);
}

namespace {
bool LoadAudioFile(const char* path, int default_rate, std::vector<char>* data_out, const char** payload, size_t* n_bytes, int* rate_out) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  data_out->assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<char>& data = *data_out;
  size_t off = 0;
  int rate = default_rate;
  if (data.size() >= 44 && memcmp(data.data(), "RIFF", 4) == 0 && memcmp(data.data() + 8, "WAVE", 4) == 0) {
    // walk the chunks: "fmt " gives the rate, "data" the payload (Audio::LoadWav, audio.cpp:622-737)
    size_t p = 12;
    off = data.size();
    size_t len = 0;
    while (p + 8 <= data.size()) {
      unsigned int sz;
      memcpy(&sz, data.data() + p + 4, 4);
      if (memcmp(data.data() + p, "fmt ", 4) == 0 && p + 8 + 16 <= data.size()) {
        unsigned short fmt, ch, bits;
        memcpy(&fmt, data.data() + p + 8, 2); memcpy(&ch, data.data() + p + 10, 2);
        memcpy(&rate, data.data() + p + 12, 4); memcpy(&bits, data.data() + p + 22, 2);
        if (ch != 1 || bits != 16) { fprintf(stderr, "only 16-bit mono wav is supported here\n"); return false; }
      } else if (memcmp(data.data() + p, "data", 4) == 0) {
        off = p + 8;
        len = std::min<size_t>(sz, data.size() - off);
        break;
      }
      p += 8 + sz + (sz & 1);
    }
    if (off >= data.size() && len == 0) return false;
    *payload = data.data() + off; *n_bytes = len; *rate_out = rate;
    return true;
  }
  *payload = data.data(); *n_bytes = data.size(); *rate_out = rate;
  return true;
}
}  // namespace

namespace {
// ---- 2-pass stream: the OFFLINE (correction) leg of FunTpassInferBuffer (funasrruntime.cpp:492-639) -----------------------------
// TpassStream (tpass-stream.h) owns the offline acoustic model, the VAD and the realtime punctuation model; the streaming acoustic
// model (ParaformerOnline) is not part of this path and is not built here: the online partial results (`msg`) stay empty, the
// corrected `tpass_msg` / `stamp` / `stamp_sents` of every closed VAD segment are produced exactly like the reference's FetchTpass
// loop produces them.
struct TpassHandle {
  OfflineHandle* offline = nullptr;                                   // acoustic model (+ micro-batcher) and the VAD network
  std::unique_ptr<funasr_b200::CTTransformerOnlineB200> punc_online;  // TpassStream::punc_online_handle (may be absent)
  ~TpassHandle() { delete offline; }
};

// TpassOnlineStream (tpass-online-stream.h): the per-connection state -- here the audio not yet consumed and the VAD state machine.
struct TpassOnlineHandle {
  TpassHandle* parent;
  pf::host::StreamingVad vad;
  std::vector<short> pcm;          // samples [base, base + pcm.size()) of the stream
  long long base = 0;              // global index of pcm[0]
  long long total = 0;             // samples received since the stream started
  int scored = 0;                  // 10 ms frames whose silence probability is final and has been pushed into `vad`
  explicit TpassOnlineHandle(TpassHandle* p, const pf::host::VadOptions& o) : parent(p), vad(o) {}
  void Restart() { vad.Reset(); pcm.clear(); base = 0; total = 0; scored = 0; }
};

constexpr int kVadLookbackFrames = 100;   // > the VAD network's left context (4 FSMN layers x 19 taps + 2 LFR frames = 78 frames)
constexpr int kVadRightFrames = 2;        // LFR 5/1 looks two frames ahead: the last two frames of an unfinished stream are provisional
}  // namespace

FUNASR_HANDLE FunTpassInit(std::map<std::string, std::string>& model_path, int thread_num) {
  // keys as TpassStream reads them (tpass-stream.cpp): "model-dir" (offline AM), "vad-dir", "punc-dir" (the REALTIME punctuation
  // model here), "online-model-dir" (ignored: the streaming AM stays on the reference host), "itn-dir" / "lm-dir" (ignored)
  if (model_path.find("vad-dir") == model_path.end()) { fprintf(stderr, "FunTpassInit: vad-dir missing (the 2-pass stream is cut by the VAD)\n"); return nullptr; }
  std::map<std::string, std::string> off(model_path);
  off.erase("punc-dir");   // the offline handle's own punctuation hook is the OFFLINE model; the 2-pass stream uses the realtime one
  std::unique_ptr<TpassHandle> t(new TpassHandle);
  t->offline = (OfflineHandle*)FunOfflineInit(off, thread_num, false, 1);
  if (!t->offline) return nullptr;
  auto pd = model_path.find("punc-dir");
  if (pd != model_path.end() && !pd->second.empty()) {
    t->punc_online.reset(new funasr_b200::CTTransformerOnlineB200(ToInt(model_path, "device", 0), ToInt(model_path, "punc-max-tokens", 0)));
    std::string err;
    if (!t->punc_online->Init(pd->second, &err)) { fprintf(stderr, "FunTpassInit: punc-dir %s: %s\n", pd->second.c_str(), err.c_str()); return nullptr; }
  }
  return t.release();
}

FUNASR_HANDLE FunTpassOnlineInit(FUNASR_HANDLE tpass_handle, std::vector<int> chunk_size) {
  (void)chunk_size;   // the streaming model's chunking; the offline leg is cut by the VAD alone
  TpassHandle* t = (TpassHandle*)tpass_handle;
  if (!t) return nullptr;
  pf::host::VadOptions vo;
  vo.speech_noise_thres = t->offline->vad_thres;
  return new TpassOnlineHandle(t, vo);
}

void FunTpassUninit(FUNASR_HANDLE handle) { delete (TpassHandle*)handle; }
void FunTpassOnlineUninit(FUNASR_HANDLE handle) { delete (TpassOnlineHandle*)handle; }

FUNASR_RESULT FunTpassInferBuffer(FUNASR_HANDLE handle, FUNASR_HANDLE online_handle, const char* sz_buf, int n_len,
                                  std::vector<std::vector<std::string>>& punc_cache, bool input_finished, int sampling_rate,
                                  std::string wav_format, ASR_TYPE mode, const std::vector<std::vector<float>>& hw_emb, bool itn,
                                  int vad_tail_sil, int vad_max_len, FUNASR_DEC_HANDLE dec_handle, std::string svs_lang, bool svs_itn) {
  (void)itn; (void)dec_handle; (void)svs_lang; (void)svs_itn;
  TpassHandle* t = (TpassHandle*)handle;
  TpassOnlineHandle* o = (TpassOnlineHandle*)online_handle;
  if (!t || !o) return nullptr;                                                     // funasrruntime.cpp:500-501
  OfflineHandle* h = t->offline;
  if (!(wav_format == "pcm" || wav_format == "PCM")) { fprintf(stderr, "Wrong wav_format: %s\n", wav_format.c_str()); return nullptr; }
  if (sampling_rate != h->model()->GetAsrSampleRate()) { fprintf(stderr, "FunTpassInferBuffer: resampling stays on the reference host path\n"); return nullptr; }
  if (punc_cache.size() < 2) punc_cache.resize(2);
  o->vad.SetOptions(vad_tail_sil, vad_max_len);                                      // vad_online_handle->SetConfig (:506)
  const long long n = n_len / 2;
  const unsigned char* bytes = (const unsigned char*)sz_buf;
  const size_t old = o->pcm.size();
  o->pcm.resize(old + (size_t)n);
  for (long long i = 0; i < n; ++i) o->pcm[old + i] = (short)((bytes[2 * i + 1] << 8) | bytes[2 * i]);   // LoadPcmwavOnline
  o->total += n;
  RecogResult* res = new RecogResult;
  res->snippet_time = (float)n / h->model()->GetAsrSampleRate();

  // ---- VAD: silence probabilities of the frames that became final with this chunk, then the E2E state machine ----
  const int f_total = o->total >= 400 ? (int)(1 + (o->total - 400) / 160) : 0;
  const int f_final = input_finished ? f_total : std::max(o->scored, f_total - kVadRightFrames);
  std::vector<std::pair<int, int>> closed;
  if (f_final > o->scored || (input_finished && f_total > 0 && f_final == o->scored)) {
    const int w0 = std::max((int)(o->base / 160), std::max(0, o->scored - kVadLookbackFrames));   // first frame of the scoring window
    const long long s0 = (long long)w0 * 160 - o->base;                                           // its first sample inside pcm
    const long long ns = (long long)o->pcm.size() - s0;
    const int wf = ns >= 400 ? (int)(1 + (ns - 400) / 160) : 0;
    std::vector<float> sil((size_t)std::max(1, wf));
    if (wf > 0) {
      int64_t off[2] = {0, (int64_t)ns};
      int32_t frame_off[2] = {0, 0};
      std::lock_guard<std::mutex> lk(h->vad_mu);
      if (b200pf_vad_scores_s16(h->vad, (const int16_t*)o->pcm.data() + s0, off, 1, sil.data(), (int64_t)sil.size(), frame_off, nullptr, nullptr) != 0) {
        fprintf(stderr, "FunTpassInferBuffer: VAD: %s\n", b200pf_last_error());
        delete res;
        return nullptr;
      }
    }
    const int first = o->scored - w0, count = f_final - o->scored;
    if (count > 0 && first >= 0 && first + count <= wf) closed = o->vad.Push(sil.data() + first, count, input_finished);
    else if (input_finished) closed = o->vad.Push(sil.data(), 0, true);
    o->scored = f_final;
  }

  // ---- the offline leg: every closed VAD segment through the offline model (FetchTpass loop, :570-636) ----
  if (mode != ASR_ONLINE) {
    std::string cur_stamp = "[";
    const long long per_ms = h->model()->GetAsrSampleRate() / 1000;
    for (const auto& sg : closed) {
      const long long sb = std::max<long long>(o->base, std::min<long long>(o->total, sg.first * per_ms));
      const long long se = std::max<long long>(sb, std::min<long long>(o->total, sg.second * per_ms));
      if (se <= sb) continue;
      const long long len = se - sb;
      std::vector<float> fl((size_t)len);
      const short* src = o->pcm.data() + (sb - o->base);
      for (long long i = 0; i < len; ++i) fl[i] = (float)src[i] / 32768.0f;
      float* buff[1] = {fl.data()};
      int l[1] = {(int)len};
      std::vector<std::string> msgs = h->batcher ? h->batcher->Forward(buff, l, true, hw_emb, nullptr, 1)
                                                 : h->model()->Forward(buff, l, true, hw_emb, nullptr, 1);
      std::string msg = msgs.empty() ? std::string() : msgs[0];
      // "text | b0, e0,b1, e1,..." (PostProcess, util.cpp:820-835): seconds relative to the segment
      std::string stamps;
      const size_t bar = msg.find(" | ");
      if (bar != std::string::npos) { stamps = msg.substr(bar + 3); msg = msg.substr(0, bar); }
      if (msg.empty() && stamps.empty() && msgs.empty()) continue;
      if (!stamps.empty()) {
        std::vector<std::string> parts;
        size_t p0 = 0;
        while (p0 <= stamps.size()) {
          size_t q = stamps.find(',', p0);
          if (q == std::string::npos) q = stamps.size();
          if (q > p0) parts.push_back(stamps.substr(p0, q - p0));
          p0 = q + 1;
        }
        for (size_t i = 0; i + 1 < parts.size(); i += 2) {
          const float begin = std::stof(parts[i]) + (float)sg.first / 1000.0f;       // frame->global_start is in ms
          const float end = std::stof(parts[i + 1]) + (float)sg.first / 1000.0f;
          cur_stamp += "[" + std::to_string((int)(1000 * begin)) + "," + std::to_string((int)(1000 * end)) + "],";
        }
      }
      if (cur_stamp != "[") {
        cur_stamp.erase(cur_stamp.length() - 1);
        res->stamp += cur_stamp + "]";
      }
      std::string msg_punc = msg;
      if (t->punc_online) {
        msg_punc = t->punc_online->AddPunc(msg.c_str(), punc_cache[1], h->model()->GetLang());
        if (input_finished) msg_punc += "\xe3\x80\x82";                            // "。" (:612-614)
      }
      res->tpass_msg = msg_punc;                                                     // the reference keeps the LAST segment's text per call
      if (!res->stamp.empty()) res->stamp_sents = pf::host::SentenceStamps(res->tpass_msg, res->stamp);
    }
  }

  // ---- drop what can no longer be needed: audio before the open segment and before the scoring look-back ----
  long long keep_from = (long long)std::max(0, o->scored - kVadLookbackFrames) * 160;
  const int open_ms = o->vad.open_start_ms();
  if (open_ms >= 0) keep_from = std::min<long long>(keep_from, (long long)open_ms * 16);
  keep_from = std::max(keep_from, o->base);
  if (keep_from - o->base > (1 << 20)) {   // trim in large steps only
    o->pcm.erase(o->pcm.begin(), o->pcm.begin() + (size_t)(keep_from - o->base));
    o->base = keep_from;
  }
  if (input_finished) o->Restart();        // audio->ResetIndex() (:641-643)
  return res;
}

const std::vector<std::vector<float>> CompileHotwordEmbedding(FUNASR_HANDLE handle, std::string& hotwords, ASR_TYPE mode) {
  std::vector<std::vector<float>> emb;
  if (!handle) return emb;
  if (mode == ASR_ONLINE) { fprintf(stderr, "Not implement: Online model does not support Hotword yet!\n"); return emb; }   // funasrruntime.cpp:482-486
  OfflineHandle* h = mode == ASR_TWO_PASS ? ((TpassHandle*)handle)->offline : (OfflineHandle*)handle;
  if (!h) return emb;
  return h->model()->CompileHotwordEmbedding(hotwords);
}

void FunOfflineUninit(FUNASR_HANDLE handle) { delete (OfflineHandle*)handle; }

const char* FunASRGetResult(FUNASR_RESULT result, int) { return result ? ((RecogResult*)result)->msg.c_str() : nullptr; }
const char* FunASRGetStamp(FUNASR_RESULT result) { return result ? ((RecogResult*)result)->stamp.c_str() : nullptr; }
const char* FunASRGetStampSents(FUNASR_RESULT result) { return result ? ((RecogResult*)result)->stamp_sents.c_str() : nullptr; }
const char* FunASRGetTpassResult(FUNASR_RESULT result, int) { return result ? ((RecogResult*)result)->tpass_msg.c_str() : nullptr; }
const int FunASRGetRetNumber(FUNASR_RESULT result) { return result ? 1 : 0; }
void FunASRFreeResult(FUNASR_RESULT result) { delete (RecogResult*)result; }
const float FunASRGetRetSnippetTime(FUNASR_RESULT result) { return result ? ((RecogResult*)result)->snippet_time : 0.0f; }

// ---- punctuation API (funasrruntime.h:92-96) ----
namespace {
struct PuncResult {  // FUNASR_PUNC_RESULT, commonfunc.h
  std::string msg;
  std::vector<std::string> arr_cache;
};
}  // namespace

FUNASR_HANDLE CTTransformerInit(std::map<std::string, std::string>& model_path, int thread_num, PUNC_TYPE type) {
  (void)thread_num;
  auto it = model_path.find("punc-dir");
  if (it == model_path.end()) { fprintf(stderr, "CTTransformerInit: punc-dir missing\n"); return nullptr; }
  std::string err;
  const int device = ToInt(model_path, "device", 0), max_tokens = ToInt(model_path, "punc-max-tokens", 0);
  if (type == PUNC_ONLINE) {   // CreatePuncModel: PUNC_ONLINE -> CTTransformerOnline (punc-model.cpp)
    std::unique_ptr<funasr_b200::CTTransformerOnlineB200> p(new funasr_b200::CTTransformerOnlineB200(device, max_tokens));
    if (!p->Init(it->second, &err)) { fprintf(stderr, "CTTransformerInit: %s\n", err.c_str()); return nullptr; }
    return (funasr_b200::PuncModel*)p.release();
  }
  std::unique_ptr<funasr_b200::CTTransformerB200> p(new funasr_b200::CTTransformerB200(device, max_tokens));
  if (!p->Init(it->second, &err)) { fprintf(stderr, "CTTransformerInit: %s\n", err.c_str()); return nullptr; }
  return (funasr_b200::PuncModel*)p.release();
}

FUNASR_RESULT CTTransformerInfer(FUNASR_HANDLE handle, const char* sz_sentence, FUNASR_MODE, QM_CALLBACK, PUNC_TYPE type, FUNASR_RESULT pre_result) {
  funasr_b200::PuncModel* punc = (funasr_b200::PuncModel*)handle;
  if (!punc) return nullptr;
  if (type == PUNC_OFFLINE) {
    PuncResult* r = new PuncResult;
    r->msg = punc->AddPunc(sz_sentence);
    return r;
  }
  PuncResult* r = pre_result ? (PuncResult*)pre_result : new PuncResult;   // funasrruntime.cpp:193-198
  r->msg = punc->AddPunc(sz_sentence, r->arr_cache);
  return r;
}

const char* CTTransformerGetResult(FUNASR_RESULT result, int) { return result ? ((PuncResult*)result)->msg.c_str() : nullptr; }
void CTTransformerFreeResult(FUNASR_RESULT result) { delete (PuncResult*)result; }
void CTTransformerUninit(FUNASR_HANDLE handle) { delete (funasr_b200::PuncModel*)handle; }

FUNASR_DEC_HANDLE FunASRWfstDecoderInit(FUNASR_HANDLE, int, float, float, float) { return nullptr; }
void FunASRWfstDecoderUninit(FUNASR_DEC_HANDLE) {}
void FunWfstDecoderLoadHwsRes(FUNASR_DEC_HANDLE, int, std::unordered_map<std::string, int>&) {}
void FunWfstDecoderUnloadHwsRes(FUNASR_DEC_HANDLE) {}
