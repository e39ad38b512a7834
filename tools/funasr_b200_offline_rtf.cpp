// funasr-b200-offline-rtf — the reference's own benchmark harness (onnxruntime/bin/funasr-onnx-offline-rtf.cpp) written
// against libfunasr_b200.so: same call sequence and the same accounting.
//   * FunOfflineInit(model_path, /*thread_num=*/1, use_gpu, batch_size)                       (rtf.cpp:187)
//   * P std::threads share the ONE handle and pull wav indices from an atomic counter           (rtf.cpp:63-68,247-250)
//   * one warm-up FunOfflineInfer per thread, then gettimeofday around every call                (rtf.cpp:55-76)
//   * total_time = max over threads of the summed call time, total_length = sum of snippet_time,
//     total_rtf = total_time / total_length, speedup = 1 / total_rtf                            (rtf.cpp:96-101,257-260)
// Usage: funasr-b200-offline-rtf --model-dir D --wav-scp wav.scp --thread-num P [--micro-batch-us U] [--devices 0,1]
#include <sys/time.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../asr-2pass_b200/csrc/host/funasrruntime_b200.h"

static std::atomic<int> wav_index(0);
static std::mutex mtx;

static void run(FUNASR_HANDLE h, const std::vector<std::string>& wavs, float* total_length, long* total_time) {
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  FUNASR_RESULT r = FunOfflineInfer(h, wavs[0].c_str(), RASR_NONE, nullptr, hw, 16000, true, 250, 20000);   // warm up
  if (r) FunASRFreeResult(r);
  float n_len = 0.f;
  long n_time = 0;
  for (;;) {
    const int i = wav_index.fetch_add(1);
    if (i >= (int)wavs.size()) break;
    struct timeval a, b;
    gettimeofday(&a, nullptr);
    r = FunOfflineInfer(h, wavs[i].c_str(), RASR_NONE, nullptr, hw, 16000, true, 250, 20000);
    gettimeofday(&b, nullptr);
    n_time += (b.tv_sec - a.tv_sec) * 1000000L + b.tv_usec - a.tv_usec;
    if (r) { n_len += FunASRGetRetSnippetTime(r); FunASRFreeResult(r); }
    else fprintf(stderr, "%s: No return data!\n", wavs[i].c_str());
  }
  std::lock_guard<std::mutex> g(mtx);
  *total_length += n_len;
  if (*total_time < n_time) *total_time = n_time;
}

int main(int argc, char** argv) {
  std::map<std::string, std::string> mp;
  std::string scp;
  int threads = 1;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k == "--model-dir") mp["model-dir"] = v;
    else if (k == "--wav-scp" || k == "--wav-path") scp = v;
    else if (k == "--thread-num") threads = atoi(v.c_str());
    else if (k == "--micro-batch-us") mp["micro-batch-us"] = v;
    else if (k == "--devices") mp["devices"] = v;
    else if (k == "--max-rows") mp["max-rows"] = v;
    else { fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  if (!mp.count("model-dir") || scp.empty()) { fprintf(stderr, "usage: --model-dir D --wav-scp wav.scp --thread-num P [--micro-batch-us U] [--devices 0,1]\n"); return 2; }
  std::vector<std::string> wavs;
  std::ifstream in(scp);
  std::string line;
  while (std::getline(in, line)) {        // "<id> <path>" per line (rtf.cpp:207-221)
    std::istringstream iss(line);
    std::string id, path;
    if (iss >> id >> path) wavs.push_back(path);
  }
  if (wavs.empty()) { fprintf(stderr, "no wavs in %s\n", scp.c_str()); return 2; }
  struct timeval a, b;
  gettimeofday(&a, nullptr);
  FUNASR_HANDLE h = FunOfflineInit(mp, 1, true, 256);
  if (!h) { fprintf(stderr, "FunOfflineInit failed\n"); return 1; }
  gettimeofday(&b, nullptr);
  fprintf(stderr, "Model initialization takes %.3f s\n", (double)((b.tv_sec - a.tv_sec) * 1000000L + b.tv_usec - a.tv_usec) / 1e6);
  float total_length = 0.f;
  long total_time = 0;
  std::vector<std::thread> th;
  gettimeofday(&a, nullptr);
  for (int t = 0; t < threads; ++t) th.emplace_back(run, h, std::cref(wavs), &total_length, &total_time);
  for (auto& t : th) t.join();
  gettimeofday(&b, nullptr);
  const double wall = (double)((b.tv_sec - a.tv_sec) * 1000000L + b.tv_usec - a.tv_usec) / 1e6;
  // the reference's own lines (rtf.cpp:257-260), then one JSON line for profiles/
  printf("total_time_wav %ld ms\n", (long)(total_length * 1000));
  printf("total_time_comput %ld ms\n", total_time / 1000);
  printf("total_rtf %05lf\n", (double)total_time / (total_length * 1000000));
  printf("speedup %05lf\n", 1.0 / ((double)total_time / (total_length * 1000000)));
  printf("{\"harness\": \"funasr-onnx-offline-rtf pattern\", \"threads\": %d, \"wavs\": %zu, \"audio_s\": %.2f, \"max_thread_s\": %.4f, \"wall_s_incl_warmup\": %.4f, "
         "\"speedup\": %.1f, \"micro_batch_us\": %s, \"devices\": \"%s\"}\n",
         threads, wavs.size(), total_length, total_time / 1e6, wall, 1.0 / ((double)total_time / (total_length * 1000000)),
         mp.count("micro-batch-us") ? mp["micro-batch-us"].c_str() : "0", mp.count("devices") ? mp["devices"].c_str() : "0");
  FunOfflineUninit(h);
  return 0;
}
