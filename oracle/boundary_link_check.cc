// ORACLE-side test infrastructure (built by `make -C oracle boundary`, run by tests/test_boundary_cpu.py).
// A caller written against the REFERENCE's own exported header -- onnxruntime/include/funasrruntime.h, included from where it
// lies -- that names every entry point the reference's servers and benchmark binaries call on the offline and the 2-pass path
// (websocket/bin/websocket-server.cpp, websocket-server-2pass.cpp, funasr-wss-server*.cpp, onnxruntime/bin/funasr-onnx-offline*.cpp).
// It must LINK against libfunasr_b200.so with no other provider of these symbols: same names, same C++ signatures.
// Run without arguments it only checks the null-handle behaviour (no GPU needed); with a model directory it runs one request.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "funasrruntime.h"

int main(int argc, char** argv) {
  std::vector<std::vector<float>> hw(1, std::vector<float>(512, 0.f));
  std::vector<std::vector<std::string>> punc_cache(2);
  std::unordered_map<std::string, int> hws;
  std::string hotwords;
  int fails = 0;
  // null handles: nullptr / empty results, never a crash (funasrruntime.cpp:216-217, 500-501)
  fails += FunOfflineInferBuffer(nullptr, "", 0, RASR_NONE, nullptr, hw, 16000, "pcm", true, 800, 60000) != nullptr;
  fails += FunOfflineInfer(nullptr, "x.wav", RASR_NONE, nullptr, hw, 16000, true, 800, 60000) != nullptr;
  fails += FunTpassInferBuffer(nullptr, nullptr, "", 0, punc_cache, true, 16000, "pcm", ASR_TWO_PASS, hw, true, 800, 60000) != nullptr;
  fails += FunASRGetResult(nullptr, 0) != nullptr;
  fails += FunASRGetStamp(nullptr) != nullptr;
  fails += FunASRGetStampSents(nullptr) != nullptr;
  fails += FunASRGetTpassResult(nullptr, 0) != nullptr;
  fails += FunASRGetRetNumber(nullptr) != 0;
  fails += FunASRGetRetSnippetTime(nullptr) != 0.0f;
  fails += !CompileHotwordEmbedding(nullptr, hotwords, ASR_OFFLINE).empty();
  fails += !CompileHotwordEmbedding(nullptr, hotwords, ASR_TWO_PASS).empty();
  fails += FunTpassOnlineInit(nullptr) != nullptr;
  FunASRFreeResult(nullptr);
  FunOfflineUninit(nullptr);
  FunTpassUninit(nullptr);
  FunTpassOnlineUninit(nullptr);
  FunOfflineReset(nullptr);
  FUNASR_DEC_HANDLE dec = FunASRWfstDecoderInit(nullptr, ASR_OFFLINE, 3.0f, 3.0f, 10.0f);
  FunWfstDecoderLoadHwsRes(dec, 20, hws);
  FunWfstDecoderUnloadHwsRes(dec);
  FunASRWfstDecoderUninit(dec);
  if (argc > 1) {   // one real request through the reference's call sequence (needs a B200)
    std::map<std::string, std::string> model_path;
    model_path["model-dir"] = argv[1];
    FUNASR_HANDLE h = FunOfflineInit(model_path, 1, true, 1);
    if (!h) { printf("FunOfflineInit failed\n"); return 2; }
    std::vector<short> pcm(16000 * 3);
    for (size_t i = 0; i < pcm.size(); ++i) pcm[i] = (short)((i * 7919u) % 4001) - 2000;
    FUNASR_RESULT r = FunOfflineInferBuffer(h, (const char*)pcm.data(), (int)pcm.size() * 2, RASR_NONE, nullptr, hw, 16000, "pcm", true, 800, 60000);
    if (!r) { printf("FunOfflineInferBuffer failed\n"); return 3; }
    printf("text_bytes=%zu snippet=%.3f\n", strlen(FunASRGetResult(r, 0)), FunASRGetRetSnippetTime(r));
    FunASRFreeResult(r);
    FunOfflineUninit(h);
  }
  printf("boundary_link_check %s (%d)\n", fails ? "FAILED" : "ok", fails);
  return fails ? 1 : 0;
}
