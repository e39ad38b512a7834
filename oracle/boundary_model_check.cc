// ORACLE-side test infrastructure (built by `make -C oracle boundary`, run by tests/test_boundary_cpu.py).
// ParaformerB200 / MicroBatcher / MultiGpuParaformer compiled with -DB200PF_WITH_REFERENCE_HEADERS against the reference's OWN
// onnxruntime/include/model.h and onnxruntime/src/wfst-decodable.h: they ARE funasr::Model subclasses (no look-alike base), a
// ParaformerB200 can sit in OfflineStream::asr_handle (std::unique_ptr<funasr::Model>, offline-stream.h) and the dynamic_cast of
// FunASRWfstDecoderInit (funasrruntime.cpp:841) finds the WfstDecodable interface.  No GPU is touched: nothing is initialised.
#include <cstdio>
#include <memory>

#include "../asr-2pass_b200/csrc/host/micro_batcher.h"
#include "../asr-2pass_b200/csrc/host/multi_gpu.h"
#include "../asr-2pass_b200/csrc/host/paraformer_b200.h"

int main() {
  std::unique_ptr<funasr::Model> asr_handle(new funasr_b200::ParaformerB200(0, 0, 0));   // the member OfflineStream holds
  funasr::WfstDecodable* dec = dynamic_cast<funasr::WfstDecodable*>(asr_handle.get());
  int fails = 0;
  fails += dec == nullptr;
  fails += dec && dec->GetLm() != nullptr;             // no LM loaded -> the CtcPrefixDecoder branch (funasrruntime.cpp:847-849)
  fails += asr_handle->GetBatchSize() != 1;
  asr_handle->SetBatchSize(8);
  fails += asr_handle->GetBatchSize() != 8;
  fails += asr_handle->Rescoring() != "";
  std::unique_ptr<funasr::Model> pool(new funasr_b200::MultiGpuParaformer(std::vector<int>{0, 1}, 0, 0));
  fails += pool == nullptr;
  printf("boundary_model_check %s (%d)\n", fails ? "FAILED" : "ok", fails);
  return fails ? 1 : 0;
}
