// ORACLE-side test infrastructure: the two glog symbols the reference's benchmark binaries (onnxruntime/bin/*.cpp) reference
// beyond what oracle/funasr_text_ref_shim.cc already defines (google::InitGoogleLogging, FLAGS_logtostderr).  The vendored glog
// library itself is not built.
#include <glog/logging.h>
namespace google {
void InitGoogleLogging(const char*) {}
}  // namespace google
namespace fLB { bool FLAGS_logtostderr = true; }
