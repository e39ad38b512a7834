// Kernel launch helper: every kernel of the forward pass is launched with programmatic dependent launch
// (PDL), so kernel N+1's prologue (barrier init, TMEM allocation, tensor-map prefetch, table loads) overlaps
// kernel N's tail.  Each kernel executes pdl_wait() before it touches any activation memory and
// pdl_launch_dependents() right after, which keeps the stream's data dependencies exactly as without PDL.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace pf {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200PF_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// cudaFuncSetAttribute applies to the CURRENT device only: with one engine per GPU inside one process (MultiGpuParaformer)
// every device needs its own call, and two host threads may reach their first launch on one device at the same time (the VAD
// engine under its own lock next to the acoustic model).  `once` is a per-call-site static; `init` runs exactly once per device,
// under a lock, and the device is only marked done after it succeeded -- later launches take the lock-free path.
struct PerDeviceOnce {
  std::mutex mu;
  std::atomic<bool> done[64];
  PerDeviceOnce() { for (auto& d : done) d.store(false); }
};
template <class F>
inline int once_per_device(PerDeviceOnce& once, F&& init) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return init();
  if (once.done[d].load(std::memory_order_acquire)) return 0;
  std::lock_guard<std::mutex> lock(once.mu);
  if (once.done[d].load(std::memory_order_relaxed)) return 0;
  const int rc = init();
  if (rc == 0) once.done[d].store(true, std::memory_order_release);
  return rc;
}

template <typename... KArgs, typename... Args>
inline int launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace pf
