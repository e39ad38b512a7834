// Kernel launch helper: every kernel of the forward pass is launched with programmatic dependent launch
// (PDL), so kernel N+1's prologue (barrier init, TMEM allocation, tensor-map prefetch, table loads) overlaps
// kernel N's tail.  Each kernel executes pdl_wait() before it touches any activation memory and
// pdl_launch_dependents() right after, which keeps the stream's data dependencies exactly as without PDL.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace pf {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200PF_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// cudaFuncSetAttribute applies to the CURRENT device only: with one engine per GPU inside one process (MultiGpuParaformer)
// every device needs its own call.  `flags` is a per-call-site static array; returns true the first time on each device.
inline bool first_use_on_device(bool (&flags)[64]) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
  if (flags[d]) return false;
  flags[d] = true;
  return true;
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
