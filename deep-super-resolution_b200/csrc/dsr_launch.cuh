// dsr_launch.cuh -- programmatic dependent launch (PDL) plumbing shared by every kernel of the DIP step.
//
// One DIP iteration is ~180 dependent launches, most of them a few microseconds long (the 16x16 .. 128x128
// levels), so the launch -> drain -> launch gap between consecutive kernels is a first-order cost.  Every kernel
// is therefore launched with cudaLaunchAttributeProgrammaticStreamSerialization and follows one protocol:
//
//   prologue that touches NO data produced by earlier kernels of the step except weights packed >= 2 launches
//   earlier (barrier init, TMEM allocation, descriptor prefetch, resident-weight TMA)
//   pdl_wait()      -- griddepcontrol.wait: the predecessor grid has completed and its writes are visible
//   pdl_trigger()   -- griddepcontrol.launch_dependents: the successor may start ITS prologue now
//   body
//
// Because a kernel only triggers after its own wait, at most two grids overlap (N running, N+1 in its prologue)
// and "N+1 started" implies "N-1 completed": anything produced two or more launches earlier may be read before
// the wait.  The same attribute is honoured under stream capture (programmatic graph edges), so the CUDA-graph
// replay of the iteration keeps the overlap.  DSR_PDL=0 turns the attribute off (plain stream order) for A/B.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace dsr {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DSR_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace dsr
