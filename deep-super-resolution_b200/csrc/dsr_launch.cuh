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
// ---- in-kernel stamps (diagnostics, -DDSR_KSTAMP builds only; DSR_TIMELINE=2 at run time): thread 0 of block 0 of
// every kernel records %globaltimer at entry and after its griddepcontrol.wait (record index = arrival order), and
// ks_end() adds the time at which block 0 finished its body.  No extra launches, so the programmatic overlap of the
// replayed graph is untouched.
#ifdef DSR_KSTAMP
static __device__ unsigned long long* g_ks_buf = nullptr;   // [0] counter, then 8 words per record (entry, wait done, end, grid|block, 4 marks)
__device__ __forceinline__ unsigned long long ks_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ unsigned int g_ks_slot;                    // record index of the running kernel (block 0 writes it)
#endif
__device__ __forceinline__ void pdl_sync() {
#ifdef DSR_KSTAMP
  const bool rec = (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) && g_ks_buf != nullptr;
  unsigned long long t0 = 0;
  if (rec) t0 = ks_now();
#endif
  pdl_wait();
#ifdef DSR_KSTAMP
  if (rec) {
    const unsigned long long t1 = ks_now();
    const unsigned long long idx = atomicAdd(g_ks_buf, 1ull) % 4000ull;     // ring: the dump reads the last iteration
    {
      g_ks_buf[1 + 8 * idx] = t0;
      g_ks_buf[2 + 8 * idx] = t1;
      g_ks_buf[3 + 8 * idx] = 0;
      for (int m = 0; m < 4; ++m) g_ks_buf[5 + 8 * idx + m] = 0;
      g_ks_buf[4 + 8 * idx] = (static_cast<unsigned long long>(gridDim.x * gridDim.y * gridDim.z) << 32) | blockDim.x;
      g_ks_slot = static_cast<unsigned int>(idx);
    }
  }
#endif
  pdl_trigger();
}
// end-of-body stamp of block 0 (call where every thread of the block is done, thread 0 records)
__device__ __forceinline__ void ks_end() {
#ifdef DSR_KSTAMP
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && g_ks_buf != nullptr && g_ks_slot < 4000)
    g_ks_buf[3 + 8 * g_ks_slot] = ks_now();
#endif
}
// intermediate stamp m (0..3) of block 0 (diagnostics of one kernel's phases; call from thread 0's control path)
__device__ __forceinline__ void ks_mark(int m) {
#ifdef DSR_KSTAMP
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && g_ks_buf != nullptr && g_ks_slot < 4000)
    g_ks_buf[5 + 8 * g_ks_slot + m] = ks_now();
#else
  (void)m;
#endif
}
// the same from any ONE thread of block 0 chosen by the caller (warp-specialised kernels)
__device__ __forceinline__ void ks_mark_here(int m) {
#ifdef DSR_KSTAMP
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && g_ks_buf != nullptr && g_ks_slot < 4000)
    g_ks_buf[5 + 8 * g_ks_slot + m] = ks_now();
#else
  (void)m;
#endif
}
#ifdef DSR_KSTAMP
#define DSR_KSTAMP_SETTER(name) \
  void name(unsigned long long* buf) { cudaMemcpyToSymbol(g_ks_buf, &buf, sizeof(buf)); }
#else
#define DSR_KSTAMP_SETTER(name) \
  void name(unsigned long long*) {}
#endif

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DSR_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

// ---- in-graph timeline (diagnostics, DSR_TIMELINE=1): a one-thread stamp kernel after every launch writes
// %globaltimer into a slot; consecutive stamps bracket each kernel as it runs inside the replayed graph (warm caches,
// real dependencies; the stamp costs ~1 us and ends the programmatic overlap with the next launch).
struct Timeline {
  unsigned long long* buf = nullptr;     // device, kTimelineSlots entries
  int n = 0;
  const void* fn[2048];
  unsigned grid[2048];
  cudaStream_t stream[2048];
};
constexpr int kTimelineSlots = 2048;
inline Timeline g_timeline;
inline int timeline_mode() {
  static const int mode = [] {
    const char* e = getenv("DSR_TIMELINE");
    return e != nullptr ? atoi(e) : 0;
  }();
  return mode;
}
inline bool timeline_enabled() { return timeline_mode() != 0; }
static __global__ void timeline_stamp_kernel(unsigned long long* buf, int idx) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  buf[idx] = t;
}
inline void timeline_stamp(const void* fn, unsigned grid, cudaStream_t s) {
  Timeline& t = g_timeline;
  if (t.buf == nullptr) {
    if (cudaMalloc(&t.buf, kTimelineSlots * sizeof(unsigned long long)) != cudaSuccess) return;
    cudaMemset(t.buf, 0, kTimelineSlots * sizeof(unsigned long long));
  }
  if (t.n >= kTimelineSlots) return;
  t.fn[t.n] = fn;
  t.grid[t.n] = grid;
  t.stream[t.n] = s;
  if (timeline_mode() == 1) timeline_stamp_kernel<<<1, 1, 0, s>>>(t.buf, t.n);
  ++t.n;
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
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (timeline_enabled()) timeline_stamp(reinterpret_cast<const void*>(kernel), grid.x * grid.y * grid.z, s);
  return e;
}

// the same with a thread-block cluster of `cluster_x` CTAs (kernels without __cluster_dims__)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                    unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  if (timeline_enabled()) timeline_stamp(reinterpret_cast<const void*>(kernel), grid.x * grid.y * grid.z, s);
  return e;
}

}  // namespace dsr
