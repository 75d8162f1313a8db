// Floor of a chain of dependent kernel launches on B200: CUDA graph of N kernels, with / without programmatic
// dependent launch, empty body vs a tiny dependent load-store body.   nvcc -arch=sm_100a -o launch_floor launch_floor.cu
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_empty(float* p, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}
__global__ void k_body(float* p, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float coef = p[1 << 20];                 // "statistics" produced by the predecessor
  float v = p[i] * 0.5f + coef;                   // dependent load
  p[i] = v;
  if (threadIdx.x == 0) atomicAdd(&p[(1 << 20) + 1], v);
}
template <typename K>
float run(K kern, int nk, int grid, int pdl, float* d, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < nk; ++i) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, d, pdl);
  }
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / (20.f * nk);
}
int main() {
  float* d; cudaMalloc(&d, (2 << 20) * sizeof(float)); cudaMemset(d, 0, (2 << 20) * sizeof(float));
  cudaStream_t s; cudaStreamCreate(&s);
  for (int grid : {8, 148, 1184})
    for (int pdl : {0, 1}) {
      printf("grid %4d pdl %d: empty %.2f us/launch, body %.2f us/launch\n", grid, pdl, run(k_empty, 150, grid, pdl, d, s),
             run(k_body, 150, grid, pdl, d, s));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
