// Contended global reductions: 148 CTAs x 8 warps x 16 lanes x 16 values onto 256 addresses (the conv epilogue's
// BatchNorm-sum pattern), fp32 RED vs 64-bit integer RED.   nvcc -arch=sm_100a -o atomic_probe atomic_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* f, unsigned long long* u, int reps) {
  const int lane = threadIdx.x & 31;
  if (lane & 1) return;
  const int col = lane >> 1;
  for (int r = 0; r < reps; ++r)
    for (int c = 0; c < 8; ++c) {
      const int idx = c * 16 + col;
      if (MODE == 0) { atomicAdd(&f[idx], 1.f); atomicAdd(&f[128 + idx], 2.f); }
      else { atomicAdd(&u[idx], 3ull); atomicAdd(&u[128 + idx], 5ull); }
    }
}
int main() {
  float* f; unsigned long long* u;
  cudaMalloc(&f, 4096); cudaMalloc(&u, 4096); cudaMemset(f, 0, 4096); cudaMemset(u, 0, 4096);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      for (int i = 0; i < 20; ++i) { if (mode == 0) k<0><<<148, 256>>>(f, u, 1); else k<1><<<148, 256>>>(f, u, 1); }
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("mode %d (%s): %.2f us per launch (303104 atomics on 256 addresses)\n", mode, mode ? "u64 RED" : "f32 RED", ms * 1000 / 20);
    }
  return 0;
}
