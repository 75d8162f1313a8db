// umma_probe.cu -- hardware probe (not part of the library): does a tcgen05 K-major SWIZZLE_128B A descriptor
// whose start address is shifted by whole 128-byte rows (not 1024-aligned), with an arbitrary stride between
// 8-row groups, read the rows one expects?  Needed for the halo-tile implicit-GEMM (one smem tile, 9 shifted views).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <vector>
#include "../deep-super-resolution_b200/csrc/dsr_ptx.cuh"
using namespace dsr;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kRows = 256;   // smem A rows (128 B each)

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap mapA,
                                               const __grid_constant__ CUtensorMap mapB, float* out, int shift_rows,
                                               int sbo_bytes, int base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 256 x 128 B
  uint8_t* sB = smem + kRows * 128;         // 64 x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], kRows * 128 + 64 * 128);
    tma_load_2d(&mapA, &bars[0], sA, 0, 0);
    tma_load_2d(&mapB, &bars[0], sB, 0, 0);
    mbar_wait(&bars[0], 0, nullptr, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, 64, 0, 0, 0, 0);
    uint64_t da = make_smem_desc(smem_u32(sA) + shift_rows * 128, 16, sbo_bytes, SWZ_128B);
    da |= static_cast<uint64_t>(base_offset & 7) << 49;
    const uint64_t db = make_smem_desc(smem_u32(sB), 16, 1024, SWZ_128B);
    for (int j = 0; j < 4; ++j) umma_f16(tmem, da + j * 2, db + j * 2, idesc, j != 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, nullptr, 0);
  tc_fence_after();
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16];
    tmem_ld16(taddr + c * 16, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c * 16 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  // A_g[r][c]: pass 0 = r, pass 1 = c ; B = identity [64][64]
  std::vector<__half> hA(kRows * 64), hB(64 * 64);
  __half *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  for (int i = 0; i < 64; ++i)
    for (int j = 0; j < 64; ++j) hB[i * 64 + j] = __float2half(i == j ? 1.f : 0.f);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  cuuint64_t dimsA[2] = {64, kRows}, strA[1] = {128}, dimsB[2] = {64, 64};
  cuuint32_t boxA[2] = {64, kRows}, boxB[2] = {64, 64}, es[2] = {1, 1};
  enc(&mA, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&mB, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = kRows * 128 + 64 * 128 + 1024 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 8, 10, 11, 21};
  const int sbos[] = {1024, 1280, 2048};
  for (int sbo : sbos)
    for (int shift : shifts)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        const int bo = bo_mode ? (shift & 7) : 0;
        if (bo_mode && bo == 0) continue;
        int bad_rows = 0, bad_cols = 0;
        for (int pass = 0; pass < 2; ++pass) {
          for (int r = 0; r < kRows; ++r)
            for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2half(pass == 0 ? static_cast<float>(r) : static_cast<float>(c));
          cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
          probe<<<1, 128, smem>>>(mA, mB, dOut, shift, sbo, bo);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m) {
            const int want_row = shift + (m / 8) * (sbo / 128) + (m % 8);
            for (int n = 0; n < 64; ++n) {
              const float want = pass == 0 ? static_cast<float>(want_row) : static_cast<float>(n);
              if (want_row < kRows && out[m * 64 + n] != want) (pass == 0 ? bad_rows : bad_cols)++;
            }
          }
        }
        printf("sbo %4d shift %2d base_offset %d : wrong-row elems %5d wrong-col elems %5d %s\n", sbo, shift, bo,
               bad_rows, bad_cols, (bad_rows | bad_cols) ? "" : "OK");
      }
  return 0;
}
