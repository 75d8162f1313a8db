// dsr_host.h -- internal host-side declarations shared by the translation units of libdsr_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsr_conv.cuh"

namespace dsr {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Function attributes (opt-in shared memory) are per device: "configured once" flags are kept per device ordinal so
// that one process can use the library on several GPUs.
constexpr int kMaxDevices = 64;
inline int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

int ensure_driver_api();
int make_act_map(CUtensorMap* m, const void* base, int elem_is_16bit, int C, int Wp, int Hp, int step, int box_c,
                 int box_w, int box_h);
int make_wgt_map(CUtensorMap* m, const void* base, int K, int rows, int box_k, int box_rows);
int launch_conv_gemm(const ConvGemmParams& p, int num_sms, cudaStream_t stream);
int launch_wgrad(const WgradParams& p, cudaStream_t stream);
int launch_wgrad_halo(const WgHaloParams& p, cudaStream_t stream);
int wgrad_cluster_capacity();
int launch_conv_halo(const HaloParams& p, int num_sms, cudaStream_t stream);
int halo_smem_bytes(int n_part, int n_wide, int n_narrow, int wide_slots, int ntaps, int slot_bytes = 0);

// Element formats of the 16-bit tensors (tcgen05 kind::f16 operand format codes).
enum : uint32_t { FMT_F16 = 0, FMT_BF16 = 1 };

}  // namespace dsr
