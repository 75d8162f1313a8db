// dsr_debug.cu -- naive CUDA-core CHECKER kernels for the two tcgen05 kernels.  They evaluate the
// same implicit GEMMs from the same packed operands with plain loops (one thread per output
// element), reading through the 5-D coordinate arithmetic that the TMA descriptors encode.
// Used only by csrc/selftest.cu and by tests through dsr_plan_set_debug_conv(); never on the
// product path.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "dsr_debug.h"

namespace dsr {

__device__ __forceinline__ float ld16(const uint16_t* p, long long i, int bf16) {
  if (bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// element (c, px, x, py, y) of a padded NHWC tensor viewed with parity step; 0 outside.
__device__ __forceinline__ float act_at(const ActRef& a, int c, int px, int x, int py, int y) {
  const int nx = (a.Wp + a.step - 1) / a.step, ny = (a.Hp + a.step - 1) / a.step;
  if (x < 0 || y < 0 || x >= nx || y >= ny || c < 0 || c >= a.C) return 0.f;
  const long long off = (static_cast<long long>(py + a.step * y) * a.Wp + (px + a.step * x)) * a.C + c;
  if (off >= static_cast<long long>(a.Hp) * a.Wp * a.C) return 0.f;   // TMA bounds are per-dimension on the view
  return ld16(a.ptr, off, a.bf16);
}

__global__ void conv_ref_kernel(ConvGemmParams p, ActRef a, WgtRef b) {
  const long long total = static_cast<long long>(p.out_h) * p.out_w * p.n_store;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % p.n_store);
    const int x = static_cast<int>((i / p.n_store) % p.out_w);
    const int y = static_cast<int>(i / (static_cast<long long>(p.n_store) * p.out_w));
    float acc = 0.f;
    for (int k = 0; k < p.nkb; ++k) {
      const KBlk kb = p.kb[k];
      const int kc = kb.wide ? 64 : 16;
      for (int c = 0; c < kc; ++c) {
        const float av = act_at(a, kb.a_c + c, kb.a_px, x + kb.a_dx, kb.a_py, y + kb.a_dy);
        const int bk = kb.b_k + c;
        const float bv = (bk < b.K) ? ld16(b.ptr, static_cast<long long>(kb.b_row + n) * b.K + bk, b.bf16) : 0.f;
        acc = fmaf(av, bv, acc);
      }
    }
    const long long o = static_cast<long long>(y) * p.out_sy + static_cast<long long>(x) * p.out_sx + n;
    float stored;
    if (p.out_bf16) {
      const __nv_bfloat16 h = __float2bfloat16_rn(acc);
      reinterpret_cast<__nv_bfloat16*>(p.out)[o] = h;
      stored = __bfloat162float(h);
    } else {
      const __half h = __float2half_rn(acc);
      reinterpret_cast<__half*>(p.out)[o] = h;
      stored = __half2float(h);
    }
    if (p.stats != nullptr) {
      acc_add_f(&p.stats[n], stored);
      acc_add_f(&p.stats[p.n_mma + n], stored * stored);
    }
  }
}

__global__ void wgrad_ref_kernel(WgradParams p, ActRef dr, ActRef x) {
  const int ncols = p.n64 * 64 + p.n16 * 16;
  const long long total = static_cast<long long>(p.ngroups) * p.ntaps * 128 * ncols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % ncols);
    const int co = static_cast<int>((i / ncols) % 128);
    const int t = static_cast<int>((i / (static_cast<long long>(ncols) * 128)) % p.ntaps);
    const int g = static_cast<int>(i / (static_cast<long long>(ncols) * 128 * p.ntaps));
    const WgTap tp = p.taps[g][t];
    const int ci = (col < p.n64 * 64) ? col : p.c16_base + (col - p.n64 * 64);
    float acc = 0.f;
    for (int yy = 0; yy < p.pb_y * p.ph; ++yy)
      for (int xx = 0; xx < p.pb_x * p.pw; ++xx) {
        const float d = act_at(dr, co, 0, xx + 1, 0, yy + 1);
        if (d == 0.f) continue;
        acc = fmaf(d, act_at(x, ci, tp.px, xx + tp.dx, tp.py, yy + tp.dy), acc);
      }
    atomicAdd(&p.dw[(static_cast<long long>(tp.w_tap) * 128 + co) * p.ldw + col], acc);
  }
}

int launch_conv_ref(const ConvGemmParams& p, const ActRef& a, const WgtRef& b, cudaStream_t s) {
  const long long total = static_cast<long long>(p.out_h) * p.out_w * p.n_store;
  if (total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  conv_ref_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, a, b);
  return static_cast<int>(cudaGetLastError());
}

int launch_wgrad_ref(const WgradParams& p, const ActRef& dr, const ActRef& x, cudaStream_t s) {
  const long long total = static_cast<long long>(p.ngroups) * p.ntaps * 128 * (p.n64 * 64 + p.n16 * 16);
  if (total <= 0) return 0;
  long long blocks = (total + 127) / 128;
  if (blocks > 148 * 64) blocks = 148 * 64;
  wgrad_ref_kernel<<<static_cast<int>(blocks), 128, 0, s>>>(p, dr, x);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace dsr
