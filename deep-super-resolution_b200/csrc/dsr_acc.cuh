// dsr_acc.cuh -- order-independent accumulators.
//
// Every cross-CTA reduction of the DIP step (BatchNorm sums of the forward pass, the BatchNorm-backward sums, the
// gradients of the small 1x1 layers) is accumulated in 64-bit FIXED POINT: integer addition is associative, so the
// result does not depend on the order in which the CTAs' atomics land and two runs of one binary give bit-identical
// statistics.  (fp32 atomics gave sums that differed in the last bits from run to run; the freshly initialised
// network amplifies that into 1e-3 relative differences of the output and a few percent of the gradient.)
//
// Two scales: forward sums (sum x, sum x^2 of fp16 activations over up to 2^18 pixels) use 2^-20 quanta, range
// +-8.8e12; backward sums (S * gradient units, S = the dynamic loss scale) use 2^-24 quanta, range +-5.5e11.
// A contribution is rounded to the quantum once (round-to-nearest); non-finite contributions poison the slot
// with a huge value, which the gradient-scale logic then sees as an overflowed pass.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsr {

typedef unsigned long long acc_t;

constexpr float kAccFwdScale = 1048576.f;           // 2^20
constexpr float kAccFwdInv = 1.f / 1048576.f;
constexpr float kAccBwdScale = 16777216.f;          // 2^24
constexpr float kAccBwdInv = 1.f / 16777216.f;

#ifdef __CUDACC__
__device__ __forceinline__ acc_t acc_fix(float v, float scale) {
  // saturating conversion (cvt.rni.s64.f32 saturates; NaN -> 0)
  return static_cast<acc_t>(__float2ll_rn(v * scale));
}
// 64-bit integer -> float through two 32-bit conversions (I2F.S64 is a slow multi-instruction sequence and sits at the
// head of every consumer kernel's dependency chain): |value| = hi * 2^32 + lo
// of the magnitude (sign handled separately: the pieces of a small negative number would cancel catastrophically)
__device__ __forceinline__ float acc_val(acc_t a, float inv) {
  const long long sv = static_cast<long long>(a);
  const unsigned long long mag = static_cast<unsigned long long>(sv < 0 ? -sv : sv);
  const unsigned int hi = static_cast<unsigned int>(mag >> 32), lo = static_cast<unsigned int>(mag);
  const float m = fmaf(static_cast<float>(hi), 4294967296.f, static_cast<float>(lo)) * inv;
  return sv < 0 ? -m : m;
}

__device__ __forceinline__ void acc_add_f(acc_t* p, float v) { atomicAdd(p, acc_fix(v, kAccFwdScale)); }
__device__ __forceinline__ void acc_add_b(acc_t* p, float v) { atomicAdd(p, acc_fix(v, kAccBwdScale)); }
__device__ __forceinline__ float acc_get_f(const acc_t* p) { return acc_val(*p, kAccFwdInv); }
__device__ __forceinline__ float acc_get_b(const acc_t* p) { return acc_val(*p, kAccBwdInv); }
#endif

}  // namespace dsr
