// dsr_metrics.cu -- on-device image-quality metrics for the logging branch of the DIP loop (DIP.py:71-87,183-185):
// PSNR and SSIM as single-launch bandwidth kernels that leave ONE float on the device, so that logging no longer
// costs a chain of library launches and host synchronisations per metric.
//
// Semantics follow the objects DIP.py constructs (torchmetrics 1.x, a third-party dependency that is absent from the
// reference checkout and from this image; its published algorithm is restated here; the tests hold a CPU restatement):
//   PeakSignalNoiseRatio()            data_range = max(target) - min(min(target), 0);  10 log10(range^2 / mean((p - t)^2))
//   StructuralSimilarityIndexMeasure  11 x 11 Gaussian window (sigma 1.5, normalised), k1 = 0.01, k2 = 0.03, statistics
//     (data_range = 1.)               E[p], E[t], E[pp], E[tt], E[pt] by the window, variances clamped at 0,
//                                     ssim = (2 mu_p mu_t + c1)(2 cov + c2) / ((mu_p^2 + mu_t^2 + c1)(var_p + var_t + c2)),
//                                     averaged over the pixels whose window lies inside the image (the reflect-padded
//                                     border of 5 pixels is cropped again, torchmetrics functional/image/ssim.py)
// Sums are accumulated in 64-bit fixed point (order-independent), the last block to arrive finalises.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/dsr_b200.h"
#include "dsr_acc.cuh"
#include "dsr_launch.cuh"

namespace dsr {
namespace {

constexpr float kMetScale = 4294967296.f;        // 2^32 quanta: sums of squared errors / ssim values up to 2^31
constexpr int kSsimK = 11, kSsimPad = 5;
constexpr int kTile = 32;                        // output tile (pixels); staged patch is (32 + 10)^2

struct MetricWs {                                // 32 bytes, zero between launches
  acc_t sum;
  unsigned int ticket;
  unsigned int tmin_bits, tmax_bits;             // orderable encodings of min / max target
  unsigned int pad_;
};

__device__ __forceinline__ unsigned int f2ord(float f) {      // monotone float -> uint map
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// ws->tmin_bits holds ~ord(min) so that zero-initialised memory is the identity of both reductions
__global__ void psnr_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n,
                            float data_range, MetricWs* ws, float* __restrict__ out) {
  pdl_sync();
  float se = 0.f, tmin = INFINITY, tmax = -INFINITY;
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(pred) + i);
    const float4 t = __ldg(reinterpret_cast<const float4*>(target) + i);
    const float d0 = p.x - t.x, d1 = p.y - t.y, d2 = p.z - t.z, d3 = p.w - t.w;
    se += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    tmin = fminf(tmin, fminf(fminf(t.x, t.y), fminf(t.z, t.w)));
    tmax = fmaxf(tmax, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float d = pred[i] - target[i];
      se += d * d;
      tmin = fminf(tmin, target[i]);
      tmax = fmaxf(tmax, target[i]);
    }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, d);
    tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, d));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ws->sum, acc_fix(se, kMetScale));
    atomicMax(&ws->tmin_bits, ~f2ord(tmin));
    atomicMax(&ws->tmax_bits, f2ord(tmax));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      const float sse = acc_val(atomicExch(&ws->sum, 0ull), 1.f / kMetScale);
      const float mn = ord2f(~atomicExch(&ws->tmin_bits, 0u)), mx = ord2f(atomicExch(&ws->tmax_bits, 0u));
      ws->ticket = 0u;
      const float range = data_range > 0.f ? data_range : fmaxf(mx, 0.f) - fminf(mn, 0.f);
      const float mse = sse / static_cast<float>(n);
      *out = 10.f * (2.f * log10f(range) - log10f(mse));
    }
  }
}

__constant__ float c_gauss[kSsimK];

// one block = one 32 x 32 tile of interior pixels of one plane; separable window: rows, then columns
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                   int planes, int H, int W, float c1, float c2, MetricWs* ws,
                                                   float* __restrict__ out) {
  pdl_sync();
  constexpr int P = kTile + 2 * kSsimPad;        // 42
  __shared__ float sp[P][P + 1], st[P][P + 1];
  __shared__ float rowf[5][P][kTile + 1];        // row-filtered p, t, pp, tt, pt
  const int ih = H - 2 * kSsimPad, iw = W - 2 * kSsimPad;          // interior extent
  const int tiles_x = (iw + kTile - 1) / kTile, tiles_y = (ih + kTile - 1) / kTile;
  const int plane = blockIdx.x / (tiles_x * tiles_y);
  const int tt = blockIdx.x % (tiles_x * tiles_y);
  const int y0 = (tt / tiles_x) * kTile, x0 = (tt % tiles_x) * kTile;   // interior coordinates = image coords of the window's corner
  const float* pp = pred + static_cast<long long>(plane) * H * W;
  const float* tp = target + static_cast<long long>(plane) * H * W;
  for (int i = threadIdx.x; i < P * P; i += blockDim.x) {
    const int r = i / P, c = i - r * P;
    const int y = min(y0 + r, H - 1), x = min(x0 + c, W - 1);       // clamped reads only feed masked outputs
    sp[r][c] = __ldg(pp + static_cast<long long>(y) * W + x);
    st[r][c] = __ldg(tp + static_cast<long long>(y) * W + x);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P * kTile; i += blockDim.x) {
    const int r = i / kTile, c = i - r * kTile;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimK; ++k) {
      const float g = c_gauss[k], p = sp[r][c + k], t = st[r][c + k];
      a = fmaf(g, p, a);
      b = fmaf(g, t, b);
      aa = fmaf(g, p * p, aa);
      bb = fmaf(g, t * t, bb);
      ab = fmaf(g, p * t, ab);
    }
    rowf[0][r][c] = a; rowf[1][r][c] = b; rowf[2][r][c] = aa; rowf[3][r][c] = bb; rowf[4][r][c] = ab;
  }
  __syncthreads();
  float local = 0.f;
  for (int i = threadIdx.x; i < kTile * kTile; i += blockDim.x) {
    const int r = i / kTile, c = i - r * kTile;
    if (y0 + r >= ih || x0 + c >= iw) continue;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kSsimK; ++k) {
      const float g = c_gauss[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(g, rowf[q][r + k][c], m[q]);
    }
    const float mu_p2 = m[0] * m[0], mu_t2 = m[1] * m[1], mu_pt = m[0] * m[1];
    const float var_p = fmaxf(m[2] - mu_p2, 0.f), var_t = fmaxf(m[3] - mu_t2, 0.f), cov = m[4] - mu_pt;
    local += ((2.f * mu_pt + c1) * (2.f * cov + c2)) / ((mu_p2 + mu_t2 + c1) * (var_p + var_t + c2));
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0) atomicAdd(&ws->sum, acc_fix(local, kMetScale));
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      const float total = acc_val(atomicExch(&ws->sum, 0ull), 1.f / kMetScale);
      ws->ticket = 0u;
      *out = total / (static_cast<float>(planes) * ih * iw);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// eval_GAN.py:50-53 image writer path: resolved image [B][C][H][W] fp32 -> [B][H][W][C] uint8 ON THE DEVICE, so that a
// quarter of the bytes crosses PCIe and the host only encodes the PNG.
//   clip = 0: (x.transpose(1, 2, 0) * 255).astype(np.uint8) of eval_GAN.py:52 -- numpy's cast truncates towards zero and
//             keeps the low byte (the generator ends in tanh: negative values wrap, -127.5 -> 129), NaN -> 0
//   clip = 1: np.clip(x * 255, 0, 255).astype(np.uint8) of utils/common.py:81 (np_to_pil)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int to_u8(float v, int clip) {
  float t = v * 255.f;
  if (clip) t = fminf(fmaxf(t, 0.f), 255.f);            // fmaxf(NaN, 0) = 0, as np.clip propagates NaN -> cast 0
  return static_cast<unsigned int>(__float2int_rz(t)) & 0xFFu;
}
__global__ void image_to_u8_hwc_kernel(const float* __restrict__ chw, int C, long long hw, int clip,
                                       unsigned char* __restrict__ out) {
  pdl_sync();
  const float* src = chw + static_cast<long long>(blockIdx.y) * C * hw;
  unsigned char* dst = out + static_cast<long long>(blockIdx.y) * C * hw;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  if (C == 3 && (hw & 3) == 0) {
    // 4 pixels per thread: one float4 per plane in, three 32-bit words (12 bytes) out
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < (hw >> 2); q += stride) {
      float4 v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __ldg(reinterpret_cast<const float4*>(src + c * hw) + q);
      const float f[12] = {v[0].x, v[1].x, v[2].x, v[0].y, v[1].y, v[2].y, v[0].z, v[1].z, v[2].z, v[0].w, v[1].w, v[2].w};
      unsigned int w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        w[k] = to_u8(f[4 * k], clip) | (to_u8(f[4 * k + 1], clip) << 8) | (to_u8(f[4 * k + 2], clip) << 16) |
               (to_u8(f[4 * k + 3], clip) << 24);
      unsigned int* o = reinterpret_cast<unsigned int*>(dst) + q * 3;
      o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
    }
    return;
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw; i += stride)
    for (int c = 0; c < C; ++c) dst[i * C + c] = static_cast<unsigned char>(to_u8(src[c * hw + i], clip));
}

}  // namespace
}  // namespace dsr

using namespace dsr;

extern "C" {

size_t dsr_metric_workspace_bytes(void) { return sizeof(MetricWs); }

int dsr_psnr(const float* pred, const float* target, long long n, float data_range, void* workspace, float* out,
             void* stream) {
  if (!pred || !target || !workspace || !out || n < 1) return -1;
  if ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) return -3;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  launch_k(psnr_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), pred, target,
           n, data_range, static_cast<MetricWs*>(workspace), out);
  return static_cast<int>(cudaGetLastError());
}

int dsr_ssim(const float* pred, const float* target, int planes, int H, int W, float data_range, void* workspace,
             float* out, void* stream) {
  if (!pred || !target || !workspace || !out || planes < 1 || data_range <= 0.f) return -1;
  if (H <= 2 * kSsimPad || W <= 2 * kSsimPad) return -1;
  static bool table_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !table_dev[dev]) {
    float g[kSsimK];
    double s = 0.0;
    for (int i = 0; i < kSsimK; ++i) {             // torchmetrics _gaussian(): exp(-(d / sigma)^2 / 2), d = -5 .. 5, sigma 1.5
      const double d = (i - kSsimPad) / 1.5;
      g[i] = static_cast<float>(exp(-0.5 * d * d));
      s += g[i];
    }
    for (int i = 0; i < kSsimK; ++i) g[i] = static_cast<float>(g[i] / s);
    cudaError_t e = cudaMemcpyToSymbol(c_gauss, g, sizeof(g));
    if (e != cudaSuccess) return static_cast<int>(e);
    table_dev[dev] = true;
  }
  const int ih = H - 2 * kSsimPad, iw = W - 2 * kSsimPad;
  const int tiles = ((iw + kTile - 1) / kTile) * ((ih + kTile - 1) / kTile);
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  launch_k(ssim_kernel, dim3(static_cast<unsigned>(planes * tiles)), dim3(256), 0, static_cast<cudaStream_t>(stream), pred,
           target, planes, H, W, c1, c2, static_cast<MetricWs*>(workspace), out);
  return static_cast<int>(cudaGetLastError());
}


int dsr_image_to_u8_hwc(const float* chw, int B, int C, int H, int W, int clip, unsigned char* out_hwc, void* stream) {
  if (!chw || !out_hwc || B < 1 || C < 1 || H < 1 || W < 1) return -1;
  const long long hw = static_cast<long long>(H) * W;
  if ((reinterpret_cast<uintptr_t>(chw) & 15) || (reinterpret_cast<uintptr_t>(out_hwc) & 3)) return -3;
  long long blocks = ((C == 3 && (hw & 3) == 0 ? hw / 4 : hw) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(image_to_u8_hwc_kernel, dim3(static_cast<unsigned>(blocks), static_cast<unsigned>(B)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), chw, C, hw, clip, out_hwc);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
