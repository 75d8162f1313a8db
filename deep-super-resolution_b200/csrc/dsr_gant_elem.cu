// dsr_gant_elem.cu -- bandwidth and CUDA-core kernels of the SRGAN training step (train_GAN.py:38-71): everything of
// the generator (models/GAN/generator.py), discriminator (models/GAN/discriminator.py) and perceptual loss
// (utils/GAN.py:6-123) that is not a 64..512-channel convolution: train-mode BatchNorm forward / backward, PReLU,
// LeakyReLU, PixelShuffle, tanh, the VGG preprocessing transform and its transpose, MaxPool, the feature MSE, the
// dense head of the discriminator, BCE, weight packing and the weight gradients of the 3-channel layers.
// Tensor convention: dsr_gant.cuh (bf16 NHWC tall grid, gap rows never written).
#include "dsr_gant_elem.cuh"
#include <cuda_fp16.h>

#include "dsr_host.h"
#include "dsr_launch.cuh"

namespace dsr {

namespace {

constexpr int kT = 256;
constexpr float kLrelu = 0.2f;

#define GL_CHECK() return static_cast<int>(cudaGetLastError())

__device__ __forceinline__ long long tg_pix_off(const TG& g, long long q) {
  const int hw = g.H * g.W;
  const int b = static_cast<int>(q / hw);
  const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
  const int y = r / g.W, x = r - y * g.W;
  return (static_cast<long long>(b * g.P + y) * g.W + x) * g.C;
}
__device__ __forceinline__ void load8(const bf16_t* p, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8(bf16_t* p, const float (&f)[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}
__device__ __forceinline__ void load8h(const bf16_t* p, float (&f)[8]) {       // the same 16 bytes read as IEEE half
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8h(bf16_t* p, const float (&f)[8]) {
  uint4 q;
  __half2* h = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}
inline int grid_for(long long items, int per_sm = 4) {
  long long b = (items + kT - 1) / kT;
  const long long cap = 148LL * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// Threads of a block that share `group = tid % ngroups` add their NV partials: v -> sm[tid][NV]; afterwards
// group_sum(sm, ngroups, g, i) is the block's sum of value i of group g.
template <int NV>
__device__ __forceinline__ void group_store(const float (&v)[NV], float* sm) {
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) sm[threadIdx.x * NV + i] = v[i];
  __syncthreads();
}
template <int NV>
__device__ __forceinline__ float group_sum(const float* sm, int ngroups, int g, int i) {
  float s = 0.f;
  for (int t = g; t < kT; t += ngroups) s += sm[t * NV + i];
  return s;
}

// ---------------------------------------------------------------------------------------------
// images
// ---------------------------------------------------------------------------------------------
__global__ void g_pack_image_kernel(const float* __restrict__ in, bf16_t* __restrict__ out, TG g) {
  pdl_sync();
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  const long long hw = static_cast<long long>(g.H) * g.W;
  for (long long q = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; q < np; q += static_cast<long long>(gridDim.x) * kT) {
    const long long b = q / hw, r = q - b * hw;
    float f[8] = {in[(b * 3 + 0) * hw + r], in[(b * 3 + 1) * hw + r], in[(b * 3 + 2) * hw + r], 0.f, 0.f, 0.f, 0.f, 0.f};
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bf16_t* o = out + tg_pix_off(g, q);
    store8(o, f);
    store8(o + 8, z);
  }
}

__global__ void g_tanh_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, bf16_t* __restrict__ dz,
                                  float* __restrict__ dbias3, TG g) {
  __shared__ float sm[kT * 3];
  pdl_sync();
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  const long long hw = static_cast<long long>(g.H) * g.W;
  float acc[3] = {0.f, 0.f, 0.f};
  for (long long q = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; q < np; q += static_cast<long long>(gridDim.x) * kT) {
    const long long b = q / hw, r = q - b * hw;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float o = out[(b * 3 + c) * hw + r];
      f[c] = dout[(b * 3 + c) * hw + r] * (1.f - o * o);
    }
    const float zz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bf16_t* o = dz + tg_pix_off(g, q);
    store8(o, f);
    store8(o + 8, zz);
    // the bias gradient sums the values the weight-gradient kernels see (bf16-rounded)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += __bfloat162float(__float2bfloat16_rn(f[c]));
  }
  group_store<3>(acc, sm);
  if (threadIdx.x < 3) atomicAdd(&dbias3[threadIdx.x], group_sum<3>(sm, 1, 0, threadIdx.x));
}

// ---------------------------------------------------------------------------------------------
// BatchNorm (train mode: statistics of the batch, biased variance, eps 1e-5)
// ---------------------------------------------------------------------------------------------
struct BnTabG { float mean[512], rstd[512], sc[512], sh[512]; };

__device__ __forceinline__ void bn_tab_fill_g(BnTabG& t, const double* stats, const float* gamma, const float* beta, int C,
                                              double inv_count) {
  for (int c = threadIdx.x; c < C; c += kT) {
    const double m = stats[c] * inv_count;
    double var = stats[C + c] * inv_count - m * m;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + 1e-5));
    const float sc = gamma[c] * rstd;
    t.mean[c] = static_cast<float>(m);
    t.rstd[c] = rstd;
    t.sc[c] = sc;
    t.sh[c] = beta[c] - static_cast<float>(m) * sc;
  }
  __syncthreads();
}

template <int ACT>
__global__ void g_bn_apply_kernel(const bf16_t* __restrict__ raw, bf16_t* __restrict__ out, const bf16_t* __restrict__ res,
                                  const double* __restrict__ stats, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, const float* __restrict__ slope_p, TG g, double inv_count,
                                  const float* __restrict__ conv_bias, float* __restrict__ rm, float* __restrict__ rv,
                                  int run_times) {
  __shared__ BnTabG tab;
  pdl_sync();
  bn_tab_fill_g(tab, stats, gamma, beta, g.C, inv_count);
  if (blockIdx.x == 0 && rm != nullptr && run_times > 0) {
    // running_mean / running_var as torch.nn.BatchNorm2d updates them (momentum 0.1, unbiased variance), `run_times`
    // times with the same batch statistics; the conv bias dropped by the conv kernels is added back to the mean
    const double count = 1.0 / inv_count;
    for (int c = threadIdx.x; c < g.C; c += kT) {
      const double m = stats[c] * inv_count;
      double var = stats[g.C + c] * inv_count - m * m;
      if (var < 0.0) var = 0.0;
      const float mean = static_cast<float>(m) + (conv_bias ? conv_bias[c] : 0.f);
      const float uvar = static_cast<float>(var * (count / (count - 1.0)));
      float a = rm[c], b = rv[c];
      for (int t = 0; t < run_times; ++t) {
        a = 0.9f * a + 0.1f * mean;
        b = 0.9f * b + 0.1f * uvar;
      }
      rm[c] = a;
      rv[c] = b;
    }
  }
  const float slope = (ACT == GACT_PRELU) ? slope_p[0] : kLrelu;
  const int g8 = g.C >> 3;
  const long long total = static_cast<long long>(g.B) * g.H * g.W * g8;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i / g8;
    const int c0 = static_cast<int>(i - q * g8) * 8;
    const long long off = tg_pix_off(g, q) + c0;
    float x[8];
    load8(raw + off, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = x[j] * tab.sc[c0 + j] + tab.sh[c0 + j];
      if (ACT != GACT_NONE) y = y > 0.f ? y : y * slope;
      x[j] = y;
    }
    if (res != nullptr) {
      float r[8];
      load8(res + off, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += r[j];
    }
    store8(out + off, x);
  }
}

// pass 1 of the backward: sums[c] = sum g, sums[C + c] = sum g * xhat, sums[2 C] = d(loss)/d(PReLU slope),
// with g = dy * act'(z), z = gamma * xhat + beta
template <int ACT>
__global__ void g_bn_bwd_stats_kernel(const bf16_t* __restrict__ dy, const bf16_t* __restrict__ raw,
                                      const double* __restrict__ stats, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ slope_p, TG g,
                                      double inv_count, double* __restrict__ sums) {
  __shared__ BnTabG tab;
  __shared__ float sm[kT * 17];
  pdl_sync();
  bn_tab_fill_g(tab, stats, gamma, beta, g.C, inv_count);
  const float slope = (ACT == GACT_PRELU) ? slope_p[0] : kLrelu;
  const int g8 = g.C >> 3;
  const int cg = threadIdx.x % g8;
  const int c0 = cg * 8;
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  const long long pstride = static_cast<long long>(gridDim.x) * kT / g8;
  float v[17];
#pragma unroll
  for (int j = 0; j < 17; ++j) v[j] = 0.f;
  for (long long q = (static_cast<long long>(blockIdx.x) * kT + threadIdx.x) / g8; q < np; q += pstride) {
    const long long off = tg_pix_off(g, q) + c0;
    float x[8], d[8];
    load8(raw + off, x);
    load8(dy + off, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - tab.mean[c0 + j]) * tab.rstd[c0 + j];
      float gj = d[j];
      if (ACT != GACT_NONE) {
        const float z = x[j] * tab.sc[c0 + j] + tab.sh[c0 + j];
        if (!(z > 0.f)) {
          if (ACT == GACT_PRELU) v[16] += d[j] * z;
          gj *= slope;
        }
      }
      v[j] += gj;
      v[8 + j] += gj * xh;
    }
  }
  group_store<17>(v, sm);
  for (int e = threadIdx.x; e < g8 * 16; e += kT) {
    const int gg = e / 16, i = e % 16;
    const float s = group_sum<17>(sm, g8, gg, i);
    atomicAdd(&sums[(i < 8 ? 0 : g.C) + gg * 8 + (i & 7)], static_cast<double>(s));
  }
  if (ACT == GACT_PRELU && threadIdx.x == 0) {
    float s = 0.f;
    for (int t = 0; t < kT; ++t) s += sm[t * 17 + 16];
    atomicAdd(&sums[2 * g.C], static_cast<double>(s));
  }
}

template <int ACT>
__global__ void g_bn_bwd_apply_kernel(const bf16_t* __restrict__ dy, const bf16_t* __restrict__ raw, bf16_t* __restrict__ draw,
                                      const double* __restrict__ stats, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ slope_p, TG g,
                                      double inv_count, const double* __restrict__ sums, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, float* __restrict__ dslope, double* __restrict__ next_sums) {
  __shared__ BnTabG tab;
  __shared__ float s_mg[512], s_mgx[512];
  pdl_sync();
  // the OTHER half of the ping-pong scratch is cleared here for the next BatchNorm backward of this network (its last
  // reader, the previous apply kernel, completed before this call's statistics kernel started; its next writer starts
  // after this kernel): no memset node between the kernels, so the programmatic launch overlap of the chain survives
  if (blockIdx.x == 0 && next_sums != nullptr)
    for (int c = threadIdx.x; c < kBnSumsHalf; c += kT) next_sums[c] = 0.0;      // the whole half: the next layer's C may be larger
  for (int c = threadIdx.x; c < g.C; c += kT) {
    s_mg[c] = static_cast<float>(sums[c] * inv_count);
    s_mgx[c] = static_cast<float>(sums[g.C + c] * inv_count);
    if (blockIdx.x == 0) {               // parameter gradients (+=): d gamma = sum g xhat, d beta = sum g
      dgamma[c] += static_cast<float>(sums[g.C + c]);
      dbeta[c] += static_cast<float>(sums[c]);
    }
  }
  if (ACT == GACT_PRELU && blockIdx.x == 0 && threadIdx.x == 0 && dslope != nullptr)
    dslope[0] += static_cast<float>(sums[2 * g.C]);
  bn_tab_fill_g(tab, stats, gamma, beta, g.C, inv_count);
  const float slope = (ACT == GACT_PRELU) ? slope_p[0] : kLrelu;
  const int g8 = g.C >> 3;
  const long long total = static_cast<long long>(g.B) * g.H * g.W * g8;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i / g8;
    const int c0 = static_cast<int>(i - q * g8) * 8;
    const long long off = tg_pix_off(g, q) + c0;
    float x[8], d[8];
    load8(raw + off, x);
    load8(dy + off, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (x[j] - tab.mean[c0 + j]) * tab.rstd[c0 + j];
      float gj = d[j];
      if (ACT != GACT_NONE) {
        const float z = x[j] * tab.sc[c0 + j] + tab.sh[c0 + j];
        if (!(z > 0.f)) gj *= slope;
      }
      x[j] = tab.sc[c0 + j] * (gj - s_mg[c0 + j] - xh * s_mgx[c0 + j]);
    }
    store8(draw + off, x);
  }
}

// ---------------------------------------------------------------------------------------------
// PixelShuffle(2) + PReLU (generator.py:27-41): u[c][2y+dy][2x+dx] = prelu(s[4c + 2dy + dx][y][x])
// ---------------------------------------------------------------------------------------------
__global__ void g_shuffle_fwd_kernel(const bf16_t* __restrict__ s, bf16_t* __restrict__ u, const float* __restrict__ slope_p,
                                     TG gs, TG gu) {
  pdl_sync();
  const float a = slope_p[0];
  const long long total = static_cast<long long>(gs.B) * gs.H * gs.W * 8;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i >> 3;
    const int gi = static_cast<int>(i & 7);
    const int hw = gs.H * gs.W;
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
    const int y = r / gs.W, x = r - y * gs.W;
    const bf16_t* sp = s + (static_cast<long long>(b * gs.P + y) * gs.W + x) * gs.C + gi * 32;
    float f[32];
#pragma unroll
    for (int k = 0; k < 4; ++k) load8(sp + 8 * k, *reinterpret_cast<float(*)[8]>(&f[8 * k]));
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = f[4 * k + pos];
        o[k] = v > 0.f ? v : v * a;
      }
      const int oy = 2 * y + (pos >> 1), ox = 2 * x + (pos & 1);
      store8(u + (static_cast<long long>(b * gu.P + oy) * gu.W + ox) * gu.C + gi * 8, o);
    }
  }
}

__global__ void g_shuffle_bwd_kernel(const bf16_t* __restrict__ du, const bf16_t* __restrict__ s, bf16_t* __restrict__ ds,
                                     const float* __restrict__ slope_p, float* __restrict__ dbias, float* __restrict__ dslope,
                                     TG gs, TG gu) {
  __shared__ float sm[kT * 33];
  pdl_sync();
  const float a = slope_p[0];
  const int gi = threadIdx.x & 7;
  const long long np = static_cast<long long>(gs.B) * gs.H * gs.W;
  float v[33];
#pragma unroll
  for (int j = 0; j < 33; ++j) v[j] = 0.f;
  for (long long q = (static_cast<long long>(blockIdx.x) * kT + threadIdx.x) >> 3; q < np;
       q += (static_cast<long long>(gridDim.x) * kT) >> 3) {
    const int hw = gs.H * gs.W;
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
    const int y = r / gs.W, x = r - y * gs.W;
    const long long soff = (static_cast<long long>(b * gs.P + y) * gs.W + x) * gs.C + gi * 32;
    float f[32], o[32];
#pragma unroll
    for (int k = 0; k < 4; ++k) load8(s + soff + 8 * k, *reinterpret_cast<float(*)[8]>(&f[8 * k]));
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float d[8];
      const int oy = 2 * y + (pos >> 1), ox = 2 * x + (pos & 1);
      load8(du + (static_cast<long long>(b * gu.P + oy) * gu.W + ox) * gu.C + gi * 8, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float sv = f[4 * k + pos];
        float gsv = d[k];
        if (!(sv > 0.f)) { v[32] += d[k] * sv; gsv *= a; }
        o[4 * k + pos] = gsv;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) store8(ds + soff + 8 * k, *reinterpret_cast<float(*)[8]>(&o[8 * k]));
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __bfloat162float(__float2bfloat16_rn(o[j]));
  }
  group_store<33>(v, sm);
  for (int e = threadIdx.x; e < 256; e += kT) atomicAdd(&dbias[e], group_sum<33>(sm, 8, e >> 5, e & 31));
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kT; ++k) t += sm[k * 33 + 32];
    atomicAdd(dslope, t);
  }
}

__global__ void g_prelu_fwd_kernel(const bf16_t* __restrict__ z, bf16_t* __restrict__ out, const float* __restrict__ slope_p,
                                   TG g) {
  pdl_sync();
  const float a = slope_p[0];
  const int g8 = g.C >> 3;
  const long long total = static_cast<long long>(g.B) * g.H * g.W * g8;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i / g8;
    const long long off = tg_pix_off(g, q) + static_cast<int>(i - q * g8) * 8;
    float x[8];
    load8(z + off, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = x[j] > 0.f ? x[j] : x[j] * a;
    store8(out + off, x);
  }
}

__global__ void g_prelu_bwd_kernel(const bf16_t* __restrict__ d1, const bf16_t* __restrict__ d2, const bf16_t* __restrict__ z,
                                   bf16_t* __restrict__ dz, const float* __restrict__ slope_p, float* __restrict__ dbias,
                                   float* __restrict__ dslope, TG g) {
  __shared__ float sm[kT * 9];
  pdl_sync();
  const float a = slope_p[0];
  const int g8 = g.C >> 3;
  const int cg = threadIdx.x % g8;
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  float v[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) v[j] = 0.f;
  for (long long q = (static_cast<long long>(blockIdx.x) * kT + threadIdx.x) / g8; q < np;
       q += static_cast<long long>(gridDim.x) * kT / g8) {
    const long long off = tg_pix_off(g, q) + cg * 8;
    float x[8], d[8];
    load8(z + off, x);
    load8(d1 + off, d);
    if (d2 != nullptr) {
      float e[8];
      load8(d2 + off, e);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] += e[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!(x[j] > 0.f)) { v[8] += d[j] * x[j]; d[j] *= a; }
    }
    store8(dz + off, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __bfloat162float(__float2bfloat16_rn(d[j]));
  }
  group_store<9>(v, sm);
  for (int e = threadIdx.x; e < g.C; e += kT) atomicAdd(&dbias[e], group_sum<9>(sm, g8, e >> 3, e & 7));
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kT; ++k) t += sm[k * 9 + 8];
    atomicAdd(dslope, t);
  }
}

__global__ void g_chan_sum_kernel(const bf16_t* __restrict__ t, float* __restrict__ out, TG g) {
  __shared__ float sm[kT * 8];
  pdl_sync();
  const int g8 = g.C >> 3;
  const int cg = threadIdx.x % g8;
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  for (long long q = (static_cast<long long>(blockIdx.x) * kT + threadIdx.x) / g8; q < np;
       q += static_cast<long long>(gridDim.x) * kT / g8) {
    float x[8];
    load8(t + tg_pix_off(g, q) + cg * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += x[j];
  }
  group_store<8>(v, sm);
  for (int e = threadIdx.x; e < g.C; e += kT) atomicAdd(&out[e], group_sum<8>(sm, g8, e >> 3, e & 7));
}

// ---------------------------------------------------------------------------------------------
// VGG19 preprocessing (utils/GAN.py:76-77: VGG19_Weights.IMAGENET1K_V1.transforms() = torchvision ImageClassification:
// bilinear resize of the smaller edge to 256 (antialias has no effect when enlarging), centre crop 224, normalise)
// ---------------------------------------------------------------------------------------------
__constant__ float kVggMean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float kVggStd[3] = {0.229f, 0.224f, 0.225f};

__device__ __forceinline__ void bil_src(int d, float scale, int n, int* i0, int* i1, float* l1) {
  float src = (static_cast<float>(d) + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  int a = static_cast<int>(src);
  if (a > n - 1) a = n - 1;
  *i0 = a;
  *i1 = a + 1 < n ? a + 1 : n - 1;
  *l1 = src - static_cast<float>(a);
}

__global__ void g_vgg_pre_fwd_kernel(const float* __restrict__ img, bf16_t* __restrict__ out, TG g, int Hi, int Wi, int Hr,
                                     int Wr, int top, int left, int f16) {
  pdl_sync();
  const float sy = static_cast<float>(Hi) / static_cast<float>(Hr), sx = static_cast<float>(Wi) / static_cast<float>(Wr);
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  const int hw = g.H * g.W;
  for (long long q = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; q < np; q += static_cast<long long>(gridDim.x) * kT) {
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
    const int y = r / g.W, x = r - y * g.W;
    int y0, y1, x0, x1;
    float ly, lx;
    bil_src(y + top, sy, Hi, &y0, &y1, &ly);
    bil_src(x + left, sx, Wi, &x0, &x1, &lx);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = img + (static_cast<long long>(b) * 3 + c) * Hi * Wi;
      const float v = (1.f - ly) * ((1.f - lx) * pl[y0 * Wi + x0] + lx * pl[y0 * Wi + x1]) +
                      ly * ((1.f - lx) * pl[y1 * Wi + x0] + lx * pl[y1 * Wi + x1]);
      f[c] = (v - kVggMean[c]) / kVggStd[c];
    }
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bf16_t* o = out + tg_pix_off(g, q);
    if (f16) store8h(o, f);
    else store8(o, f);
    store8(o + 8, z);
  }
}

// transpose of the above in gather form: one thread per SOURCE pixel sums the crop-window pixels that read it
__global__ void g_vgg_pre_bwd_kernel(const float* __restrict__ dpre, float* __restrict__ dimg, TG g, int Hi, int Wi, int Hr,
                                     int Wr, int top, int left, int accumulate) {
  pdl_sync();
  const float sy = static_cast<float>(Hi) / static_cast<float>(Hr), sx = static_cast<float>(Wi) / static_cast<float>(Wr);
  const long long np = static_cast<long long>(g.B) * Hi * Wi;
  for (long long q = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; q < np; q += static_cast<long long>(gridDim.x) * kT) {
    const int b = static_cast<int>(q / (Hi * Wi));
    const int r = static_cast<int>(q - static_cast<long long>(b) * Hi * Wi);
    const int py = r / Wi, px = r - py * Wi;
    int dy_lo = static_cast<int>(floorf((py - 0.5f) / sy - 0.5f)) - 1, dy_hi = static_cast<int>(ceilf((py + 1.5f) / sy - 0.5f)) + 1;
    int dx_lo = static_cast<int>(floorf((px - 0.5f) / sx - 0.5f)) - 1, dx_hi = static_cast<int>(ceilf((px + 1.5f) / sx - 0.5f)) + 1;
    if (dy_lo < top) dy_lo = top;
    if (dx_lo < left) dx_lo = left;
    if (dy_hi > top + g.H - 1) dy_hi = top + g.H - 1;
    if (dx_hi > left + g.W - 1) dx_hi = left + g.W - 1;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int dy = dy_lo; dy <= dy_hi; ++dy) {
      int y0, y1;
      float ly;
      bil_src(dy, sy, Hi, &y0, &y1, &ly);
      const float wy = (y0 == py ? 1.f - ly : 0.f) + (y1 == py ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int dx = dx_lo; dx <= dx_hi; ++dx) {
        int x0, x1;
        float lx;
        bil_src(dx, sx, Wi, &x0, &x1, &lx);
        const float wx = (x0 == px ? 1.f - lx : 0.f) + (x1 == px ? lx : 0.f);
        if (wx == 0.f) continue;
        const float4 d = *reinterpret_cast<const float4*>(
            dpre + (static_cast<long long>(b * g.P + (dy - top)) * g.W + (dx - left)) * g.C);
        acc[0] += wy * wx * d.x;
        acc[1] += wy * wx * d.y;
        acc[2] += wy * wx * d.z;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float* o = dimg + (static_cast<long long>(b) * 3 + c) * Hi * Wi + r;
      const float v = acc[c] / kVggStd[c];
      *o = accumulate ? *o + v : v;
    }
  }
}

__global__ void g_maxpool_fwd_kernel(const bf16_t* __restrict__ in, bf16_t* __restrict__ out, TG gi, TG go) {
  pdl_sync();
  const int g8 = go.C >> 3;
  const long long total = static_cast<long long>(go.B) * go.H * go.W * g8;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i / g8;
    const int c0 = static_cast<int>(i - q * g8) * 8;
    const int hw = go.H * go.W;
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
    const int y = r / go.W, x = r - y * go.W;
    float m[8];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float f[8];
      load8(in + (static_cast<long long>(b * gi.P + 2 * y + (pos >> 1)) * gi.W + 2 * x + (pos & 1)) * gi.C + c0, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = (pos == 0) ? f[j] : fmaxf(m[j], f[j]);
    }
    store8(out + (static_cast<long long>(b * go.P + y) * go.W + x) * go.C + c0, m);
  }
}

// dz_in = (first position of the window maximum ? dout : 0) * [y_in > 0]   (MaxPool2d backward + the ReLU in front of it)
__global__ void g_maxpool_bwd_kernel(const bf16_t* __restrict__ dout, const bf16_t* __restrict__ yin, bf16_t* __restrict__ dzin,
                                     TG gi, TG go) {
  pdl_sync();
  const int g8 = go.C >> 3;
  const long long total = static_cast<long long>(go.B) * go.H * go.W * g8;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i / g8;
    const int c0 = static_cast<int>(i - q * g8) * 8;
    const int hw = go.H * go.W;
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
    const int y = r / go.W, x = r - y * go.W;
    float f[4][8], d[8];
    int arg[8];
    float m[8];
    load8(dout + (static_cast<long long>(b * go.P + y) * go.W + x) * go.C + c0, d);
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      load8(yin + (static_cast<long long>(b * gi.P + 2 * y + (pos >> 1)) * gi.W + 2 * x + (pos & 1)) * gi.C + c0, f[pos]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (pos == 0 || f[pos][j] > m[j]) { m[j] = f[pos][j]; arg[j] = pos; }
      }
    }
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (arg[j] == pos && f[pos][j] > 0.f) ? d[j] : 0.f;
      store8(dzin + (static_cast<long long>(b * gi.P + 2 * y + (pos >> 1)) * gi.W + 2 * x + (pos & 1)) * gi.C + c0, o);
    }
  }
}

// loss_acc += sum (f1 - f2)^2 ;  dz = 2 (f1 - f2) / N * [f1 > 0]   (nn.MSELoss on the relu5_4 maps + the ReLU in front)
__global__ void g_feat_mse_kernel(const bf16_t* __restrict__ f1, const bf16_t* __restrict__ f2, bf16_t* __restrict__ dz, TG g,
                                  float two_over_n, double* __restrict__ loss_acc, int f16) {
  __shared__ float sm[kT];
  pdl_sync();
  const int g8 = g.C >> 3;
  const long long total = static_cast<long long>(g.B) * g.H * g.W * g8;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const long long q = i / g8;
    const long long off = tg_pix_off(g, q) + static_cast<int>(i - q * g8) * 8;
    float a[8], b[8];
    if (f16) { load8h(f1 + off, a); load8h(f2 + off, b); }
    else { load8(f1 + off, a); load8(f2 + off, b); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = a[j] - b[j];
      acc += d * d;
      b[j] = a[j] > 0.f ? d * two_over_n : 0.f;
    }
    store8(dz + off, b);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kT; ++k) t += sm[k];
    atomicAdd(loss_acc, static_cast<double>(t));
  }
}

__global__ void g_finish_double_kernel(const double* __restrict__ acc, float* __restrict__ out, float scale, int accumulate) {
  pdl_sync();
  const float v = static_cast<float>(acc[0] * static_cast<double>(scale));
  out[0] = accumulate ? out[0] + v : v;
}

// ---------------------------------------------------------------------------------------------
// discriminator head (discriminator.py:61-74): flatten (NCHW order) -> dense1 -> LeakyReLU -> dense2 -> sigmoid
// ---------------------------------------------------------------------------------------------
__global__ void g_flatten_kernel(const bf16_t* __restrict__ h, float* __restrict__ flat, TG g) {
  pdl_sync();
  const long long chw = static_cast<long long>(g.C) * g.H * g.W;
  const long long total = chw * g.B;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    const int b = static_cast<int>(i / chw);
    const long long r = i - b * chw;
    const int c = static_cast<int>(r / (g.H * g.W));
    const int p = static_cast<int>(r - static_cast<long long>(c) * g.H * g.W);
    const int y = p / g.W, x = p - y * g.W;
    flat[i] = __bfloat162float(h[(static_cast<long long>(b * g.P + y) * g.W + x) * g.C + c]);
  }
}
__global__ void g_unflatten_kernel(const float* __restrict__ dflat, bf16_t* __restrict__ dh, TG g) {
  pdl_sync();
  const long long chw = static_cast<long long>(g.C) * g.H * g.W;
  const long long total = chw * g.B;
  for (long long i = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * kT) {
    // i enumerates NHWC so that the stores coalesce
    const int c = static_cast<int>(i % g.C);
    const long long pq = i / g.C;
    const int hw = g.H * g.W;
    const int b = static_cast<int>(pq / hw);
    const int p = static_cast<int>(pq - static_cast<long long>(b) * hw);
    const int y = p / g.W, x = p - y * g.W;
    dh[(static_cast<long long>(b * g.P + y) * g.W + x) * g.C + c] =
        __float2bfloat16_rn(dflat[b * chw + static_cast<long long>(c) * hw + p]);
  }
}

// 16-byte global -> shared copies that need no registers while in flight (the dense head streams a 302 MB matrix with
// 8 warps per SM: the bytes in flight have to live somewhere else than the register file)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// z1[b][j] = bias[j] + sum_k W[j][k] x[b][k].  grid (J / kD1Rows, kD1KSplit): a block owns kD1Rows rows of W (fp32,
// streamed once) and a ninth of the columns; partial sums are added atomically into z1 (zeroed first), B <= 8.
// Every thread keeps a PRIVATE ring of kD1Stages x (8 W + 8 x) 16-byte slots in shared memory, filled by cp.async: a
// thread only ever reads the slots it copied itself, so the ring needs no barrier, and 2 blocks x 128 threads x 2 stages
// x 256 B = 128 KB per SM are in flight (the register-staged form had 159.8 us for 304 MB = 1.9 TB/s at 22 % of
// the warp slots, every warp waiting on its own loads: profiles/r02_gan_train_dense1_fwd_before_raw.csv).
constexpr int kD1Rows = 8, kD1KSplit = 9, kD1T = 128, kD1Stages = 3;
constexpr int kD1FwdSmem = kD1Stages * 16 * kD1T * 16;
__global__ void __launch_bounds__(kD1T) g_dense1_fwd_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                            const float* __restrict__ x, float* __restrict__ z1, int B, int K,
                                                            int J) {
  extern __shared__ float4 d1ring[];               // [stage][16][kD1T]
  __shared__ float sm[kD1T / 32][kD1Rows * 8];
  pdl_sync();
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * kD1Rows;
  float acc[kD1Rows][8];
#pragma unroll
  for (int r = 0; r < kD1Rows; ++r)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[r][b] = 0.f;
  const int k4n = K >> 2;
  const int kb = static_cast<int>((static_cast<long long>(k4n) * blockIdx.y) / gridDim.y);
  const int ke = static_cast<int>((static_cast<long long>(k4n) * (blockIdx.y + 1)) / gridDim.y);
  const int niter = (ke - kb + kD1T - 1) / kD1T;
  const float4* W4 = reinterpret_cast<const float4*>(W) + static_cast<long long>(j0) * k4n;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  auto issue = [&](int it) {
    const int k4 = kb + it * kD1T + tid;
    if (it < niter && k4 < ke) {
      float4* st = d1ring + (it % kD1Stages) * 16 * kD1T + tid;
#pragma unroll
      for (int r = 0; r < kD1Rows; ++r) cp_async16(st + r * kD1T, W4 + static_cast<long long>(r) * k4n + k4);
#pragma unroll
      for (int b = 0; b < 8; ++b)
        if (b < B) cp_async16(st + (8 + b) * kD1T, x4 + static_cast<long long>(b) * k4n + k4);
    }
    cp_async_commit();
  };
  for (int it = 0; it < kD1Stages - 1; ++it) issue(it);
  for (int it = 0; it < niter; ++it) {
    issue(it + kD1Stages - 1);
    cp_async_wait<kD1Stages - 1>();
    const int k4 = kb + it * kD1T + tid;
    if (k4 < ke) {
      const float4* st = d1ring + (it % kD1Stages) * 16 * kD1T + tid;
      float4 w[kD1Rows];
#pragma unroll
      for (int r = 0; r < kD1Rows; ++r) w[r] = st[r * kD1T];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (b < B) {
          const float4 xv = st[(8 + b) * kD1T];
#pragma unroll
          for (int r = 0; r < kD1Rows; ++r) acc[r][b] += w[r].x * xv.x + w[r].y * xv.y + w[r].z * xv.z + w[r].w * xv.w;
        }
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < kD1Rows; ++r)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      float v = acc[r][b];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sm[warp][r * 8 + b] = v;
    }
  __syncthreads();
  if (threadIdx.x < kD1Rows * 8) {
    const int r = threadIdx.x >> 3, b = threadIdx.x & 7;
    float v = 0.f;
    for (int w = 0; w < kD1T / 32; ++w) v += sm[w][threadIdx.x];
    if (blockIdx.y == 0) v += bias[j0 + r];
    if (b < B) atomicAdd(&z1[b * J + j0 + r], v);
  }
}

__global__ void g_dense2_fwd_kernel(const float* __restrict__ z1, const float* __restrict__ w2, const float* __restrict__ b2,
                                    float* __restrict__ prob, int J) {
  __shared__ float sm[kT];
  pdl_sync();
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int j = threadIdx.x; j < J; j += kT) {
    const float z = z1[b * J + j];
    acc += w2[j] * (z > 0.f ? z : z * kLrelu);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kT; ++k) t += sm[k];
    t += b2[0];
    prob[b] = 1.f / (1.f + expf(-t));
  }
}

__global__ void g_dense2_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dprob, float target,
                                    const float* __restrict__ z1, const float* __restrict__ w2, float* __restrict__ dz1,
                                    float* __restrict__ dw2, float* __restrict__ db2, int B, int J) {
  __shared__ float dl[8];
  pdl_sync();
  if (threadIdx.x < 8) {
    float v = 0.f;
    if (threadIdx.x < B) {
      const float p = prob[threadIdx.x];
      v = (dprob != nullptr) ? dprob[threadIdx.x] * p * (1.f - p) : (p - target) / static_cast<float>(B);
    }
    dl[threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = blockIdx.x * kT + threadIdx.x; j < J; j += gridDim.x * kT) {
    float g = 0.f;
    for (int b = 0; b < B; ++b) {
      const float z = z1[b * J + j];
      g += dl[b] * (z > 0.f ? z : z * kLrelu);
      dz1[b * J + j] = dl[b] * w2[j] * (z > 0.f ? 1.f : kLrelu);
    }
    dw2[j] += g;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += dl[b];
    db2[0] += t;
  }
}

// dW[j][k] += sum_b dz1[b][j] x[b][k];  dx[b][k] = sum_j dz1[b][j] W[j][k];  db1[j] += sum_b dz1[b][j], for NG groups
// of B <= 8 samples in ONE sweep over W / dW (the real and the fake pass of do_epoch: the 302 MB matrix and its gradient
// are streamed once instead of twice).  grid (ceil(K / 1024), JS): a block owns 1024 columns and J / JS rows; dx
// partials are added atomically (dx zeroed first).
constexpr int kD1Split = 8;
constexpr int kD1BR = 4, kD1BS = 4;             // rows per ring stage, stages: 4 x 4 x (W + dW) x 16 B = 512 B per thread
constexpr int kD1BwdRing = kD1BS * 2 * kD1BR * kT * 16;
struct D1Grp { const float* x; const float* dz1; float* dx; };
// W and dW rows reach the thread through a private cp.async ring (see g_dense1_fwd_kernel): 3 stages x 4 rows x 32 B x
// 256 threads = 96 KB in flight per SM instead of the one row (32 B per thread) of the plain loop, which ran the
// 0.9 GB sweep at 2.25 TB/s (402 us).
template <int NG>
__global__ void __launch_bounds__(kT) g_dense1_bwd_kernel(const float* __restrict__ W, D1Grp g0, D1Grp g1,
                                                          float* __restrict__ dW, float* __restrict__ db1, int B, int K,
                                                          int J, int acc) {      // acc = 0: dW is written, not added to
  extern __shared__ float4 d1bring[];            // [stage][2 * kD1BR][kT], then dzs [jn][8 * NG]
  float* dzs = reinterpret_cast<float*>(d1bring + kD1BS * 2 * kD1BR * kT);
  pdl_sync();
  const D1Grp grp[2] = {g0, g1};
  const int jn = (J + kD1Split - 1) / kD1Split;
  const int jb = blockIdx.y * jn;
  const int je = (jb + jn < J) ? jb + jn : J;
  for (int i = threadIdx.x; i < jn * 8 * NG; i += kT) {
    const int j = jb + i / (8 * NG), r = i % (8 * NG), g = r >> 3, b = r & 7;
    dzs[i] = (j < J && b < B) ? grp[g].dz1[b * J + j] : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int j = jb + threadIdx.x; j < je; j += kT) {
      float t = 0.f;
      for (int b = 0; b < 8 * NG; ++b) t += dzs[(j - jb) * 8 * NG + b];
      db1[j] += t;
    }
  }
  const int k4 = blockIdx.x * kT + threadIdx.x;
  if (k4 * 4 >= K) return;
  const int k4n = K >> 2;
  const float4* W4 = reinterpret_cast<const float4*>(W) + k4;
  float4* G4 = reinterpret_cast<float4*>(dW) + k4;
  const int nrg = (je - jb + kD1BR - 1) / kD1BR;
  auto issue = [&](int rg) {
    if (rg < nrg) {
      float4* st = d1bring + (rg % kD1BS) * 2 * kD1BR * kT + threadIdx.x;
#pragma unroll
      for (int r = 0; r < kD1BR; ++r) {
        const int j = jb + rg * kD1BR + r;
        if (j < je) {
          cp_async16(st + (2 * r) * kT, W4 + static_cast<long long>(j) * k4n);
          if (acc) cp_async16(st + (2 * r + 1) * kT, G4 + static_cast<long long>(j) * k4n);
        }
      }
    }
    cp_async_commit();
  };
  for (int rg = 0; rg < kD1BS - 1; ++rg) issue(rg);
  float4 xv[NG][8], da[NG][8];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      xv[g][b] = (b < B) ? __ldg(reinterpret_cast<const float4*>(grp[g].x + static_cast<long long>(b) * K) + k4)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      da[g][b] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  for (int rg = 0; rg < nrg; ++rg) {
    issue(rg + kD1BS - 1);
    cp_async_wait<kD1BS - 1>();
    const float4* st = d1bring + (rg % kD1BS) * 2 * kD1BR * kT + threadIdx.x;
#pragma unroll
    for (int r = 0; r < kD1BR; ++r) {
      const int j = jb + rg * kD1BR + r;
      if (j < je) {
        const float4 w = st[(2 * r) * kT];
        float4 o = acc ? st[(2 * r + 1) * kT] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* dzj = reinterpret_cast<const float4*>(&dzs[(j - jb) * 8 * NG]);
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 d4 = dzj[g * 2 + h];
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int b = h * 4 + q;
              const float d = dd[q];
              da[g][b].x += d * w.x; da[g][b].y += d * w.y; da[g][b].z += d * w.z; da[g][b].w += d * w.w;
              o.x += d * xv[g][b].x; o.y += d * xv[g][b].y; o.z += d * xv[g][b].z; o.w += d * xv[g][b].w;
            }
          }
        __stcs(G4 + static_cast<long long>(j) * k4n, o);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      if (b < B) {
        float* o = grp[g].dx + static_cast<long long>(b) * K + k4 * 4;
        atomicAdd(o, da[g][b].x); atomicAdd(o + 1, da[g][b].y); atomicAdd(o + 2, da[g][b].z); atomicAdd(o + 3, da[g][b].w);
      }
    }
}

// nn.BCELoss (mean reduction, log clamped at -100) against a constant target (utils/GAN.py:96-107)
__global__ void g_bce_kernel(const float* __restrict__ prob, float target, int B, float* __restrict__ loss, int accumulate) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float t = 0.f;
  for (int b = 0; b < B; ++b) {
    const float p = prob[b];
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);
    t -= target * lp + (1.f - target) * lq;
  }
  t /= static_cast<float>(B);
  loss[0] = accumulate ? loss[0] + t : t;
}

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
// One launch for all convolutions of a network (a launch at these sizes costs 4 us whatever it does; the generator has
// 36 layers): block b works on the item whose [blk0, blk0 + nblk) holds it, grid-striding over the item's blocks.
__global__ void g_pack_group_kernel(const float* __restrict__ params, const GPackItem* __restrict__ items, int n) {
  pdl_sync();
  int it = 0;
  while (it + 1 < n && static_cast<int>(blockIdx.x) >= items[it + 1].blk0) ++it;
  const GPackItem q = items[it];
  const float* w = params + q.w_off;
  const int kk = q.ks * q.ks;
  const long long total = static_cast<long long>(kk) * q.cout_pad * q.cin_pad;
  for (long long i = static_cast<long long>(static_cast<int>(blockIdx.x) - q.blk0) * kT + threadIdx.x; i < total;
       i += static_cast<long long>(q.nblk) * kT) {
    const int ci = static_cast<int>(i % q.cin_pad);
    const long long r = i / q.cin_pad;
    const int co = static_cast<int>(r % q.cout_pad);
    const int tap = static_cast<int>(r / q.cout_pad);
    float v = 0.f;
    if (co < q.cout && ci < q.cin) v = w[(static_cast<long long>(co) * q.cin + ci) * kk + tap];
    const bf16_t h = __float2bfloat16_rn(v);
    if (q.w_f != nullptr) {
      if (q.f16_fwd) reinterpret_cast<__half*>(q.w_f)[i] = __float2half_rn(v);
      else q.w_f[i] = h;
    }
    if (q.w_d != nullptr) q.w_d[(static_cast<long long>(tap) * q.cin_pad + ci) * q.cout_pad + co] = h;
    if (tap == 0 && ci == 0) q.bias_pad[co] = (q.b_off >= 0 && co < q.cout) ? params[q.b_off + co] : 0.f;
  }
}
__global__ void g_unpack_group_kernel(float* __restrict__ grads, const GUnpackItem* __restrict__ items, int n) {
  pdl_sync();
  int it = 0;
  while (it + 1 < n && static_cast<int>(blockIdx.x) >= items[it + 1].blk0) ++it;
  const GUnpackItem q = items[it];
  float* g = grads + q.g_off;
  const int kk = q.ks * q.ks;
  const long long total = static_cast<long long>(q.cout) * q.cin * kk;
  for (long long i = static_cast<long long>(static_cast<int>(blockIdx.x) - q.blk0) * kT + threadIdx.x; i < total;
       i += static_cast<long long>(q.nblk) * kT) {
    const int tap = static_cast<int>(i % kk);
    const long long r = i / kk;
    const int ci = static_cast<int>(r % q.cin);
    const int co = static_cast<int>(r / q.cin);
    g[i] += q.dw[(static_cast<long long>(tap) * q.cout + co) * q.cin + ci];
  }
}

// ---------------------------------------------------------------------------------------------
// 9 x 9, 64 -> 3 output convolution (generator.py:60), backward: the kx taps FOLDED into the channel dimension.
//   dz9[(y, x)][kx * 3 + co] = dz[(y, x - kx + 4)][co]      (27 of 64 channels; zero outside the image)
// turns both gradients into 9-tap (ky) tensor-core problems with 64-channel operands:
//   dW[co][ci][ky][kx] = sum_q dz9[q][kx*3+co] * x[q + (ky-4, 0)][ci]            -> gwgrad_kernel
//   dx[q][ci]          = sum_ky sum_j dz9[q - (ky-4, 0)][j] * w9[ky][ci][j]      -> gconv_kernel
// ---------------------------------------------------------------------------------------------
// dst[(y, x)][kx * 3 + c] = src[(y, x + sign * (kx - pad))][c], c < 3, kx < KS; the remaining channels of the first
// 16 (KS = 3) / 32 (KS = 9) are written as zero, the rest of dst's pitch stays zero from bind time
template <int KS>
__global__ void g_expand_kx_kernel(const bf16_t* __restrict__ src, bf16_t* __restrict__ dst, TG g, int sC, int sign, int pad) {
  constexpr int NV = (KS * 3 + 15) / 16 * 16;
  pdl_sync();
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  const int hw = g.H * g.W;
  for (long long q = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; q < np; q += static_cast<long long>(gridDim.x) * kT) {
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - static_cast<long long>(b) * hw);
    const int y = r / g.W, x = r - y * g.W;
    const bf16_t* row = src + static_cast<long long>(b * g.P + y) * g.W * sC;
    __align__(16) bf16_t v[NV];
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      const int xs = x + sign * (kx - pad);
      const bool in = xs >= 0 && xs < g.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[kx * 3 + c] = in ? row[static_cast<long long>(xs) * sC + c] : __float2bfloat16_rn(0.f);
    }
#pragma unroll
    for (int j = KS * 3; j < NV; ++j) v[j] = __float2bfloat16_rn(0.f);
    uint4* d4 = reinterpret_cast<uint4*>(dst + (static_cast<long long>(b * g.P + y) * g.W + x) * g.C);
    const uint4* s4 = reinterpret_cast<const uint4*>(v);
#pragma unroll
    for (int k = 0; k < NV / 8; ++k) d4[k] = s4[k];
  }
}
// discriminator conv0 (3 x 3, 3 -> 64) weight gradient from the folded layout: g[co][ci][ky][kx] += dw3[ky][co][kx * 3 + ci]
__global__ void g_unpack3_kernel(const float* __restrict__ dw3, float* __restrict__ g) {
  pdl_sync();
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= 64 * 27) return;
  const int kx = i % 3, ky = (i / 3) % 3, ci = (i / 9) % 3, co = i / 27;
  g[i] += dw3[(ky * 64 + co) * 64 + kx * 3 + ci];
}
// FORWARD of the same layer, kx folded into N:  S[q][kx*3+co] = sum_ky sum_ci x[q + (ky-4, 0)][ci] * w[co][ci][ky][kx]
// (gconv_kernel, 9 taps, N = 32, fp32 out), then out[b][co][y][x] = tanh(bias[co] + sum_kx S[(y, x + kx - 4)][kx*3+co])
// (S is zero outside the image columns because x is) -- the tanh / NCHW output pass does the 9-term fold.
__global__ void g_fold9_tanh_kernel(const float* __restrict__ S, const float* __restrict__ bias, float* __restrict__ z_out,
                                    float* __restrict__ out, TG g) {
  pdl_sync();
  const long long np = static_cast<long long>(g.B) * g.H * g.W;
  const long long hw = static_cast<long long>(g.H) * g.W;
  for (long long q = static_cast<long long>(blockIdx.x) * kT + threadIdx.x; q < np; q += static_cast<long long>(gridDim.x) * kT) {
    const int b = static_cast<int>(q / hw);
    const int r = static_cast<int>(q - b * hw);
    const int y = r / g.W, x = r - y * g.W;
    const float* row = S + static_cast<long long>(b * g.P + y) * g.W * g.C;
    float a0 = bias[0], a1 = bias[1], a2 = bias[2];
#pragma unroll
    for (int kx = 0; kx < 9; ++kx) {
      const int xs = x + kx - 4;
      if (xs >= 0 && xs < g.W) {
        const float* p3 = row + static_cast<long long>(xs) * g.C + kx * 3;
        a0 += p3[0]; a1 += p3[1]; a2 += p3[2];
      }
    }
    out[(static_cast<long long>(b) * 3 + 0) * hw + r] = tanhf(a0);
    out[(static_cast<long long>(b) * 3 + 1) * hw + r] = tanhf(a1);
    out[(static_cast<long long>(b) * 3 + 2) * hw + r] = tanhf(a2);
    if (z_out != nullptr) {                 // pre-activation for tests: fp32, 16-channel pitch
      float* zp = z_out + (static_cast<long long>(b * g.P + y) * g.W + x) * 16;
      zp[0] = a0; zp[1] = a1; zp[2] = a2;
    }
  }
}
// w9f[ky][kx * 3 + co][ci] = w[co][ci][ky][kx]  (bf16 [9][32][64], rows >= 27 zero)
__global__ void g_pack9f_kernel(const float* __restrict__ w, bf16_t* __restrict__ w9f) {
  pdl_sync();
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= 9 * 32 * 64) return;
  const int ci = i & 63, j = (i >> 6) & 31, ky = i >> 11;
  float v = 0.f;
  if (j < 27) {
    const int kx = j / 3, co = j - 3 * kx;
    v = w[((co * 64 + ci) * 9 + ky) * 9 + kx];
  }
  w9f[i] = __float2bfloat16_rn(v);
}
// w9[ky][ci][kx * 3 + co] = w[co][ci][ky][kx]  (bf16 [9][64][64], columns >= 27 zero)
__global__ void g_pack9_kernel(const float* __restrict__ w, bf16_t* __restrict__ w9) {
  pdl_sync();
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= 9 * 64 * 64) return;
  const int j = i & 63, ci = (i >> 6) & 63, ky = i >> 12;
  float v = 0.f;
  if (j < 27) {
    const int kx = j / 3, co = j - 3 * kx;
    v = w[((co * 64 + ci) * 9 + ky) * 9 + kx];
  }
  w9[i] = __float2bfloat16_rn(v);
}
// g[co][ci][ky][kx] += dw9[ky][kx * 3 + co][ci]
__global__ void g_unpack9_kernel(const float* __restrict__ dw9, float* __restrict__ g) {
  pdl_sync();
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= 3 * 64 * 81) return;
  const int kx = i % 9, ky = (i / 9) % 9, ci = (i / 81) % 64, co = i / (81 * 64);
  g[i] += dw9[(ky * 64 + kx * 3 + co) * 64 + ci];
}

// ---------------------------------------------------------------------------------------------
// CUDA-core weight gradients of the 3-channel layers (0.3 % of the step's FLOPs; K = 3 is no tensor-core shape)
// ---------------------------------------------------------------------------------------------
// dW[co][ci][ky][kx] += sum_p dy[p][co] x[p + (ky, kx) - pad][ci];  dy: 64 channels, x: 16-channel pitch (3 used)
template <int KS>
__global__ void __launch_bounds__(kT) g_wgrad_in3_kernel(const bf16_t* __restrict__ dy, const bf16_t* __restrict__ x,
                                                         float* __restrict__ g, TG gy, int xC) {
  constexpr int PW = 8 + KS - 1, NC = 3 * KS * KS, NM = (NC + 3) / 4, PAD = (KS - 1) / 2;
  __shared__ float xs[PW * PW * 3];
  __shared__ __align__(16) bf16_t dys[64 * 64];
  __shared__ short offs[NC];
  pdl_sync();
  const int co = threadIdx.x & 63, qd = threadIdx.x >> 6;
  const int rows = gy.B * gy.P;
  const int tiles_x = (gy.W + 7) / 8, tiles_y = (rows + 7) / 8;
  for (int m = threadIdx.x; m < NC; m += kT) {
    const int ci = m % 3, tap = m / 3;
    offs[m] = static_cast<short>(((tap / KS) * PW + (tap % KS)) * 3 + ci);
  }
  float acc[NM];
#pragma unroll
  for (int i = 0; i < NM; ++i) acc[i] = 0.f;
  for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
    const int x0 = (tile % tiles_x) * 8, y0 = (tile / tiles_x) * 8;
    __syncthreads();
    {
      const int p = threadIdx.x >> 2, part = threadIdx.x & 3;
      const int yy = y0 + (p >> 3), xx = x0 + (p & 7);
      uint4 a = make_uint4(0, 0, 0, 0), b = a;
      if (yy < rows && xx < gy.W) {
        const uint4* src = reinterpret_cast<const uint4*>(dy + (static_cast<long long>(yy) * gy.W + xx) * 64 + part * 16);
        a = src[0]; b = src[1];
      }
      uint4* dst = reinterpret_cast<uint4*>(&dys[p * 64 + part * 16]);
      dst[0] = a; dst[1] = b;
    }
    for (int i = threadIdx.x; i < PW * PW; i += kT) {
      const int yy = y0 + i / PW - PAD, xx = x0 + i % PW - PAD;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
      if (yy >= 0 && yy < rows && xx >= 0 && xx < gy.W) {
        const bf16_t* src = x + (static_cast<long long>(yy) * gy.W + xx) * xC;
        v0 = __bfloat162float(src[0]); v1 = __bfloat162float(src[1]); v2 = __bfloat162float(src[2]);
      }
      xs[i * 3] = v0; xs[i * 3 + 1] = v1; xs[i * 3 + 2] = v2;
    }
    __syncthreads();
    for (int p = 0; p < 64; ++p) {
      const float d = __bfloat162float(dys[p * 64 + co]);
      const int base = ((p >> 3) * PW + (p & 7)) * 3;
#pragma unroll
      for (int mm = 0; mm < NM; ++mm) {
        const int m = qd + 4 * mm;
        if (m < NC) acc[mm] += d * xs[base + offs[m]];
      }
    }
  }
#pragma unroll
  for (int mm = 0; mm < NM; ++mm) {
    const int m = qd + 4 * mm;
    if (m < NC) atomicAdd(&g[(co * 3 + (m % 3)) * KS * KS + m / 3], acc[mm]);
  }
}

}  // namespace

// =============================================================================================
// launch wrappers
// =============================================================================================
int gl_pack_image(const float* nchw, const GT& out, cudaStream_t s) {
  launch_k(g_pack_image_kernel, dim3(grid_for(static_cast<long long>(out.B) * out.H * out.W)), dim3(kT), 0, s, nchw,
           static_cast<bf16_t*>(out.ptr), tg_of(out));
  GL_CHECK();
}
int gl_tanh_bwd(const float* dout, const float* out, const GT& dz, float* dbias3, cudaStream_t s) {
  launch_k(g_tanh_bwd_kernel, dim3(grid_for(static_cast<long long>(dz.B) * dz.H * dz.W, 2)), dim3(kT), 0, s, dout, out,
           static_cast<bf16_t*>(dz.ptr), dbias3, tg_of(dz));
  GL_CHECK();
}

int gl_bn_apply(const GT& raw, const GT& out, const bf16_t* res, const double* stats, const float* gamma, const float* beta,
                int act, const float* slope, const float* conv_bias, float* rm, float* rv, int run_times, cudaStream_t s) {
  const long long items = static_cast<long long>(raw.B) * raw.H * raw.W * (raw.C / 8);
  const double inv = 1.0 / (static_cast<double>(raw.B) * raw.H * raw.W);
  const dim3 grid(grid_for(items));
  const bf16_t* r = static_cast<const bf16_t*>(raw.ptr);
  bf16_t* o = static_cast<bf16_t*>(out.ptr);
  if (raw.C > 512 || (raw.C & 7)) return -51;
  if (act == GACT_NONE) launch_k(g_bn_apply_kernel<GACT_NONE>, grid, dim3(kT), 0, s, r, o, res, stats, gamma, beta, slope, tg_of(raw), inv, conv_bias, rm, rv, run_times);
  else if (act == GACT_LRELU) launch_k(g_bn_apply_kernel<GACT_LRELU>, grid, dim3(kT), 0, s, r, o, res, stats, gamma, beta, slope, tg_of(raw), inv, conv_bias, rm, rv, run_times);
  else launch_k(g_bn_apply_kernel<GACT_PRELU>, grid, dim3(kT), 0, s, r, o, res, stats, gamma, beta, slope, tg_of(raw), inv, conv_bias, rm, rv, run_times);
  GL_CHECK();
}
int gl_bn_bwd(const GT& dy, const GT& raw, const GT& draw, const double* stats, const float* gamma, const float* beta,
              int act, const float* slope, double* sums2, int* parity, float* dgamma, float* dbeta, float* dslope,
              cudaStream_t s) {
  const int C = raw.C;
  if (C > 512 || (C & 7) || (kT % (C / 8))) return -51;
  // sums2: two halves of kBnSumsHalf doubles, both zero at bind time; *parity selects the half this call accumulates
  // into, the apply kernel clears the other one
  double* sums = sums2 + (*parity ? kBnSumsHalf : 0);
  double* other = sums2 + (*parity ? 0 : kBnSumsHalf);
  *parity ^= 1;
  const long long np = static_cast<long long>(raw.B) * raw.H * raw.W;
  const double inv = 1.0 / static_cast<double>(np);
  const dim3 g1(grid_for(np * (C / 8), 2)), g2(grid_for(np * (C / 8)));
  const bf16_t* d = static_cast<const bf16_t*>(dy.ptr);
  const bf16_t* r = static_cast<const bf16_t*>(raw.ptr);
  bf16_t* o = static_cast<bf16_t*>(draw.ptr);
  const TG g = tg_of(raw);
  if (act == GACT_NONE) {
    launch_k(g_bn_bwd_stats_kernel<GACT_NONE>, g1, dim3(kT), 0, s, d, r, stats, gamma, beta, slope, g, inv, sums);
    launch_k(g_bn_bwd_apply_kernel<GACT_NONE>, g2, dim3(kT), 0, s, d, r, o, stats, gamma, beta, slope, g, inv, static_cast<const double*>(sums), dgamma, dbeta, dslope, other);
  } else if (act == GACT_LRELU) {
    launch_k(g_bn_bwd_stats_kernel<GACT_LRELU>, g1, dim3(kT), 0, s, d, r, stats, gamma, beta, slope, g, inv, sums);
    launch_k(g_bn_bwd_apply_kernel<GACT_LRELU>, g2, dim3(kT), 0, s, d, r, o, stats, gamma, beta, slope, g, inv, static_cast<const double*>(sums), dgamma, dbeta, dslope, other);
  } else {
    launch_k(g_bn_bwd_stats_kernel<GACT_PRELU>, g1, dim3(kT), 0, s, d, r, stats, gamma, beta, slope, g, inv, sums);
    launch_k(g_bn_bwd_apply_kernel<GACT_PRELU>, g2, dim3(kT), 0, s, d, r, o, stats, gamma, beta, slope, g, inv, static_cast<const double*>(sums), dgamma, dbeta, dslope, other);
  }
  GL_CHECK();
}

int gl_shuffle_fwd(const GT& sraw, const GT& u, const float* slope, cudaStream_t s) {
  if (sraw.C != 256 || u.C != 64 || u.W != 2 * sraw.W || u.H != 2 * sraw.H) return -52;
  launch_k(g_shuffle_fwd_kernel, dim3(grid_for(static_cast<long long>(sraw.B) * sraw.H * sraw.W * 8)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(sraw.ptr), static_cast<bf16_t*>(u.ptr), slope, tg_of(sraw), tg_of(u));
  GL_CHECK();
}
int gl_shuffle_bwd(const GT& du, const GT& sraw, const GT& ds, const float* slope, float* dbias256, float* dslope,
                   cudaStream_t s) {
  if (sraw.C != 256 || du.C != 64) return -52;
  launch_k(g_shuffle_bwd_kernel, dim3(grid_for(static_cast<long long>(sraw.B) * sraw.H * sraw.W * 8, 2)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(du.ptr), static_cast<const bf16_t*>(sraw.ptr), static_cast<bf16_t*>(ds.ptr), slope,
           dbias256, dslope, tg_of(sraw), tg_of(du));
  GL_CHECK();
}
int gl_prelu_fwd(const GT& z, const GT& out, const float* slope, cudaStream_t s) {
  launch_k(g_prelu_fwd_kernel, dim3(grid_for(static_cast<long long>(z.B) * z.H * z.W * (z.C / 8))), dim3(kT), 0, s,
           static_cast<const bf16_t*>(z.ptr), static_cast<bf16_t*>(out.ptr), slope, tg_of(z));
  GL_CHECK();
}
int gl_prelu_bwd(const GT& d1, const bf16_t* d2, const GT& z, const GT& dz, const float* slope, float* dbias, float* dslope,
                 cudaStream_t s) {
  if (kT % (z.C / 8)) return -51;
  launch_k(g_prelu_bwd_kernel, dim3(grid_for(static_cast<long long>(z.B) * z.H * z.W * (z.C / 8), 2)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(d1.ptr), d2, static_cast<const bf16_t*>(z.ptr), static_cast<bf16_t*>(dz.ptr), slope,
           dbias, dslope, tg_of(z));
  GL_CHECK();
}
int gl_chan_sum(const GT& t, float* out_c, cudaStream_t s) {
  if (kT % (t.C / 8)) return -51;
  launch_k(g_chan_sum_kernel, dim3(grid_for(static_cast<long long>(t.B) * t.H * t.W * (t.C / 8), 2)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(t.ptr), out_c, tg_of(t));
  GL_CHECK();
}

int gl_vgg_pre_fwd(const float* img, int Hi, int Wi, int Hr, int Wr, int top, int left, const GT& out, cudaStream_t s) {
  launch_k(g_vgg_pre_fwd_kernel, dim3(grid_for(static_cast<long long>(out.B) * out.H * out.W)), dim3(kT), 0, s, img,
           static_cast<bf16_t*>(out.ptr), tg_of(out), Hi, Wi, Hr, Wr, top, left, out.f16);
  GL_CHECK();
}
int gl_vgg_pre_bwd(const GT& dpre, int Hi, int Wi, int Hr, int Wr, int top, int left, float* dimg, int accumulate,
                   cudaStream_t s) {
  launch_k(g_vgg_pre_bwd_kernel, dim3(grid_for(static_cast<long long>(dpre.B) * Hi * Wi)), dim3(kT), 0, s,
           static_cast<const float*>(dpre.ptr), dimg, tg_of(dpre), Hi, Wi, Hr, Wr, top, left, accumulate);
  GL_CHECK();
}
int gl_maxpool_fwd(const GT& in, const GT& out, cudaStream_t s) {
  if (in.C != out.C || in.H != 2 * out.H || in.W != 2 * out.W) return -53;
  launch_k(g_maxpool_fwd_kernel, dim3(grid_for(static_cast<long long>(out.B) * out.H * out.W * (out.C / 8))), dim3(kT), 0, s,
           static_cast<const bf16_t*>(in.ptr), static_cast<bf16_t*>(out.ptr), tg_of(in), tg_of(out));
  GL_CHECK();
}
int gl_maxpool_bwd(const GT& dout, const GT& y_in, const GT& dz_in, cudaStream_t s) {
  if (y_in.C != dout.C || y_in.H != 2 * dout.H || y_in.W != 2 * dout.W) return -53;
  launch_k(g_maxpool_bwd_kernel, dim3(grid_for(static_cast<long long>(dout.B) * dout.H * dout.W * (dout.C / 8))), dim3(kT),
           0, s, static_cast<const bf16_t*>(dout.ptr), static_cast<const bf16_t*>(y_in.ptr), static_cast<bf16_t*>(dz_in.ptr),
           tg_of(y_in), tg_of(dout));
  GL_CHECK();
}
int gl_feat_mse(const GT& f1, const GT& f2, const GT& dz, double* loss_acc, cudaStream_t s) {
  const long long n = static_cast<long long>(f1.B) * f1.H * f1.W * f1.C;
  launch_k(g_feat_mse_kernel, dim3(grid_for(n / 8, 2)), dim3(kT), 0, s, static_cast<const bf16_t*>(f1.ptr),
           static_cast<const bf16_t*>(f2.ptr), static_cast<bf16_t*>(dz.ptr), tg_of(f1), 2.f / static_cast<float>(n), loss_acc, f1.f16);
  GL_CHECK();
}
int gl_finish_double(const double* acc, float* out, float scale, int accumulate, cudaStream_t s) {
  launch_k(g_finish_double_kernel, dim3(1), dim3(1), 0, s, acc, out, scale, accumulate);
  GL_CHECK();
}

int gl_flatten(const GT& h, float* flat, cudaStream_t s) {
  launch_k(g_flatten_kernel, dim3(grid_for(static_cast<long long>(h.B) * h.C * h.H * h.W)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(h.ptr), flat, tg_of(h));
  GL_CHECK();
}
int gl_unflatten(const float* dflat, const GT& dh, cudaStream_t s) {
  launch_k(g_unflatten_kernel, dim3(grid_for(static_cast<long long>(dh.B) * dh.C * dh.H * dh.W)), dim3(kT), 0, s, dflat,
           static_cast<bf16_t*>(dh.ptr), tg_of(dh));
  GL_CHECK();
}
int gl_dense1_fwd(const float* W, const float* bias, const float* x, float* z1, int B, int K, int J, cudaStream_t s) {
  if (B > 8 || (K & 3) || (J % kD1Rows)) return -54;
  static bool attr_dev[kMaxDevices] = {};            // opt-in shared memory is a per-device function attribute
  bool& attr = attr_dev[device_slot()];
  if (!attr) {
    if (cudaFuncSetAttribute(g_dense1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kD1FwdSmem) != cudaSuccess) return -59;
    attr = true;
  }
  cudaMemsetAsync(z1, 0, static_cast<size_t>(B) * J * sizeof(float), s);
  launch_k(g_dense1_fwd_kernel, dim3(J / kD1Rows, kD1KSplit), dim3(kD1T), kD1FwdSmem, s, W, bias, x, z1, B, K, J);
  GL_CHECK();
}
int gl_dense2_fwd(const float* z1, const float* w2, const float* b2, float* prob, int B, int J, cudaStream_t s) {
  launch_k(g_dense2_fwd_kernel, dim3(B), dim3(kT), 0, s, z1, w2, b2, prob, J);
  GL_CHECK();
}
int gl_dense2_bwd(const float* prob, const float* dprob, float target, const float* z1, const float* w2, float* dz1,
                  float* dw2, float* db2, int B, int J, cudaStream_t s) {
  if (B > 8) return -54;
  launch_k(g_dense2_bwd_kernel, dim3((J + kT - 1) / kT), dim3(kT), 0, s, prob, dprob, target, z1, w2, dz1, dw2, db2, B, J);
  GL_CHECK();
}
static int dense1_bwd_attr() {
  static bool attr_dev[kMaxDevices] = {};
  bool& attr = attr_dev[device_slot()];
  if (attr) return 0;
  const int cap = kD1BwdRing + 1024 * 16 * static_cast<int>(sizeof(float));
  if (cudaFuncSetAttribute(g_dense1_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess) return -59;
  if (cudaFuncSetAttribute(g_dense1_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap) != cudaSuccess) return -59;
  attr = true;
  return 0;
}
int gl_dense1_bwd(const float* W, const float* x, const float* dz1, float* dW, float* db1, float* dx, int B, int K, int J,
                  cudaStream_t s) {
  if (B > 8 || (K & 3)) return -54;
  const int jn = (J + kD1Split - 1) / kD1Split;
  if (jn > 1024) return -54;
  if (dense1_bwd_attr()) return -59;
  cudaMemsetAsync(dx, 0, static_cast<size_t>(B) * K * sizeof(float), s);
  const D1Grp g0{x, dz1, dx};
  launch_k(g_dense1_bwd_kernel<1>, dim3((K / 4 + kT - 1) / kT, kD1Split), dim3(kT),
           kD1BwdRing + static_cast<size_t>(jn) * 8 * sizeof(float), s, W, g0, g0, dW, db1, B, K, J, 1);
  GL_CHECK();
}
int gl_dense1_bwd2(const float* W, const float* x0, const float* dz0, float* dx0, const float* x1, const float* dz1,
                   float* dx1, float* dW, float* db1, int B, int K, int J, int accumulate, cudaStream_t s) {
  if (B > 8 || (K & 3)) return -54;
  const int jn = (J + kD1Split - 1) / kD1Split;
  if (jn > 1024) return -54;
  if (dense1_bwd_attr()) return -59;
  cudaMemsetAsync(dx0, 0, static_cast<size_t>(B) * K * sizeof(float), s);
  cudaMemsetAsync(dx1, 0, static_cast<size_t>(B) * K * sizeof(float), s);
  const D1Grp g0{x0, dz0, dx0}, g1{x1, dz1, dx1};
  launch_k(g_dense1_bwd_kernel<2>, dim3((K / 4 + kT - 1) / kT, kD1Split), dim3(kT),
           kD1BwdRing + static_cast<size_t>(jn) * 16 * sizeof(float), s, W, g0, g1, dW, db1, B, K, J, accumulate);
  GL_CHECK();
}
int gl_bce(const float* prob, float target, int B, float* loss_out, int accumulate, cudaStream_t s) {
  launch_k(g_bce_kernel, dim3(1), dim3(32), 0, s, prob, target, B, loss_out, accumulate);
  GL_CHECK();
}

int gl_group_blocks(long long items) {
  long long b = (items + 4 * kT - 1) / (4 * kT);
  return static_cast<int>(b < 1 ? 1 : (b > 1024 ? 1024 : b));
}
int gl_pack_group(const float* params, const GPackItem* items_dev, int n, int total_blocks, cudaStream_t s) {
  if (n <= 0) return 0;
  launch_k(g_pack_group_kernel, dim3(total_blocks), dim3(kT), 0, s, params, items_dev, n);
  GL_CHECK();
}
int gl_unpack_group(float* grads, const GUnpackItem* items_dev, int n, int total_blocks, cudaStream_t s) {
  if (n <= 0) return 0;
  launch_k(g_unpack_group_kernel, dim3(total_blocks), dim3(kT), 0, s, grads, items_dev, n);
  GL_CHECK();
}
int gl_expand9(const GT& dz16, const GT& dz9, cudaStream_t s) {
  if (dz9.C != 64 || dz9.W != dz16.W || dz9.P != dz16.P || dz9.H != dz16.H) return -56;
  launch_k(g_expand_kx_kernel<9>, dim3(grid_for(static_cast<long long>(dz9.B) * dz9.H * dz9.W)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(dz16.ptr), static_cast<bf16_t*>(dz9.ptr), tg_of(dz9), dz16.C, -1, 4);
  GL_CHECK();
}
int gl_expand3(const GT& img16, const GT& x9, cudaStream_t s) {
  if (x9.C != 64 || x9.W != img16.W || x9.P != img16.P || x9.H != img16.H) return -56;
  launch_k(g_expand_kx_kernel<3>, dim3(grid_for(static_cast<long long>(x9.B) * x9.H * x9.W)), dim3(kT), 0, s,
           static_cast<const bf16_t*>(img16.ptr), static_cast<bf16_t*>(x9.ptr), tg_of(x9), img16.C, 1, 1);
  GL_CHECK();
}
int gl_unpack3(const float* dw3, float* g_oihw, cudaStream_t s) {
  launch_k(g_unpack3_kernel, dim3((64 * 27 + kT - 1) / kT), dim3(kT), 0, s, dw3, g_oihw);
  GL_CHECK();
}
int gl_pack9(const float* w_oihw, bf16_t* w9, cudaStream_t s) {
  launch_k(g_pack9_kernel, dim3((9 * 64 * 64 + kT - 1) / kT), dim3(kT), 0, s, w_oihw, w9);
  GL_CHECK();
}
int gl_pack9f(const float* w_oihw, bf16_t* w9f, cudaStream_t s) {
  launch_k(g_pack9f_kernel, dim3((9 * 32 * 64 + kT - 1) / kT), dim3(kT), 0, s, w_oihw, w9f);
  GL_CHECK();
}
int gl_fold9_tanh(const GT& S, const float* bias3, float* z16_f32, float* out_nchw, cudaStream_t s) {
  if (S.C != 32 || !S.f32) return -56;
  launch_k(g_fold9_tanh_kernel, dim3(grid_for(static_cast<long long>(S.B) * S.H * S.W)), dim3(kT), 0, s,
           static_cast<const float*>(S.ptr), bias3, z16_f32, out_nchw, tg_of(S));
  GL_CHECK();
}
int gl_unpack9(const float* dw9, float* g_oihw, cudaStream_t s) {
  launch_k(g_unpack9_kernel, dim3((3 * 64 * 81 + kT - 1) / kT), dim3(kT), 0, s, dw9, g_oihw);
  GL_CHECK();
}
int gl_wgrad_in3(const GT& dy64, const GT& x16, float* g, int ks, cudaStream_t s) {
  if (dy64.C != 64 || dy64.W != x16.W || dy64.P != x16.P) return -55;
  const int tiles = ((dy64.W + 7) / 8) * ((dy64.rows() + 7) / 8);
  const int grid = tiles < 296 ? tiles : 296;
  if (ks == 3) launch_k(g_wgrad_in3_kernel<3>, dim3(grid), dim3(kT), 0, s, static_cast<const bf16_t*>(dy64.ptr),
                        static_cast<const bf16_t*>(x16.ptr), g, tg_of(dy64), x16.C);
  else if (ks == 9) launch_k(g_wgrad_in3_kernel<9>, dim3(grid), dim3(kT), 0, s, static_cast<const bf16_t*>(dy64.ptr),
                             static_cast<const bf16_t*>(x16.ptr), g, tg_of(dy64), x16.C);
  else return -55;
  GL_CHECK();
}
}  // namespace dsr
