// dsr_gan.cu -- SRResNet generator inference (models/GAN/generator.py of the reference, eval mode as run by
// eval_GAN.py:87-94) on the same tcgen05 halo-tile convolution kernel as the DIP step.
//
//   x [B,3,h,w] fp32 --conv1 9x9 + PReLU (CUDA cores: K = 3 x 81)--> X0
//   16 x ResidualBlock: T = PReLU(BN(conv3x3(X)));  X' = X + BN(conv3x3(T))          conv_halo2_kernel<1>
//   Z = X0 + BN(conv3x3(X))                                                          conv_halo2_kernel<1>
//   3 (x8) or 4 (x16) x PixelShuffleBlock: PReLU(shuffle2(conv3x3 64->256))          conv_halo2_kernel<1>, 2 launches
//   y = tanh(conv 9x9 64->3)  -> [B,3,f*h,f*w] fp32                                  conv_halo2_kernel<2>
//
// Layout: fp16 NHWC, 64 channels, a batch is ONE tall pixel grid: image b occupies rows [b (H+G), b (H+G) + H), the G
// rows between images stay zero (G = 1 before a 3x3 conv, 4 before the 9x9 conv) and ARE the convolution's zero
// padding; left / right / top / bottom padding comes from TMA's out-of-bounds zero fill.  So every convolution of the
// batch is one launch over one tensor map, with no per-image loop and no halo-writing pass.
// Eval-mode BatchNorm is folded into the fp16 weights (per-output-channel scale) and an fp32 bias vector at load time;
// PixelShuffle is an output ADDRESSING mode: launch dy (weight rows permuted to n = dx * 64 + c) writes the 128
// contiguous channels of output pixels (2y + dy, 2x), (2y + dy, 2x + 1).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dsr_b200.h"
#include "dsr_host.h"
#include "dsr_launch.cuh"
#include "dsr_ptx.cuh"

namespace dsr {

constexpr int kGF = 64;                 // generator feature channels (generator.py:7,47)
constexpr float kGenBnEps = 1e-5f;

// ---------------------------------------------------------------------------------------------
// weight packing (load time): fold eval-mode BatchNorm, convert to fp16 GEMM layout [tap][n][ci]
//   mode 0: n = co;  mode 1 (PixelShuffle conv, 2 launches dy): out [dy][tap][128][ci], co = (n & 63) * 4 + dy * 2 + (n >> 6);
//   mode 2 (9x9 output conv, tap-by-tap kernel): [tap][16][ci], rows n >= cout are zero;
//   mode 3 (9x9 output conv, conv9_out_kernel): ktaps = 9 row taps, [ky][32][ci] with n = kx * 3 + co (27 rows used).
// ---------------------------------------------------------------------------------------------
__global__ void gen_pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                     const float* __restrict__ bn, __half* __restrict__ wout, float* __restrict__ bout,
                                     int cout, int cin, int ktaps, int mode) {
  const int n_rows = (mode == 1) ? 128 : (mode == 2 ? 16 : (mode == 3 ? 32 : cout));
  const int n_dy = (mode == 1) ? 2 : 1;
  const long long total = static_cast<long long>(n_dy) * ktaps * n_rows * cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    const int n = static_cast<int>((i / cin) % n_rows);
    const int tap = static_cast<int>((i / (static_cast<long long>(cin) * n_rows)) % ktaps);
    const int dy = static_cast<int>(i / (static_cast<long long>(cin) * n_rows * ktaps));
    int co = n;
    if (mode == 1) co = (n & 63) * 4 + dy * 2 + (n >> 6);
    float v = 0.f;
    if (mode == 3) {
      const int kx = n / 3;
      co = n - 3 * kx;
      if (n < 27) v = w[(static_cast<long long>(co) * cin + ci) * 81 + tap * 9 + kx];
      if (n >= 3) co = cout;                        // bias slots: n < 3 only
    } else if (co < cout) {
      v = w[(static_cast<long long>(co) * cin + ci) * ktaps + tap];
      if (bn != nullptr) v *= bn[co] * rsqrtf(bn[3 * cout + co] + kGenBnEps);
    }
    wout[i] = __float2half_rn(v);
    if (ci == 0 && tap == 0) {
      float bb = 0.f;
      if (co < cout) {
        bb = b[co];
        if (bn != nullptr) bb = (bb - bn[2 * cout + co]) * (bn[co] * rsqrtf(bn[3 * cout + co] + kGenBnEps)) + bn[cout + co];
      }
      bout[dy * n_rows + n] = bb;
    }
  }
}

// conv1 weights [64][3][9][9] -> [tap = ci * 81 + ky * 9 + kx][64] fp32 (the CUDA-core kernel broadcasts rows of 64)
__global__ void gen_pack_conv1_kernel(const float* __restrict__ w, float* __restrict__ wout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 243 * kGF) return;
  const int co = i % kGF, tap = i / kGF;
  wout[i] = w[co * 243 + tap];
}

__global__ void gen_copy_scalars_kernel(const float* __restrict__ state, const long long* __restrict__ offs,
                                        float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = state[offs[i]];
}

// ---------------------------------------------------------------------------------------------
// conv1: 9x9, 3 -> 64, zero padding 4, + PReLU, fp32 NCHW in -> fp16 tall NHWC out.  K = 243 with 3-channel pixels
// is no tensor-core shape (0.3 % of the network's FLOPs): one thread = one pixel x 64 output channels, the 16 x 8
// pixel block's 24 x 16 x 3 input patch and the whole weight table (62 KB) live in shared memory, weight rows are
// broadcast LDS.128.
// ---------------------------------------------------------------------------------------------
constexpr int kC1TX = 16, kC1TY = 8;
constexpr int kC1PW = kC1TX + 8, kC1PH = kC1TY + 8;
constexpr int kC1Smem = (243 * kGF + 3 * kC1PW * kC1PH) * 4;

__global__ void __launch_bounds__(kC1TX* kC1TY) gen_conv1_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                                  const float* __restrict__ bias,
                                                                  const float* __restrict__ slope_p, __half* __restrict__ out,
                                                                  int h, int w, int img_rows) {
  extern __shared__ float c1_smem[];
  float* wsm = c1_smem;                       // [243][64]
  float* patch = c1_smem + 243 * kGF;         // [3][kC1PH][kC1PW]
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kC1TX, y0 = blockIdx.y * kC1TY;
  for (int i = tid; i < 243 * kGF / 4; i += kC1TX * kC1TY)
    reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
  pdl_sync();
  const float* xb = x + static_cast<long long>(b) * 3 * h * w;
  for (int i = tid; i < 3 * kC1PH * kC1PW; i += kC1TX * kC1TY) {
    const int px = i % kC1PW, py = (i / kC1PW) % kC1PH, c = i / (kC1PW * kC1PH);
    const int gx = x0 + px - 4, gy = y0 + py - 4;
    patch[i] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? __ldg(xb + (static_cast<long long>(c) * h + gy) * w + gx) : 0.f;
  }
  __syncthreads();
  const int tx = tid % kC1TX, ty = tid / kC1TX;
  float acc[kGF];
#pragma unroll
  for (int i = 0; i < kGF; ++i) acc[i] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int ky = 0; ky < 9; ++ky) {
      const float* prow = patch + (c * kC1PH + ty + ky) * kC1PW + tx;
      const float4* wrow = reinterpret_cast<const float4*>(wsm + (c * 81 + ky * 9) * kGF);
#pragma unroll
      for (int kx = 0; kx < 9; ++kx) {
        const float v = prow[kx];
#pragma unroll
        for (int j = 0; j < kGF / 4; ++j) {
          const float4 w4 = wrow[kx * (kGF / 4) + j];
          acc[4 * j + 0] = fmaf(v, w4.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(v, w4.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, w4.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, w4.w, acc[4 * j + 3]);
        }
      }
    }
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx < w && gy < h) {
    const float slope = __ldg(slope_p);
    uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * img_rows + gy) * w + gx) * kGF);
#pragma unroll
    for (int j = 0; j < kGF / 8; ++j) {
      uint32_t pk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a0 = acc[8 * j + 2 * k] + __ldg(bias + 8 * j + 2 * k);
        float a1 = acc[8 * j + 2 * k + 1] + __ldg(bias + 8 * j + 2 * k + 1);
        a0 = a0 > 0.f ? a0 : slope * a0;
        a1 = a1 > 0.f ? a1 : slope * a1;
        __half2 hh = __floats2half2_rn(a0, a1);
        pk[k] = *reinterpret_cast<uint32_t*>(&hh);
      }
      dst[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

}  // namespace dsr

using namespace dsr;

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct GenTensorInfo {
  std::string name;
  long long offset, numel;
};

struct GenTall {             // one activation tensor: fp16 [rows][W][64], image period H + gap
  __half* ptr = nullptr;
  int H = 0, W = 0, gap = 1, rows = 0;
  size_t off = 0, bytes = 0;
};

struct dsr_gen_plan {
  int factor = 8, nres = 16, nshuf = 3, B = 1, h = 0, w = 0, num_sms = 148;
  std::vector<GenTensorInfo> tensors;
  long long numel = 0;
  // state offsets
  long long o_conv1_w = 0, o_conv1_b = 0, o_prelu1 = 0, o_conv2_w = 0, o_conv2_b = 0, o_bn1 = 0, o_conv3_w = 0, o_conv3_b = 0;
  std::vector<long long> o_rc1_w, o_rc1_b, o_rbn1, o_rprelu, o_rc2_w, o_rc2_b, o_rbn2, o_sc_w, o_sc_b, o_sprelu;
  // workspace
  size_t ws_bytes = 0;
  uint8_t* ws = nullptr;
  bool bound = false, loaded = false;
  bool conv3_taps = false;       // DSR_GEN_CONV3_TAPS=1: tap-by-tap 9x9 kernel (conv_halo2_kernel<2>) instead of conv9_out_kernel
  GenTall x0, xa, xb, t1, s[4];
  size_t off_w33 = 0, off_wsh = 0, off_w3 = 0, off_bias = 0, off_w1 = 0, off_slopes = 0, off_slope_offs = 0, off_err = 0;
  __half *w33 = nullptr, *wsh = nullptr, *w3 = nullptr;
  float *bias = nullptr, *w1 = nullptr, *slopes = nullptr;
  long long* slope_offs = nullptr;
  int* err = nullptr;
  std::vector<HaloParams> launches;
  int last_launches = 0;
};

namespace {

size_t up1k(size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); }

void gen_layout(dsr_gen_plan* p) {
  long long off = 0;
  auto add = [&](const std::string& name, long long n) {
    p->tensors.push_back(GenTensorInfo{name, off, n});
    const long long o = off;
    off += n;
    return o;
  };
  auto add_bn = [&](const std::string& pre) {
    const long long o = add(pre + ".weight", kGF);
    add(pre + ".bias", kGF);
    add(pre + ".running_mean", kGF);
    add(pre + ".running_var", kGF);
    return o;
  };
  // state_dict order of the reference module (generator.py:45-66; registration order of the attributes)
  p->o_conv1_w = add("conv1.weight", kGF * 3 * 81);
  p->o_conv1_b = add("conv1.bias", kGF);
  p->o_prelu1 = add("prelu1.weight", 1);
  for (int i = 0; i < p->nres; ++i) {
    const std::string pre = "residual_blocks." + std::to_string(i) + ".";
    p->o_rc1_w.push_back(add(pre + "conv1.weight", kGF * kGF * 9));
    p->o_rc1_b.push_back(add(pre + "conv1.bias", kGF));
    p->o_rbn1.push_back(add_bn(pre + "bn1"));
    p->o_rprelu.push_back(add(pre + "prelu1.weight", 1));
    p->o_rc2_w.push_back(add(pre + "conv2.weight", kGF * kGF * 9));
    p->o_rc2_b.push_back(add(pre + "conv2.bias", kGF));
    p->o_rbn2.push_back(add_bn(pre + "bn2"));
  }
  p->o_conv2_w = add("conv2.weight", kGF * kGF * 9);
  p->o_conv2_b = add("conv2.bias", kGF);
  p->o_bn1 = add_bn("bn1");
  for (int i = 0; i < p->nshuf; ++i) {
    const std::string pre = "pixel_shuffle_blocks." + std::to_string(i) + ".";
    p->o_sc_w.push_back(add(pre + "conv1.weight", 4 * kGF * kGF * 9));
    p->o_sc_b.push_back(add(pre + "conv1.bias", 4 * kGF));
    p->o_sprelu.push_back(add(pre + "prelu1.weight", 1));
  }
  p->o_conv3_w = add("conv3.weight", 3 * kGF * 81);
  p->o_conv3_b = add("conv3.bias", 3);
  p->numel = off;
}

void tall_init(GenTall& t, int B, int H, int W, int gap, size_t& off) {
  t.H = H;
  t.W = W;
  t.gap = gap;
  t.rows = B * (H + gap) - gap;
  t.bytes = static_cast<size_t>(t.rows) * W * kGF * sizeof(__half);
  t.off = off;
  off += up1k(t.bytes);
}

// bias arena (floats): [2 nres + 1][64] | [nshuf][2][128] | [16]
long long bias_off_33(int i) { return static_cast<long long>(i) * kGF; }
long long bias_off_sh(const dsr_gen_plan* p, int i, int dy) { return (2LL * p->nres + 1) * kGF + (2LL * i + dy) * 128; }
long long bias_off_3(const dsr_gen_plan* p) { return (2LL * p->nres + 1) * kGF + 2LL * p->nshuf * 128; }

// One convolution launch over a tall input tensor.
int gen_make_conv(dsr_gen_plan* p, HaloParams& h, const GenTall& in, const __half* wts, int N, int k, int ep_mode) {
  memset(&h, 0, sizeof(h));
  int rc;
  h.n_wide = 1;
  h.n_narrow = 0;
  h.n_part = N / 2;
  h.parts = 2;
  h.ntaps = k * k;
  h.halo_w = kHaloTW + (k - 1);
  h.halo_h = kHaloTH + (k - 1);
  h.wide_slot_bytes = static_cast<int>(up1k(static_cast<size_t>(h.halo_w) * h.halo_h * 128));
  // one 64-channel chunk per tile: the A ring depth is the number of TILES in flight -- deep enough to cover the TMA
  // latency at ~0.7 us of MMA work per tile (DSR_GEN_SLOTS for A/B)
  h.wide_slots = (k == 3) ? (N <= 64 ? 6 : 4) : 2;
  if (k == 3 && getenv("DSR_GEN_SLOTS")) h.wide_slots = atoi(getenv("DSR_GEN_SLOTS"));
  if ((rc = make_act_map(&h.a64, in.ptr, 1, kGF, in.W, in.rows, 1, 64, h.halo_w, h.halo_h))) return rc;
  h.a16 = h.a64;
  if ((rc = make_wgt_map(&h.b64, wts, kGF, h.ntaps * N, 64, h.n_part))) return rc;
  h.b16 = h.b64;
  if (k == 3)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx)
        h.taps[ky * 3 + kx] = HaloTap{static_cast<int8_t>(ky), static_cast<int8_t>(kx), static_cast<int16_t>((ky * 3 + kx) * N)};
  h.org_x = h.org_y = -(k - 1) / 2;
  h.tiles_x = (in.W + kHaloTW - 1) / kHaloTW;
  h.tiles_y = (in.rows + kHaloTH - 1) / kHaloTH;
  h.out_h = in.rows;
  h.out_w = in.W;
  h.img_rows = in.H + in.gap;
  h.img_h = in.H;
  h.n_store = N;
  h.stats = nullptr;
  h.idesc_wide = h.idesc_narrow = make_idesc_f16(256, N, FMT_F16, FMT_F16, 0, 0);
  h.smem_bytes = halo_smem_bytes(h.n_part, h.n_wide, h.n_narrow, h.wide_slots, h.ntaps, h.wide_slot_bytes);
  h.pair = 1;
  h.ep_mode = ep_mode;
  h.err = p->err;
  return 0;
}

void gen_same_out(HaloParams& h, const GenTall& out) {          // 64 -> 64: the output has the input's geometry
  h.out = out.ptr;
  h.out_img_stride = static_cast<long long>(out.H + out.gap) * out.W * kGF;
  h.out_sy = static_cast<long long>(out.W) * kGF;
  h.out_sx = kGF;
}

}  // namespace

extern "C" {

int dsr_gen_plan_create(dsr_gen_plan_t** out, int factor, int residual_blocks, int batch, int h, int w) {
  if (out == nullptr) return -1;
  *out = nullptr;
  if (factor != 8 && factor != 16) return -5;         // generator.py:55-58 builds nothing else
  if (residual_blocks < 1 || residual_blocks > 64 || batch < 1 || h < 8 || w < 8) return -1;
  const long long out_rows = static_cast<long long>(batch) * (static_cast<long long>(h) * factor + 4);
  if (out_rows > (1LL << 30) || static_cast<long long>(w) * factor > (1 << 20)) return -1;
  dsr_gen_plan* p = new dsr_gen_plan();
  p->factor = factor;
  p->nres = residual_blocks;
  p->nshuf = (factor == 8) ? 3 : 4;
  p->B = batch;
  p->h = h;
  p->w = w;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
    p->num_sms = prop.multiProcessorCount;
  p->conv3_taps = getenv("DSR_GEN_CONV3_TAPS") != nullptr;
  gen_layout(p);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += up1k(bytes);
    return o;
  };
  p->off_err = take(256);
  p->off_w33 = take(static_cast<size_t>(2 * p->nres + 1) * 9 * kGF * kGF * 2);
  p->off_wsh = take(static_cast<size_t>(p->nshuf) * 2 * 9 * 128 * kGF * 2);
  p->off_w3 = take(static_cast<size_t>(81) * 16 * kGF * 2);          // either layout: [81][16][64] or [9][32][64]
  p->off_bias = take(static_cast<size_t>(bias_off_3(p) + 32) * 4);
  p->off_w1 = take(static_cast<size_t>(243 + 1) * kGF * 4);          // [243][64] weights + [64] bias
  p->off_slopes = take(static_cast<size_t>(1 + p->nres + p->nshuf) * 4);
  p->off_slope_offs = take(static_cast<size_t>(1 + p->nres + p->nshuf) * 8);
  tall_init(p->x0, batch, h, w, 1, off);
  tall_init(p->xa, batch, h, w, 1, off);
  tall_init(p->xb, batch, h, w, 1, off);
  tall_init(p->t1, batch, h, w, 1, off);
  int H = h, W = w;
  for (int i = 0; i < p->nshuf; ++i) {
    H *= 2;
    W *= 2;
    tall_init(p->s[i], batch, H, W, (i == p->nshuf - 1) ? 4 : 1, off);
  }
  p->ws_bytes = off;
  *out = p;
  return 0;
}

void dsr_gen_plan_destroy(dsr_gen_plan_t* p) { delete p; }

long long dsr_gen_state_numel(const dsr_gen_plan_t* p) { return p ? p->numel : -1; }
int dsr_gen_num_tensors(const dsr_gen_plan_t* p) { return p ? static_cast<int>(p->tensors.size()) : -1; }
int dsr_gen_tensor_info(const dsr_gen_plan_t* p, int idx, char* name, int name_cap, long long* offset, long long* numel) {
  if (p == nullptr || idx < 0 || idx >= static_cast<int>(p->tensors.size())) return -1;
  const GenTensorInfo& t = p->tensors[idx];
  if (name != nullptr && name_cap > 0) {
    strncpy(name, t.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (offset) *offset = t.offset;
  if (numel) *numel = t.numel;
  return 0;
}
size_t dsr_gen_workspace_bytes(const dsr_gen_plan_t* p) { return p ? p->ws_bytes : 0; }

int dsr_gen_bind(dsr_gen_plan_t* p, void* workspace, size_t bytes, void* stream) {
  if (p == nullptr || workspace == nullptr) return -1;
  if (bytes < p->ws_bytes) return -8;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return -3;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, p->ws_bytes, s);      // gap rows must be (and stay) zero
  if (e != cudaSuccess) return static_cast<int>(e);
  p->ws = static_cast<uint8_t*>(workspace);
  p->err = reinterpret_cast<int*>(p->ws + p->off_err);
  p->w33 = reinterpret_cast<__half*>(p->ws + p->off_w33);
  p->wsh = reinterpret_cast<__half*>(p->ws + p->off_wsh);
  p->w3 = reinterpret_cast<__half*>(p->ws + p->off_w3);
  p->bias = reinterpret_cast<float*>(p->ws + p->off_bias);
  p->w1 = reinterpret_cast<float*>(p->ws + p->off_w1);
  p->slopes = reinterpret_cast<float*>(p->ws + p->off_slopes);
  p->slope_offs = reinterpret_cast<long long*>(p->ws + p->off_slope_offs);
  GenTall* all[8] = {&p->x0, &p->xa, &p->xb, &p->t1, &p->s[0], &p->s[1], &p->s[2], &p->s[3]};
  for (GenTall* t : all)
    if (t->bytes) t->ptr = reinterpret_cast<__half*>(p->ws + t->off);

  // ---- the launch list ----
  p->launches.clear();
  int rc;
  const size_t w33_stride = static_cast<size_t>(9) * kGF * kGF;
  HaloParams hp;
  const GenTall* cur = &p->x0;
  GenTall* pp[2] = {&p->xa, &p->xb};
  int flip = 0;
  for (int i = 0; i < p->nres; ++i) {
    // T = PReLU(BN(conv1(X)))   (generator.py:15-18)
    if ((rc = gen_make_conv(p, hp, *cur, p->w33 + (2 * i) * w33_stride, kGF, 3, 1))) return rc;
    gen_same_out(hp, p->t1);
    hp.ep_bias = p->bias + bias_off_33(2 * i);
    hp.ep_slope = p->slopes + 1 + i;
    p->launches.push_back(hp);
    // X' = X + BN(conv2(T))     (generator.py:20-23)
    if ((rc = gen_make_conv(p, hp, p->t1, p->w33 + (2 * i + 1) * w33_stride, kGF, 3, 1))) return rc;
    gen_same_out(hp, *pp[flip]);
    hp.ep_bias = p->bias + bias_off_33(2 * i + 1);
    hp.ep_res = cur->ptr;
    p->launches.push_back(hp);
    cur = pp[flip];
    flip ^= 1;
  }
  // Z = X0 + BN(conv2(X))        (generator.py:73-76)
  if ((rc = gen_make_conv(p, hp, *cur, p->w33 + (2 * p->nres) * w33_stride, kGF, 3, 1))) return rc;
  gen_same_out(hp, *pp[flip]);
  hp.ep_bias = p->bias + bias_off_33(2 * p->nres);
  hp.ep_res = p->x0.ptr;
  p->launches.push_back(hp);
  cur = pp[flip];
  // PixelShuffle blocks          (generator.py:36-41)
  for (int i = 0; i < p->nshuf; ++i) {
    const GenTall& o = p->s[i];
    for (int dy = 0; dy < 2; ++dy) {
      const __half* wts = p->wsh + (static_cast<size_t>(2 * i + dy)) * 9 * 128 * kGF;
      if ((rc = gen_make_conv(p, hp, *cur, wts, 128, 3, 1))) return rc;
      hp.out = o.ptr + static_cast<long long>(dy) * o.W * kGF;
      hp.out_img_stride = static_cast<long long>(o.H + o.gap) * o.W * kGF;
      hp.out_sy = 2LL * o.W * kGF;
      hp.out_sx = 2 * kGF;
      hp.ep_bias = p->bias + bias_off_sh(p, i, dy);
      hp.ep_slope = p->slopes + 1 + p->nres + i;
      p->launches.push_back(hp);
    }
    cur = &p->s[i];
  }
  // y = tanh(conv3(.))           (generator.py:80-82); output pointer patched per call
  if (p->conv3_taps) {
    if ((rc = gen_make_conv(p, hp, *cur, p->w3, 16, 9, 2))) return rc;
  } else {                       // conv9_out_kernel: kx folded into N, 24 x 8 pixel tiles, 32 x 16 halo boxes
    if ((rc = gen_make_conv(p, hp, *cur, p->w3, 16, 9, 3))) return rc;
    if ((rc = make_act_map(&hp.a64, cur->ptr, 1, kGF, cur->W, cur->rows, 1, 64, 32, 16))) return rc;
    if ((rc = make_wgt_map(&hp.b64, p->w3, kGF, 9 * 32, 64, 16))) return rc;
    hp.a16 = hp.a64;
    hp.b16 = hp.b64;
    hp.tiles_x = (cur->W + 23) / 24;
    hp.tiles_y = (cur->rows + 7) / 8;
    hp.idesc_wide = hp.idesc_narrow = make_idesc_f16(256, 32, FMT_F16, FMT_F16, 0, 0);
  }
  hp.out = nullptr;
  hp.out_img_stride = 3LL * cur->H * cur->W;
  hp.out_sy = cur->W;
  hp.out_sx = 1;
  hp.ep_plane = static_cast<long long>(cur->H) * cur->W;
  hp.n_store = 3;
  hp.ep_bias = p->bias + bias_off_3(p);
  p->launches.push_back(hp);
  p->bound = true;
  p->loaded = false;
  return 0;
}

int dsr_gen_load_weights(dsr_gen_plan_t* p, const float* st, void* stream) {
  if (p == nullptr || st == nullptr) return -1;
  if (!p->bound) return -6;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t w33_stride = static_cast<size_t>(9) * kGF * kGF;
  auto pack = [&](const float* w, const float* b, const float* bn, __half* wout, float* bout, int cout, int ktaps, int mode) {
    const long long total = static_cast<long long>(mode == 1 ? 2 * 128 : (mode == 2 ? 16 : (mode == 3 ? 32 : cout))) * ktaps * kGF;
    const int blocks = static_cast<int>((total + 255) / 256);
    gen_pack_conv_kernel<<<blocks, 256, 0, s>>>(w, b, bn, wout, bout, cout, kGF, ktaps, mode);
  };
  for (int i = 0; i < p->nres; ++i) {
    pack(st + p->o_rc1_w[i], st + p->o_rc1_b[i], st + p->o_rbn1[i], p->w33 + (2 * i) * w33_stride,
         p->bias + bias_off_33(2 * i), kGF, 9, 0);
    pack(st + p->o_rc2_w[i], st + p->o_rc2_b[i], st + p->o_rbn2[i], p->w33 + (2 * i + 1) * w33_stride,
         p->bias + bias_off_33(2 * i + 1), kGF, 9, 0);
  }
  pack(st + p->o_conv2_w, st + p->o_conv2_b, st + p->o_bn1, p->w33 + (2 * p->nres) * w33_stride,
       p->bias + bias_off_33(2 * p->nres), kGF, 9, 0);
  for (int i = 0; i < p->nshuf; ++i)
    pack(st + p->o_sc_w[i], st + p->o_sc_b[i], nullptr, p->wsh + static_cast<size_t>(2 * i) * 9 * 128 * kGF,
         p->bias + bias_off_sh(p, i, 0), 4 * kGF, 9, 1);
  if (p->conv3_taps) pack(st + p->o_conv3_w, st + p->o_conv3_b, nullptr, p->w3, p->bias + bias_off_3(p), 3, 81, 2);
  else pack(st + p->o_conv3_w, st + p->o_conv3_b, nullptr, p->w3, p->bias + bias_off_3(p), 3, 9, 3);
  gen_pack_conv1_kernel<<<(243 * kGF + 255) / 256, 256, 0, s>>>(st + p->o_conv1_w, p->w1);
  std::vector<long long> offs;
  offs.push_back(p->o_prelu1);
  for (int i = 0; i < p->nres; ++i) offs.push_back(p->o_rprelu[i]);
  for (int i = 0; i < p->nshuf; ++i) offs.push_back(p->o_sprelu[i]);
  cudaError_t e = cudaMemcpyAsync(p->slope_offs, offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaStreamSynchronize(s);                      // `offs` is a host temporary
  if (e != cudaSuccess) return static_cast<int>(e);
  gen_copy_scalars_kernel<<<1, 128, 0, s>>>(st, p->slope_offs, p->slopes, static_cast<int>(offs.size()));
  // conv1's bias sits right behind its repacked weights
  e = cudaMemcpyAsync(p->w1 + 243 * kGF, st + p->o_conv1_b, kGF * sizeof(float), cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  p->loaded = true;
  return 0;
}

int dsr_gen_forward(dsr_gen_plan_t* p, const float* x, float* y, void* stream) {
  if (p == nullptr || x == nullptr || y == nullptr) return -1;
  if (!p->bound || !p->loaded) return -6;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static bool attr_done_dev[kMaxDevices] = {};
  bool& attr_done = attr_done_dev[device_slot()];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gen_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1Smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  int n = 0;
  dim3 g1((p->w + kC1TX - 1) / kC1TX, (p->h + kC1TY - 1) / kC1TY, p->B);
  launch_k(gen_conv1_kernel, g1, dim3(kC1TX * kC1TY), kC1Smem, s, x, static_cast<const float*>(p->w1),
           static_cast<const float*>(p->w1 + 243 * kGF), static_cast<const float*>(p->slopes), p->x0.ptr, p->h, p->w,
           p->h + p->x0.gap);
  ++n;
  for (size_t i = 0; i < p->launches.size(); ++i) {
    HaloParams hp = p->launches[i];
    if (hp.ep_mode >= 2) hp.out = y;
    const int rc = launch_conv_halo(hp, p->num_sms, s);
    if (rc) return rc;
    ++n;
  }
  p->last_launches = n;
  return static_cast<int>(cudaGetLastError());
}

int dsr_gen_last_launches(const dsr_gen_plan_t* p) { return p ? p->last_launches : -1; }

int dsr_gen_debug_tensor(const dsr_gen_plan_t* p, const char* name, void** ptr, int* rows, int* img_rows, int* H, int* W) {
  if (p == nullptr || name == nullptr || !p->bound) return -1;
  const GenTall* t = nullptr;
  if (!strcmp(name, "x0")) t = &p->x0;
  else if (!strcmp(name, "xa")) t = &p->xa;
  else if (!strcmp(name, "xb")) t = &p->xb;
  else if (!strcmp(name, "t1")) t = &p->t1;
  else if (name[0] == 's' && name[1] >= '0' && name[1] < '0' + p->nshuf && name[2] == 0) t = &p->s[name[1] - '0'];
  if (t == nullptr) return -1;
  if (ptr) *ptr = t->ptr;
  if (rows) *rows = t->rows;
  if (img_rows) *img_rows = t->H + t->gap;
  if (H) *H = t->H;
  if (W) *W = t->W;
  return 0;
}

int dsr_gen_device_error(dsr_gen_plan_t* p, int* host_code) {
  if (p == nullptr || host_code == nullptr || !p->bound) return -1;
  cudaError_t e = cudaMemcpy(host_code, p->err, sizeof(int), cudaMemcpyDeviceToHost);
  return static_cast<int>(e);
}

}  // extern "C"
