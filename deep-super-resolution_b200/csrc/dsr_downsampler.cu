// dsr_downsampler.cu -- the fixed Lanczos anti-aliasing downsampler of the DIP loss, fp32.
//
// Replaces utils/downsampler.py:5-71 (Downsampler: ReplicationPad2d + dense Conv2d(n, n, k, stride=f)
// whose weight is the 2-D Lanczos table on the diagonal) and its autograd backward.  The 2-D table
// of utils/downsampler.py:103-133 is an outer product of 1-D taps, so the kernel is evaluated
// separably: a block stages the clamped input patch in shared memory, filters rows, then columns.
// The backward pass is a gather with host-precomputed per-row / per-column weight tables that fold
// the replication padding (border pixels collect the gradient of every padded copy).
#include <math.h>

#include <vector>

#include "dsr_elem.cuh"
#include "dsr_launch.cuh"

namespace dsr {

namespace {

constexpr float kLossScale = 1099511627776.f;      // 2^40: loss quanta of 9e-13, range +-8e6

// outputs per block: 16 x 8 (factor <= 8; 384 blocks at 512^2 / factor 4) or 16 x 4 (larger factors, to keep the staged patch in smem)
inline void ds_tile(int factor, int& tox, int& toy) {
  if (factor <= 8) { tox = 16; toy = 8; } else { tox = 16; toy = 4; }
}

// dynamic smem: patch [py][px] then rows-filtered [py][kTileOx]
__global__ void downsample_fwd_kernel(const float* __restrict__ x, const float* __restrict__ target,
                                      float* __restrict__ y, float* __restrict__ gy, float* __restrict__ loss, int C,
                                      int H, int W, int oh, int ow, DsTables t, int kTileOx, int kTileOy,
                                      const float* __restrict__ state) {
  pdl_sync();
  extern __shared__ float sm[];
  if (state != nullptr && loss != nullptr) loss += __float_as_int(state[0]) - 1;
  const int f = t.factor, k = t.k, pad = t.pad;
  const int pw = (kTileOx - 1) * f + k;          // patch width (input pixels)
  const int ph = (kTileOy - 1) * f + k;
  float* patch = sm;                              // [ph][pw]
  float* rowf = sm + ph * pw;                     // [ph][kTileOx]
  float* taps = rowf + ph * kTileOx;              // [k]
  const int c = blockIdx.z;
  const int ox0 = blockIdx.x * kTileOx, oy0 = blockIdx.y * kTileOy;
  const float* xc = x + static_cast<long long>(c) * H * W;
  for (int i = threadIdx.x; i < k; i += blockDim.x) taps[i] = t.taps[i];
#pragma unroll 8
  for (int i = threadIdx.x; i < ph * pw; i += blockDim.x) {      // unrolled: 8 independent clamped loads in flight
    const int py = i / pw, px = i - py * pw;
    int sy = oy0 * f + py - pad, sx = ox0 * f + px - pad;    // replicate padding = clamp
    sy = min(max(sy, 0), H - 1);
    sx = min(max(sx, 0), W - 1);
    patch[i] = xc[static_cast<long long>(sy) * W + sx];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ph * kTileOx; i += blockDim.x) {
    const int py = i / kTileOx, ox = i % kTileOx;
    float acc = 0.f;
    for (int j = 0; j < k; ++j) acc = fmaf(taps[j], patch[py * pw + ox * f + j], acc);
    rowf[i] = acc;
  }
  __syncthreads();
  float lsum = 0.f;
  for (int i = threadIdx.x; i < kTileOy * kTileOx; i += blockDim.x) {
    const int oyl = i / kTileOx, ox = i % kTileOx;
    const int oy = oy0 + oyl, oxg = ox0 + ox;
    if (oy < oh && oxg < ow) {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc = fmaf(taps[j], rowf[(oyl * f + j) * kTileOx + ox], acc);
      const long long o = (static_cast<long long>(c) * oh + oy) * ow + oxg;
      y[o] = acc;
      if (target != nullptr) {
        const float d = acc - target[o];
        const float inv_n = 1.f / (static_cast<float>(C) * oh * ow);
        gy[o] = 2.f * d * inv_n;
        lsum += d * d * inv_n;
      }
    }
  }
  if (target != nullptr) {
    // The loss is summed in 64-bit fixed point (2^-40 quanta) so that it does not depend on the order in which the
    // blocks finish: every block adds its sum to t.loss_acc; the last one to arrive (ticket) converts the total,
    // adds it to *loss and re-arms the two words for the next launch.
    __shared__ acc_t red;
    if (threadIdx.x == 0) red = 0ull;
    __syncthreads();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red, acc_fix(lsum, kLossScale));
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(t.loss_acc, red);
      __threadfence();
      const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
      if (atomicAdd(t.loss_ticket, 1u) == nblocks - 1) {
        __threadfence();
        const acc_t total = atomicExch(t.loss_acc, 0ull);
        *t.loss_ticket = 0u;
        *loss += acc_val(total, 1.f / kLossScale);
      }
    }
  }
}

__global__ void downsample_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int C, int H, int W, int oh,
                                      int ow, DsTables t) {
  pdl_sync();
  const int n = C * H * W;                      // < 2^31 (checked by the launcher)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / (W * H), rem = i - c * (W * H);
    const int y = rem / W, x = rem - y * W;
    const int y0 = t.by0[y], x0 = t.bx0[x];
    const float* g = gy + static_cast<long long>(c) * oh * ow;
    float acc = 0.f;
    for (int a = 0; a < t.nw; ++a) {
      const float wy = t.bwy[y * t.nw + a];
      const int oy = y0 + a;
      if (wy == 0.f || oy >= oh) continue;
      float r = 0.f;
      for (int b = 0; b < t.nw; ++b) {
        const int ox = x0 + b;
        if (ox < ow) r = fmaf(t.bwx[x * t.nw + b], g[static_cast<long long>(oy) * ow + ox], r);
      }
      acc = fmaf(wy, r, acc);
    }
    gx[i] = acc;
  }
}

}  // namespace

static size_t ds_smem(const DsTables& t) {
  int kTileOx, kTileOy;
  ds_tile(t.factor, kTileOx, kTileOy);
  const int pw = (kTileOx - 1) * t.factor + t.k, ph = (kTileOy - 1) * t.factor + t.k;
  return sizeof(float) * (static_cast<size_t>(ph) * pw + static_cast<size_t>(ph) * kTileOx + t.k);
}

static int ds_launch(const float* x, const float* target, float* y, float* gy, float* loss, int C, int H, int W,
                     int oh, int ow, DsTables t, cudaStream_t s, const float* state = nullptr) {
  const size_t smem = ds_smem(t);
  if (smem > 200 * 1024) return -4;
  static size_t configured_dev[64] = {};            // function attributes are per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  size_t& configured = configured_dev[dev];
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(downsample_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = smem;
  }
  int kTileOx, kTileOy;
  ds_tile(t.factor, kTileOx, kTileOy);
  dim3 grid((ow + kTileOx - 1) / kTileOx, (oh + kTileOy - 1) / kTileOy, C);
  launch_k(downsample_fwd_kernel, dim3(grid), dim3(256), smem, s, x, target, y, gy, loss, C, H, W, oh, ow, t, kTileOx, kTileOy, state);
  return static_cast<int>(cudaGetLastError());
}

int launch_downsample_fwd(const float* x, float* y, int C, int H, int W, int oh, int ow, DsTables t, cudaStream_t s) {
  return ds_launch(x, nullptr, y, nullptr, nullptr, C, H, W, oh, ow, t, s);
}
int launch_downsample_mse(const float* x, const float* target, float* y, float* gy, float* loss, int C, int H, int W,
                          int oh, int ow, DsTables t, cudaStream_t s, const float* state) {
  return ds_launch(x, target, y, gy, loss, C, H, W, oh, ow, t, s, state);
}
int launch_downsample_bwd(const float* gy, float* gx, int C, int H, int W, int oh, int ow, DsTables t, cudaStream_t s) {
  const long long n = static_cast<long long>(C) * H * W;
  if (n >= (1LL << 31)) return -5;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(downsample_bwd_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, s, gy, gx, C, H, W, oh, ow, t);
  return static_cast<int>(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// Host-side tables
// ---------------------------------------------------------------------------------------------
// 1-D taps: utils/downsampler.py:103-127 ('lanczos', phase 0.5): the 2-D entry is the product of a
// row factor and a column factor; normalising the outer product by its sum (:133) equals
// normalising each factor by its own sum.
void lanczos_taps(int factor, int support, std::vector<double>& taps) {
  const int kernel_width = 2 * support * factor + 1;     // 4f+1 (lanczos2) / 6f+1 (lanczos3), :14-22
  const int n = kernel_width - 1;                        // phase 0.5 shrinks the table, :77-78
  const double center = (kernel_width + 1) / 2.0;
  const double pi = 3.14159265358979323846;
  taps.assign(n, 1.0);
  double sum = 0.0;
  for (int i = 1; i <= n; ++i) {
    const double d = fabs(i + 0.5 - center) / factor;
    double v = 1.0;
    if (d != 0.0) v = support * sin(pi * d) * sin(pi * d / support) / (pi * pi * d * d);
    taps[i - 1] = v;
    sum += v;
  }
  for (int i = 0; i < n; ++i) taps[i] /= sum;
}

// Backward gather table for one axis of length n (output length on): for input index i, the first
// output index o0 and weights w[a] such that  gx[i] = sum_a w[a] * gy[o0 + a].
void ds_bwd_table(int n, int on, int factor, int k, int pad, const std::vector<float>& taps, int& nw,
                  std::vector<int>& o0, std::vector<float>& w) {
  // a padded position p (0 <= p < n + 2 pad) belongs to input i = clamp(p - pad); it is read by output o with
  // tap p - f*o when 0 <= p - f*o < k.
  std::vector<std::vector<std::pair<int, double>>> lists(n);
  int maxspan = 1;
  for (int i = 0; i < n; ++i) {
    int plo = i + pad, phi = i + pad;
    if (i == 0) plo = 0;
    if (i == n - 1) phi = n + 2 * pad - 1;
    int omin = 1 << 30, omax = -1;
    std::vector<double> acc(on, 0.0);
    for (int p = plo; p <= phi; ++p)
      for (int o = 0; o < on; ++o) {
        const int tp = p - factor * o;
        if (tp >= 0 && tp < k) {
          acc[o] += taps[tp];
          omin = o < omin ? o : omin;
          omax = o > omax ? o : omax;
        }
      }
    if (omax >= 0) {
      for (int o = omin; o <= omax; ++o) lists[i].push_back({o, acc[o]});
      if (omax - omin + 1 > maxspan) maxspan = omax - omin + 1;
    }
  }
  nw = maxspan;
  o0.assign(n, 0);
  w.assign(static_cast<size_t>(n) * nw, 0.f);
  for (int i = 0; i < n; ++i) {
    if (lists[i].empty()) continue;
    o0[i] = lists[i][0].first;
    for (size_t a = 0; a < lists[i].size(); ++a) w[static_cast<size_t>(i) * nw + a] = static_cast<float>(lists[i][a].second);
  }
}

DSR_KSTAMP_SETTER(kstamp_set_ds)

}  // namespace dsr
