// dsr_gant_conv.cu -- the two generic tensor-core kernels of the SRGAN training step (sm_100a, tcgen05 / TMEM / TMA):
//
//   gconv_kernel  : implicit-GEMM convolution, forward and data gradient, for every nn.Conv2d of the generator
//                   (models/GAN/generator.py:4-81), the discriminator (models/GAN/discriminator.py:4-74) and the VGG19
//                   feature extractor of the perceptual loss (utils/GAN.py:62-88): 3 x 3 stride 1 / 2 and 9 x 9, channel
//                   counts 3 (16-channel pitch), 64, 128, 256, 512.  bf16 operands, fp32 accumulation in TMEM; epilogue =
//                   bias, residual add, (Leaky)ReLU or its derivative mask, BatchNorm sum / sum of squares.
//   gwgrad_kernel : weight gradient of the same layers.
//
// Replaces torch.nn.Conv2d forward / backward (ATen / cuDNN) as called by the modules above inside
// train_GAN.py:38-71 (do_epoch).  Layout and conventions: dsr_gant.cuh.
#include "dsr_gant.cuh"
#include "dsr_ptx.cuh"
#include "dsr_host.h"
#include "dsr_launch.cuh"

#include <stdlib.h>

namespace dsr {

__device__ __forceinline__ GTap gc_tap(const GConvParams& p, int t) {
  if (p.proc_ks > 0) {
    GTap tp;
    const int ky = t / p.proc_ks, kx = t - ky * p.proc_ks;
    tp.px = 0; tp.py = 0;
    tp.dx = static_cast<int8_t>(p.proc_sign * (kx - p.proc_pad));
    tp.dy = static_cast<int8_t>(p.proc_sign * (ky - p.proc_pad));
    tp.b_row = t * p.rows_per_tap;
    return tp;
  }
  return p.taps[t];
}

__device__ __forceinline__ void gc_flush_stats(const GConvParams& p, int ntile, int lane, float (&acc_s)[8],
                                               float (&acc_q)[8]) {
  if (p.stats == nullptr || ntile < 0) return;
  const int nchunks = p.nt >> 4;
  const int ctot = p.ntiles_n * p.nt;
  if ((lane & 1) == 0) {
    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < nchunks) {
        atomicAdd(&p.stats[ntile * p.nt + c * 16 + col], static_cast<double>(acc_s[c]));
        atomicAdd(&p.stats[ctot + ntile * p.nt + c * 16 + col], static_cast<double>(acc_q[c]));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) { acc_s[c] = 0.f; acc_q[c] = 0.f; }
}

__global__ void __launch_bounds__(kGcThreads, 1) gconv_kernel(const __grid_constant__ GConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGcRingBytes);    // same offset in both modes
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGcStages;
  uint64_t* tfull_bar = bars + 2 * kGcStages;
  uint64_t* tempty_bar = bars + 2 * kGcStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGcStages + 4);
  const int nstages = p.rowhalo ? kGrStages : kGcStages;
  const int stage_bytes = p.rowhalo ? (p.a_slot + p.b_slot) : kGcStageBytes;
  const int a_off = p.rowhalo ? p.a_slot : kGcStageA;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles_pix = p.tiles_x * p.tiles_y;
  const int nitems = ntiles_pix * p.ntiles_n;
  const int kper = p.narrow ? 1 : p.kchunks;
  const uint32_t a_bytes = p.narrow ? 128u * 32u : 128u * 128u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.nt) * (p.narrow ? 32u : 128u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a);
    tma_prefetch_desc(&p.b);
    if (p.b2_loads > 0) tma_prefetch_desc(&p.b2);
    for (int s = 0; s < kGcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    (void)nstages;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tile = item % ntiles_pix;
        const int ntile = item / ntiles_pix;
        const int x0 = (tile % p.tiles_x) * p.tw;
        const int y0 = (tile / p.tiles_x) * p.th;
        if (p.rowhalo) {
          const int ks = p.proc_ks;
          const uint32_t box_bytes = static_cast<uint32_t>((8 + ks - 1) * 16) * (p.narrow ? 32u : 128u);
          for (int ky = 0; ky < ks; ++ky) {
            for (int kc = 0; kc < kper; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 21);
              uint8_t* sa = smem + stage * stage_bytes;
              uint8_t* sb = sa + a_off;
              mbar_arrive_expect_tx(&full_bar[stage], box_bytes + static_cast<uint32_t>(ks) * b_bytes);
              tma_load_5d(&p.a, &full_bar[stage], sa, kc * 64, 0, x0 - p.proc_pad, 0, y0 + p.proc_sign * (ky - p.proc_pad));
              if (p.b2_loads > 0) {        // one N tile: the row's ks weight blocks are contiguous rows of the pack
                const uint32_t part = static_cast<uint32_t>(p.b2_rows) * (p.narrow ? 32u : 128u);
                for (int l = 0; l < p.b2_loads; ++l)
                  tma_load_2d(&p.b2, &full_bar[stage], sb + l * part, kc * 64, ky * ks * p.rows_per_tap + l * p.b2_rows);
              } else {
                for (int kx = 0; kx < ks; ++kx)
                  tma_load_2d(&p.b, &full_bar[stage], sb + kx * b_bytes, kc * 64, (ky * ks + kx) * p.rows_per_tap + ntile * p.nt);
              }
              if (++stage == kGrStages) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
        for (int t = 0; t < p.ntaps; ++t) {
          const GTap tp = gc_tap(p, t);
          for (int kc = 0; kc < kper; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 21);
            uint8_t* sa = smem + stage * kGcStageBytes;
            uint8_t* sb = sa + kGcStageA;
            mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
            tma_load_5d(&p.a, &full_bar[stage], sa, kc * 64, tp.px, x0 + tp.dx, tp.py, y0 + tp.dy);
            tma_load_2d(&p.b, &full_bar[stage], sb, kc * 64, tp.b_row + ntile * p.nt);
            if (++stage == kGcStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const uint64_t hi_w = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint64_t hi_n = make_smem_desc(0, 16, 256, SWZ_32B);
    const uint64_t hi = p.narrow ? hi_n : hi_w;
    const int nk = p.rowhalo ? p.proc_ks * kper : p.ntaps * kper;
    const uint32_t rowb = p.narrow ? 32u : 128u;
    // row-halo views: 8-row groups are one box row (8 + ks - 1 pixels) apart
    const uint64_t hi_r = make_smem_desc(0, 16, static_cast<uint32_t>(8 + p.proc_ks - 1) * rowb, p.narrow ? SWZ_32B : SWZ_128B);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1, p.err, 22);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 128);
      if (p.rowhalo) {
        const int ks = p.proc_ks;
        for (int k = 0; k < nk; ++k) {
          mbar_wait(&full_bar[stage], phase, p.err, 23);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + static_cast<uint32_t>(a_off);
          for (int kx = 0; kx < ks; ++kx) {
            const uint32_t shift = static_cast<uint32_t>(p.proc_sign > 0 ? kx : ks - 1 - kx);
            const uint64_t da = hi_r | static_cast<uint64_t>(((sa + shift * rowb) & 0x3FFFF) >> 4);
            const uint64_t db = hi | static_cast<uint64_t>(((sb + static_cast<uint32_t>(kx) * b_bytes) & 0x3FFFF) >> 4);
            if (elect_one()) {
              if (!p.narrow) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma_f16(tmem_d, da + static_cast<uint64_t>(j * 2), db + static_cast<uint64_t>(j * 2), p.idesc, (k | kx | j) != 0);
              } else {
                umma_f16(tmem_d, da, db, p.idesc, (k | kx) != 0);
              }
            }
            __syncwarp();
          }
          if (elect_one()) {
            umma_commit(&empty_bar[stage]);
            if (k == nk - 1) umma_commit(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == kGrStages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      for (int k = 0; k < nk; ++k) {
        mbar_wait(&full_bar[stage], phase, p.err, 23);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kGcStageBytes);
        const uint32_t sb = sa + kGcStageA;
        const uint64_t da = hi | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
        const uint64_t db = hi | static_cast<uint64_t>((sb & 0x3FFFF) >> 4);
        if (elect_one()) {
          if (!p.narrow) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_f16(tmem_d, da + static_cast<uint64_t>(j * 2), db + static_cast<uint64_t>(j * 2), p.idesc, (k | j) != 0);
          } else {
            umma_f16(tmem_d, da, db, p.idesc, k != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (k == nk - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == kGcStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int nchunks = p.nt >> 4;
    float acc_s[8], acc_q[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { acc_s[c] = 0.f; acc_q[c] = 0.f; }
    int cur_ntile = -1;
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int tile = item % ntiles_pix;
      const int ntile = item / ntiles_pix;
      if (ntile != cur_ntile) {
        gc_flush_stats(p, cur_ntile, lane, acc_s, acc_q);
        cur_ntile = ntile;
      }
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int lx = (tile % p.tiles_x) * p.tw + (row % p.tw);
      const int ly = (tile / p.tiles_x) * p.th + (row / p.tw);
      const int fx = lx * p.ox_mul + p.ox_add;
      const int fy = ly * p.oy_mul + p.oy_add;
      const bool valid = (fx < p.out_w) && (fy < p.out_rows) && ((fy % p.out_period) < p.out_h);
      mbar_wait(&tfull_bar[as], aphase, p.err, 24);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 128);
      const long long obase = (static_cast<long long>(fy) * p.out_w + fx) * p.out_c + ntile * p.nt;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nchunks) {
          uint32_t v[16];
          tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + ntile * p.nt + c * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(bp + i);
              f[4 * i] += b4.x; f[4 * i + 1] += b4.y; f[4 * i + 2] += b4.z; f[4 * i + 3] += b4.w;
            }
          }
          if (valid) {
            if (p.addend != nullptr) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.addend + obase + c * 16);
              uint4 q[2] = {ap[0], ap[1]};
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(q);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[2 * i] += __low2float(h[i]);
                f[2 * i + 1] += __high2float(h[i]);
              }
            }
            if (p.mask != nullptr) {
              const uint4* mp = reinterpret_cast<const uint4*>(p.mask + obase + c * 16);
              uint4 q[2] = {mp[0], mp[1]};
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(q);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (!(__low2float(h[i]) > 0.f)) f[2 * i] *= p.slope;
                if (!(__high2float(h[i]) > 0.f)) f[2 * i + 1] *= p.slope;
              }
            } else if (p.act) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * p.slope;
            }
          }
          if (p.out_f32) {
            if (valid) {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + obase + c * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
          } else {
            uint32_t packed[8];
            if (p.out_f16) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                packed[i] = *reinterpret_cast<uint32_t*>(&h);
                f[2 * i] = __low2float(h);
                f[2 * i + 1] = __high2float(h);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                packed[i] = *reinterpret_cast<uint32_t*>(&h);
                f[2 * i] = __low2float(h);
                f[2 * i + 1] = __high2float(h);
              }
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<bf16_t*>(p.out) + obase + c * 16);
              dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
          }
          if (p.stats != nullptr) {
            // column sums over the warp's 32 pixels by recursive halving (as conv_gemm_kernel of dsr_conv.cu): after
            // the steps lane L (even) holds column ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)
            float s[16], q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { s[i] = valid ? f[i] : 0.f; q[i] = valid ? f[i] * f[i] : 0.f; }
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1) {
              const int d = w * 2;
              const bool hi = (lane & d) != 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (i < w) {
                  const float send_s = hi ? s[i] : s[i + w];
                  const float send_q = hi ? q[i] : q[i + w];
                  const float recv_s = __shfl_xor_sync(0xffffffffu, send_s, d);
                  const float recv_q = __shfl_xor_sync(0xffffffffu, send_q, d);
                  s[i] = (hi ? s[i + w] : s[i]) + recv_s;
                  q[i] = (hi ? q[i + w] : q[i]) + recv_q;
                }
              }
            }
            s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
            q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
            acc_s[c] += s[0];
            acc_q[c] += q[0];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    gc_flush_stats(p, cur_ntile, lane, acc_s, acc_q);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// =============================================================================================
// gwgrad_kernel
// =============================================================================================
__device__ __forceinline__ void gw_red_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kGwThreads, 1) gwgrad_kernel(const __grid_constant__ GWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGwStages * kGwStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGwStages;
  uint64_t* tfull_bar = bars + 2 * kGwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGwStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int r = blockIdx.x;
  const int split = r % p.nsplit; r /= p.nsplit;
  const int ci_t = r % p.ci_tiles; r /= p.ci_tiles;
  const int co_t = r % p.co_tiles;
  const int group = r / p.co_tiles;
  const int co0 = co_t * 128, ci0 = ci_t * 128;
  const int nco = (p.cout - co0) >= 128 ? 2 : 1;
  const int nci = (p.cin - ci0) >= 128 ? 2 : 1;
  const int npb = p.pb_x * p.pb_y;
  const int pb_begin = static_cast<int>((static_cast<long long>(npb) * split) / p.nsplit);
  const int pb_end = static_cast<int>((static_cast<long long>(npb) * (split + 1)) / p.nsplit);
  const int nkb = pb_end - pb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a);
    tma_prefetch_desc(&p.b);
    for (int s = 0; s < kGwStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (nco == 1) {
    // the second 64-channel column group of the dY operand is never loaded: keep it zero (rows 64..127 of D unused)
    for (int s = 0; s < kGwStages; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * kGwStageBytes + kGwChunk);
      for (int i = threadIdx.x; i < kGwChunk / 16; i += kGwThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        const int x0 = (pb % p.pb_x) * 8;
        const int y0 = (pb / p.pb_x) * 8;
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 31);
        uint8_t* sa = smem + stage * kGwStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>((nco + p.tpg * nci) * kGwChunk));
        for (int c = 0; c < nco; ++c) tma_load_5d(&p.a, &full_bar[stage], sa + c * kGwChunk, co0 + c * 64, 0, x0, 0, y0);
        for (int t = 0; t < p.tpg; ++t) {
          const GTap tp = p.taps[group * p.tpg + t];
          uint8_t* sb = sa + 2 * kGwChunk + t * 2 * kGwChunk;
          for (int c = 0; c < nci; ++c)
            tma_load_5d(&p.b, &full_bar[stage], sb + c * kGwChunk, ci0 + c * 64, tp.px, x0 + tp.dx, tp.py, y0 + tp.dy);
        }
        if (++stage == kGwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    // MN-major operands [pixels][channels]: 64-channel column groups kGwChunk apart (LBO), 8-pixel groups 1024 B (SBO)
    const uint64_t hi = make_smem_desc(0, kGwChunk, 1024, SWZ_128B);
    const uint32_t idesc = (nci == 2) ? p.idesc128 : p.idesc64;
    for (int k = 0; k < nkb; ++k) {
      mbar_wait(&full_bar[stage], phase, p.err, 32);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * kGwStageBytes);
      for (int t = 0; t < p.tpg; ++t) {
        const uint32_t sb = sa + 2 * kGwChunk + t * 2 * kGwChunk;
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(t * 128);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = hi | static_cast<uint64_t>(((sa + ks * 2048) & 0x3FFFF) >> 4);
            const uint64_t db = hi | static_cast<uint64_t>(((sb + ks * 2048) & 0x3FFFF) >> 4);
            umma_f16(tmem_d, da, db, idesc, (k | ks) != 0);
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(&empty_bar[stage]);
        if (k == nkb - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++stage == kGwStages) { stage = 0; phase ^= 1; }
    }
  } else if (nkb > 0) {
    const int quarter = warp & 3;
    const int co = co0 + quarter * 32 + lane;
    mbar_wait(tfull_bar, 0, p.err, 33);
    tc_fence_after();
    for (int t = 0; t < p.tpg; ++t) {
      const int w_tap = p.taps[group * p.tpg + t].b_row;
      float* drow = p.dw + (static_cast<long long>(w_tap) * p.cout + co) * p.cin + ci0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(t * 128);
      for (int c = 0; c < nci * 4; ++c) {
        uint32_t v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v);
        tmem_ld_wait();
        if (co < p.cout) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            gw_red_v4(drow + c * 16 + i * 4, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                      __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// host side
// =============================================================================================
static void pick_tile(int w, int* tw, int* th) {
  if (w % 16 == 0) { *tw = 16; *th = 8; }
  else { *tw = 8; *th = 16; }
}

static void s2_tap(int k, int8_t* par, int8_t* off) {   // input coordinate 2 o + k - 1 in the parity-split view
  if (k == 1) { *par = 0; *off = 0; }
  else { *par = 1; *off = static_cast<int8_t>(k == 0 ? -1 : 0); }
}

int make_gconv_fprop(GConvParams* g, const GT& in, const GT& out, const void* w_pack, int cin_pad, int cout_pad, int ks,
                     int stride, int* err) {
  memset(g, 0, sizeof(*g));
  if (in.C != cin_pad || out.C != cout_pad || in.B != out.B) return -41;
  if (stride == 2 && (ks != 3 || in.P != 2 * out.P || (in.W & 1) || in.W != 2 * out.W)) return -42;
  if (stride == 1 && (in.P != out.P || in.W != out.W)) return -43;
  g->narrow = (cin_pad == 16);
  if (!g->narrow && (cin_pad % 64)) return -44;
  g->kchunks = cin_pad / 64;
  g->nt = cout_pad >= 128 ? 128 : cout_pad;
  if (cout_pad % g->nt || (g->nt != 128 && g->nt != 64 && g->nt != 16)) return -45;
  g->ntiles_n = cout_pad / g->nt;
  pick_tile(out.W, &g->tw, &g->th);
  int rc;
  const int rowb = g->narrow ? 32 : 128;
  g->rowhalo = (stride == 1 && (out.W % 8) == 0 && !getenv("DSR_GANT_NO_ROWHALO")) ? 1 : 0;
  if (g->rowhalo) {
    g->tw = 8; g->th = 16;
    g->a_slot = (((8 + ks - 1) * 16 * rowb) + 1023) & ~1023;
    g->b_slot = ((ks * g->nt * rowb) + 1023) & ~1023;
    if (g->a_slot + g->b_slot > kGrMaxStage) g->rowhalo = 0;
  }
  if (g->rowhalo) {
    if ((rc = make_act_map(&g->a, in.ptr, 1, in.C, in.W, in.rows(), 1, g->narrow ? 16 : 64, 8 + ks - 1, 16))) return rc;
  } else
  if ((rc = make_act_map(&g->a, in.ptr, 1, in.C, in.W, in.rows(), stride, g->narrow ? 16 : 64, g->tw, g->th))) return rc;
  if ((rc = make_wgt_map(&g->b, w_pack, cin_pad, ks * ks * cout_pad, g->narrow ? 16 : 64, g->nt))) return rc;
  if (g->rowhalo && g->ntiles_n == 1) {
    g->b2_loads = (ks * g->nt + 255) / 256;
    g->b2_rows = ks * g->nt / g->b2_loads;
    if (g->b2_rows * g->b2_loads != ks * g->nt || (g->b2_rows & 7)) g->b2_loads = 0;
    else if ((rc = make_wgt_map(&g->b2, w_pack, cin_pad, ks * ks * cout_pad, g->narrow ? 16 : 64, g->b2_rows))) return rc;
  }
  g->ntaps = ks * ks;
  if (stride == 1) {
    g->proc_ks = ks; g->proc_pad = (ks - 1) / 2; g->proc_sign = 1; g->rows_per_tap = cout_pad;
  } else {
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        GTap t{};
        s2_tap(kx, &t.px, &t.dx);
        s2_tap(ky, &t.py, &t.dy);
        t.b_row = (ky * 3 + kx) * cout_pad;
        g->taps[ky * 3 + kx] = t;
      }
  }
  g->tiles_x = (out.W + g->tw - 1) / g->tw;
  g->tiles_y = (out.rows() + g->th - 1) / g->th;
  g->ox_mul = 1; g->oy_mul = 1;
  g->out_w = out.W; g->out_h = out.H; g->out_period = out.P; g->out_rows = out.rows();
  g->out_c = out.C;
  g->out = out.ptr;
  g->out_f32 = out.f32;
  g->out_f16 = out.f16;
  if (in.f16 != out.f16 && !out.f32) return -50;
  g->idesc = in.f16 ? make_idesc_f16(128, g->nt, FMT_F16, FMT_F16, 0, 0) : make_idesc_f16(128, g->nt, FMT_BF16, FMT_BF16, 0, 0);
  g->err = err;
  return 0;
}

// Data gradient: dx[ci] = sum over taps, co  dy(...)[co] * W[co][ci][tap]; w_pack_d is [tap][cin_pad rows][cout_pad].
// stride 2: one launch per output-parity class (nlaunch = 4).
int make_gconv_dgrad(GConvParams* g, int* nlaunch, const GT& dy, const GT& dx, const void* w_pack_d, int cin_pad,
                     int cout_pad, int ks, int stride, int* err) {
  if (dy.C != cout_pad || dx.C != cin_pad || dy.B != dx.B) return -46;
  if (stride == 2 && (ks != 3 || dx.P != 2 * dy.P || dx.W != 2 * dy.W)) return -47;
  if (stride == 1 && (dx.P != dy.P || dx.W != dy.W)) return -48;
  const int n = (stride == 1) ? 1 : 4;
  *nlaunch = n;
  for (int cls = 0; cls < n; ++cls) {
    GConvParams* q = g + cls;
    memset(q, 0, sizeof(*q));
    q->narrow = (cout_pad == 16);
    if (!q->narrow && (cout_pad % 64)) return -44;
    q->kchunks = cout_pad / 64;
    q->nt = cin_pad >= 128 ? 128 : cin_pad;
    if (cin_pad % q->nt || (q->nt != 128 && q->nt != 64 && q->nt != 16)) return -45;
    q->ntiles_n = cin_pad / q->nt;
    pick_tile(dy.W, &q->tw, &q->th);
    int rc;
    const int rowb = q->narrow ? 32 : 128;
    q->rowhalo = (stride == 1 && (dy.W % 8) == 0 && !getenv("DSR_GANT_NO_ROWHALO")) ? 1 : 0;
    if (q->rowhalo) {
      q->tw = 8; q->th = 16;
      q->a_slot = (((8 + ks - 1) * 16 * rowb) + 1023) & ~1023;
      q->b_slot = ((ks * q->nt * rowb) + 1023) & ~1023;
      if (q->a_slot + q->b_slot > kGrMaxStage) q->rowhalo = 0;
    }
    if (q->rowhalo) {
      if ((rc = make_act_map(&q->a, dy.ptr, 1, dy.C, dy.W, dy.rows(), 1, q->narrow ? 16 : 64, 8 + ks - 1, 16))) return rc;
    } else
    if ((rc = make_act_map(&q->a, dy.ptr, 1, dy.C, dy.W, dy.rows(), 1, q->narrow ? 16 : 64, q->tw, q->th))) return rc;
    if ((rc = make_wgt_map(&q->b, w_pack_d, cout_pad, ks * ks * cin_pad, q->narrow ? 16 : 64, q->nt))) return rc;
    if (q->rowhalo && q->ntiles_n == 1) {
      q->b2_loads = (ks * q->nt + 255) / 256;
      q->b2_rows = ks * q->nt / q->b2_loads;
      if (q->b2_rows * q->b2_loads != ks * q->nt || (q->b2_rows & 7)) q->b2_loads = 0;
      else if ((rc = make_wgt_map(&q->b2, w_pack_d, cout_pad, ks * ks * cin_pad, q->narrow ? 16 : 64, q->b2_rows))) return rc;
    }
    if (stride == 1) {
      q->ntaps = ks * ks;
      q->proc_ks = ks; q->proc_pad = (ks - 1) / 2; q->proc_sign = -1; q->rows_per_tap = cin_pad;
      q->ox_mul = 1; q->oy_mul = 1;
    } else {
      const int py = cls >> 1, px = cls & 1;
      int nt = 0;
      for (int ky = 0; ky < 3; ++ky) {
        if (((py + 1 - ky) & 1) != 0) continue;
        for (int kx = 0; kx < 3; ++kx) {
          if (((px + 1 - kx) & 1) != 0) continue;
          GTap t{};
          t.px = 0; t.py = 0;
          t.dy = static_cast<int8_t>((py + 1 - ky) / 2);
          t.dx = static_cast<int8_t>((px + 1 - kx) / 2);
          t.b_row = (ky * 3 + kx) * cin_pad;
          q->taps[nt++] = t;
        }
      }
      q->ntaps = nt;
      q->ox_mul = 2; q->ox_add = px; q->oy_mul = 2; q->oy_add = py;
    }
    q->tiles_x = (dy.W + q->tw - 1) / q->tw;
    q->tiles_y = (dy.rows() + q->th - 1) / q->th;
    q->out_w = dx.W; q->out_h = dx.H; q->out_period = dx.P; q->out_rows = dx.rows();
    q->out_c = dx.C;
    q->out = dx.ptr;
    q->out_f32 = dx.f32;
    q->idesc = make_idesc_f16(128, q->nt, FMT_BF16, FMT_BF16, 0, 0);
    q->err = err;
  }
  return 0;
}

int make_gwgrad(GWgradParams* g, const GT& dy, const GT& x, float* dw, int cin, int cout, int stride, int num_sms,
                int* err) {
  memset(g, 0, sizeof(*g));
  if (dy.C != cout || x.C != cin || (cin % 64) || (cout % 64)) return -49;
  if (stride == 2 && (x.P != 2 * dy.P || x.W != 2 * dy.W)) return -47;
  if (stride == 1 && (x.P != dy.P || x.W != dy.W)) return -48;
  int rc;
  if ((rc = make_act_map(&g->a, dy.ptr, 1, dy.C, dy.W, dy.rows(), 1, 64, 8, 8))) return rc;
  if ((rc = make_act_map(&g->b, x.ptr, 1, x.C, x.W, x.rows(), stride, 64, 8, 8))) return rc;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      GTap t{};
      if (stride == 1) {
        t.dx = static_cast<int8_t>(kx - 1); t.dy = static_cast<int8_t>(ky - 1);
      } else {
        s2_tap(kx, &t.px, &t.dx);
        s2_tap(ky, &t.py, &t.dy);
      }
      t.b_row = ky * 3 + kx;
      g->taps[ky * 3 + kx] = t;
    }
  g->ngroups = 3; g->tpg = 3;
  g->cout = cout; g->cin = cin;
  g->co_tiles = (cout + 127) / 128;
  g->ci_tiles = (cin + 127) / 128;
  g->pb_x = (dy.W + 7) / 8;
  g->pb_y = (dy.rows() + 7) / 8;
  const int npb = g->pb_x * g->pb_y;
  const int base = g->ngroups * g->co_tiles * g->ci_tiles;
  int nsplit = (2 * num_sms) / base;
  if (nsplit > npb / 4) nsplit = npb / 4;
  if (nsplit < 1) nsplit = 1;
  g->nsplit = nsplit;
  g->dw = dw;
  g->idesc64 = make_idesc_f16(128, 64, FMT_BF16, FMT_BF16, 1, 1);
  g->idesc128 = make_idesc_f16(128, 128, FMT_BF16, FMT_BF16, 1, 1);
  g->err = err;
  return 0;
}

// A stride-1 convolution-like GEMM over an explicit tap table (<= 9 taps, 64-channel chunks): out[p][n] = sum over taps,
// k of in[p + (dx, dy)][k] * w[tap.b_row + n][k].
int make_gconv_taps(GConvParams* g, const GT& in, const GT& out, const void* w, int k_pad, int n_pad, const GTap* taps,
                    int ntaps, int* err) {
  memset(g, 0, sizeof(*g));
  if (in.C != k_pad || out.C != n_pad || (k_pad % 64) || ntaps > 9 || in.P != out.P || in.W != out.W) return -41;
  g->kchunks = k_pad / 64;
  g->nt = n_pad >= 128 ? 128 : n_pad;
  if (n_pad % g->nt || (g->nt & 15)) return -45;
  g->ntiles_n = n_pad / g->nt;
  pick_tile(out.W, &g->tw, &g->th);
  int rc;
  if ((rc = make_act_map(&g->a, in.ptr, 1, in.C, in.W, in.rows(), 1, 64, g->tw, g->th))) return rc;
  int max_row = 0;
  for (int t = 0; t < ntaps; ++t) { g->taps[t] = taps[t]; if (taps[t].b_row > max_row) max_row = taps[t].b_row; }
  if ((rc = make_wgt_map(&g->b, w, k_pad, max_row + n_pad, 64, g->nt))) return rc;
  g->ntaps = ntaps;
  g->tiles_x = (out.W + g->tw - 1) / g->tw;
  g->tiles_y = (out.rows() + g->th - 1) / g->th;
  g->ox_mul = 1; g->oy_mul = 1;
  g->out_w = out.W; g->out_h = out.H; g->out_period = out.P; g->out_rows = out.rows();
  g->out_c = out.C;
  g->out = out.ptr;
  g->out_f32 = out.f32;
  g->idesc = make_idesc_f16(128, g->nt, FMT_BF16, FMT_BF16, 0, 0);
  g->err = err;
  return 0;
}

// Weight-gradient GEMM over an explicit tap table (ntaps a multiple of 3, <= 9): dw[tap.b_row][co][ci] += sum_p
// dy[p][co] * x[p + (dx, dy)][ci].
int make_gwgrad_taps(GWgradParams* g, const GT& dy, const GT& x, float* dw, int cin, int cout, const GTap* taps, int ntaps,
                     int num_sms, int* err) {
  int rc = make_gwgrad(g, dy, x, dw, cin, cout, 1, num_sms, err);
  if (rc) return rc;
  if (ntaps > 9 || (ntaps % 3)) return -49;
  for (int t = 0; t < ntaps; ++t) g->taps[t] = taps[t];
  g->ngroups = ntaps / 3;
  g->tpg = 3;
  return 0;
}

static int gant_set_attrs() {
  static bool done_dev[kMaxDevices] = {};
  bool& done = done_dev[device_slot()];
  if (done) return 0;
  cudaError_t e = cudaFuncSetAttribute(gconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGcSmemBytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(gwgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGwSmemBytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  done = true;
  return 0;
}

int launch_gconv(const GConvParams& p, int num_sms, cudaStream_t s) {
  int rc = gant_set_attrs();
  if (rc) return rc;
  const int nitems = p.tiles_x * p.tiles_y * p.ntiles_n;
  if (nitems <= 0) return 0;
  const int grid = nitems < num_sms ? nitems : num_sms;
  launch_k(gconv_kernel, dim3(grid), dim3(kGcThreads), kGcSmemBytes, s, p);
  return static_cast<int>(cudaGetLastError());
}

int launch_gwgrad(const GWgradParams& p, cudaStream_t s) {
  int rc = gant_set_attrs();
  if (rc) return rc;
  const int grid = p.ngroups * p.co_tiles * p.ci_tiles * p.nsplit;
  if (grid <= 0) return 0;
  launch_k(gwgrad_kernel, dim3(grid), dim3(kGwThreads), kGwSmemBytes, s, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace dsr
