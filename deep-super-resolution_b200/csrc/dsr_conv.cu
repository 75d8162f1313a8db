// dsr_conv.cu -- the two tensor-core kernels of the DIP step, hand-written for sm_100a:
//
//   conv_gemm_kernel : implicit-GEMM convolution (fprop of 3x3 s1 / 3x3 s2 / 1x1 convs, and their
//                      data-gradients), D[128 pixels][N] = sum over K-blocks A[128][kc] * B[N][kc]^T.
//                      TMA (5-D tiled maps over the padded NHWC tensor) -> 128B/32B-swizzled smem
//                      -> tcgen05.mma kind::f16 (fp32 accumulators in TMEM, double buffered)
//                      -> tcgen05.ld epilogue (16-bit store + fused BatchNorm sum / sum-of-squares).
//   wgrad_kernel     : weight gradient, D[128 co][ci] += dR^T X over pixel blocks, MN-major operands,
//                      split over pixel ranges and tap groups, fp32 vector reductions into HBM.
//
// Reference semantics being replaced: torch.nn.Conv2d forward/backward as instantiated by
// models/DIP/utils.py:83-105 (conv) and used at models/DIP/skip.py:54,60,64,79,85.
#include "dsr_conv.cuh"
#include "dsr_ptx.cuh"
#include "dsr_host.h"
#include "dsr_launch.cuh"

namespace dsr {

// In-kernel profiling stamps (bench.py roofline): the span from the moment the kernel's dependencies are satisfied
// (after griddepcontrol.wait) to its last CTA's exit, as it runs inside the pipelined iteration.
__device__ __forceinline__ unsigned long long prof_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void prof_begin(unsigned long long* prof) {
  if (prof != nullptr && threadIdx.x == 0) atomicMin(prof, prof_now());
}
__device__ __forceinline__ void prof_end(unsigned long long* prof) {
  if (prof != nullptr && threadIdx.x == 0) atomicMax(prof + 1, prof_now());
}

// =============================================================================================
// conv_gemm_kernel
// =============================================================================================
struct ConvItem { int tile, kb0, nkb, out_h, out_w; long long out_off; };
__device__ __forceinline__ ConvItem conv_item(const ConvGemmParams& p, int item) {
  ConvItem it;
  if (p.ncls == 0) {
    it.tile = item; it.kb0 = 0; it.nkb = p.nkb; it.out_h = p.out_h; it.out_w = p.out_w; it.out_off = 0;
    return it;
  }
  int c = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (k < p.ncls && item >= p.cls_tile0[k]) c = k;
  it.tile = item - p.cls_tile0[c];
  it.kb0 = p.cls_kb0[c];
  it.nkb = p.cls_nkb[c];
  it.out_h = p.cls_out_h[c];
  it.out_w = p.cls_out_w[c];
  it.out_off = p.cls_out_off[c];
  return it;
}

__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atoms (8 rows x 128 B).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kConvStages * kConvStageBytes);
  uint64_t* full_bar = bars;                       // [kConvStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kConvStages;        // [kConvStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kConvStages;    // [2]            MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kConvStages + 2;  // [2]         epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kConvStages + 4);
  float* stat_smem = reinterpret_cast<float*>(bars + 2 * kConvStages + 6);  // unused bytes after barriers
  (void)stat_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = p.ncls ? p.cls_tile0[p.ncls] : p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a64);
    tma_prefetch_desc(&p.a16);
    tma_prefetch_desc(&p.b64);
    tma_prefetch_desc(&p.b16);
    for (int s = 0; s < kConvStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], (blockDim.x >> 5) - 2);   // one arrive per epilogue warp (4 or 8)
    }
    fence_barrier_init();
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
  prof_begin(p.prof);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < ntiles; item += gridDim.x) {
        const ConvItem ci = conv_item(p, item);
        const int x0 = (ci.tile % p.tiles_x) * p.tw;
        const int y0 = (ci.tile / p.tiles_x) * p.th;
        for (int k = 0; k < ci.nkb; ++k) {
          const KBlk kb = p.kb[ci.kb0 + k];
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 1);
          uint8_t* sa = smem + stage * kConvStageBytes;
          uint8_t* sb = sa + kConvStageA;
          const uint32_t kbytes = kb.wide ? 128u : 32u;
          mbar_arrive_expect_tx(&full_bar[stage], (128u + static_cast<uint32_t>(p.n_mma)) * kbytes);
          tma_load_5d(kb.wide ? &p.a64 : &p.a16, &full_bar[stage], sa, kb.a_c, kb.a_px, x0 + kb.a_dx, kb.a_py,
                      y0 + kb.a_dy);
          tma_load_2d(kb.wide ? &p.b64 : &p.b16, &full_bar[stage], sb, kb.b_k, kb.b_row);
          if (++stage == kConvStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop (descriptors stay in uniform registers); one elected lane issues the tcgen05 instructions
    {
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      const uint64_t hi_w = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t hi_n = make_smem_desc(0, 16, 256, SWZ_32B);
      for (int item = blockIdx.x; item < ntiles; item += gridDim.x, ++t) {
        const ConvItem ci = conv_item(p, item);
        const int as = t & 1;
        const uint32_t aphase = (t >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1, p.err, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 256);
        for (int k = 0; k < ci.nkb; ++k) {
          const int wide = p.kb[ci.kb0 + k].wide;
          mbar_wait(&full_bar[stage], phase, p.err, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kConvStageBytes);
          const uint32_t sb = sa + kConvStageA;
          const uint64_t da = (wide ? hi_w : hi_n) | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
          const uint64_t db = (wide ? hi_w : hi_n) | static_cast<uint64_t>((sb & 0x3FFFF) >> 4);
          if (elect_one()) {
            if (wide) {
#pragma unroll
              for (int j = 0; j < 4; ++j)   // 4 x (K = 16 elements = 32 B) inside the 128 B swizzle row
                umma_f16(tmem_d, da + static_cast<uint64_t>(j * 2), db + static_cast<uint64_t>(j * 2), p.idesc,
                         (k | j) != 0);
            } else {
              umma_f16(tmem_d, da, db, p.idesc, k != 0);
            }
            umma_commit(&empty_bar[stage]);   // frees the smem stage once these MMAs have read it
            if (k == ci.nkb - 1) umma_commit(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == kConvStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> HBM (+ BN statistics) =====================
    // 8 epilogue warps: two per TMEM lane quarter, which take the even / the odd 16-column chunks (layers with a small
    // K -- the 32-channel first layer: 18 MMAs per tile -- are bound by this epilogue, not by the MMAs)
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;       // pixel index inside the tile
    const int nchunks = p.n_mma >> 4;
    const int cmask = (blockDim.x >> 5) > 6 ? 1 : 0, chalf = (warp - 2) >> 2;
    float acc_s[9], acc_q[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) { acc_s[c] = 0.f; acc_q[c] = 0.f; }
    int t = 0;
    for (int item = blockIdx.x; item < ntiles; item += gridDim.x, ++t) {
      const ConvItem ci = conv_item(p, item);
      const int as = t & 1;
      const uint32_t aphase = (t >> 1) & 1;
      const int x = (ci.tile % p.tiles_x) * p.tw + (row % p.tw);
      const int y = (ci.tile / p.tiles_x) * p.th + (row / p.tw);
      const bool valid = (x < ci.out_w) && (y < ci.out_h);
      mbar_wait(&tfull_bar[as], aphase, p.err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 256);
      const long long obase = ci.out_off + static_cast<long long>(y) * p.out_sy + static_cast<long long>(x) * p.out_sx;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        if (c < nchunks && (c & cmask) == (chalf & cmask)) {
          uint32_t v[16];
          tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v);
          tmem_ld_wait();
          uint32_t packed[8];
          float f[16];
          if (p.out_bf16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
              packed[i] = *reinterpret_cast<uint32_t*>(&h);
              f[2 * i] = __low2float(h);
              f[2 * i + 1] = __high2float(h);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
              packed[i] = *reinterpret_cast<uint32_t*>(&h);
              f[2 * i] = __low2float(h);
              f[2 * i + 1] = __high2float(h);
            }
          }
          if (valid && c * 16 < p.n_store) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + obase + c * 16);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            if (c * 16 + 8 < p.n_store) dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
          if (p.stats != nullptr) {
            // Column sums over the warp's 32 pixels by recursive halving: after the 5 steps lane L
            // holds the sum of column ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
            const float m = valid ? 1.f : 0.f;
            float s[16], q[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { s[i] = (m != 0.f) ? f[i] : 0.f; q[i] = (m != 0.f) ? f[i] * f[i] : 0.f; }
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1) {
              const int d = w * 2;                 // lane distance 16, 8, 4, 2
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
    if (p.stats != nullptr && (lane & 1) == 0) {
      const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        if (c < nchunks && (c & cmask) == (chalf & cmask)) {
          acc_add_f(&p.stats[c * 16 + col], acc_s[c]);
          acc_add_f(&p.stats[p.n_mma + c * 16 + col], acc_q[c]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  prof_end(p.prof);
  ks_end();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// conv_halo_kernel  (see dsr_conv.cuh)
// =============================================================================================
__device__ __forceinline__ void halo_epilogue_process(const HaloParams& p, const uint32_t (&v)[16], int c, int col0,
                                                      bool valid, long long obase, int lane, float& acc_s,
                                                      float& acc_q) {
  uint32_t packed[8];
  float f[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    packed[i] = *reinterpret_cast<uint32_t*>(&h);
    f[2 * i] = __low2float(h);
    f[2 * i + 1] = __high2float(h);
  }
  const int ch = col0 + c * 16;
  if (valid && ch < p.n_store) {
    uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + obase + ch;
    if (ch + 16 <= p.n_store) {          // one 32-byte store per lane (sm_100 256-bit vector store): one sector each
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(packed[0]),
                   "r"(packed[1]), "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]), "r"(packed[6]),
                   "r"(packed[7])
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
  if (p.stats != nullptr && !(p.dbg & 4)) {
    // column sums over the warp's 32 pixels by recursive halving (see conv_gemm_kernel)
    const float m = valid ? 1.f : 0.f;
    float s[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { s[i] = (m != 0.f) ? f[i] : 0.f; q[i] = (m != 0.f) ? f[i] * f[i] : 0.f; }
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
    acc_s += s[0];
    acc_q += q[0];
  }
}

__device__ __forceinline__ void halo_epilogue_chunk(const HaloParams& p, uint32_t taddr, int c, int col0, bool valid,
                                                    long long obase, int lane, float& acc_s, float& acc_q) {
  uint32_t v[16];
  tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v);
  tmem_ld_wait();
  halo_epilogue_process(p, v, c, col0, valid, obase, lane, acc_s, acc_q);
}

__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int wb = p.n_part * 128, nb = p.n_part * 32;                    // bytes of one resident weight block
  uint8_t* bres_w = smem;                                               // [9][n_wide] wide blocks
  uint8_t* bres_n = bres_w + p.ntaps * p.n_wide * wb;                   // [ntaps] narrow blocks
  uint8_t* a_wide = bres_n + ((p.ntaps * p.n_narrow * nb + 1023) & ~1023);   // ring of wide_slots x kHaloWideSlot
  uint8_t* a_narrow = a_wide + p.wide_slots * kHaloWideSlot;            // ring of 2 x kHaloNarrowSlot (if any)
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_narrow + (p.n_narrow ? 2 * kHaloNarrowSlot : 0));
  uint64_t* wfull = bars;                 // [3]
  uint64_t* wempty = bars + 3;            // [3]
  uint64_t* nfull = bars + 6;             // [2]
  uint64_t* nempty = bars + 8;            // [2]
  uint64_t* tfull = bars + 10;            // [kHaloAccStages]
  uint64_t* tempty = bars + 14;           // [kHaloAccStages]
  uint64_t* bres_bar = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int part = blockIdx.x % p.parts;
  const int seq0 = blockIdx.x / p.parts, seq_stride = gridDim.x / p.parts;
  const int ntiles = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a64);
    tma_prefetch_desc(&p.b64);
    if (p.n_narrow) { tma_prefetch_desc(&p.a16); tma_prefetch_desc(&p.b16); }
    for (int i = 0; i < 3; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&nfull[i], 1); mbar_init(&nempty[i], 1); }
    for (int i = 0; i < kHaloAccStages; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kHaloAccStages * 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {          // resident weights: packed >= 2 launches ago, fetched before the PDL wait
    mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(p.ntaps * (p.n_wide * wb + p.n_narrow * nb)));
    for (int t = 0; t < p.ntaps; ++t) {
      const int row = p.taps[t].b_row + part * p.n_part;
      for (int c = 0; c < p.n_wide; ++c)
        tma_load_2d(&p.b64, bres_bar, bres_w + (t * p.n_wide + c) * wb, c * 64, row);
      if (p.n_narrow) tma_load_2d(&p.b16, bres_bar, bres_n + t * nb, p.n_wide * 64, row);
    }
  }
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ws = 0, ns = 0;
      uint32_t wph = 0, nph = 0;
      for (int tile = seq0; tile < ntiles; tile += seq_stride) {
        const int x0 = (tile % p.tiles_x) * kHaloTW + p.org_x;
        const int y0 = (tile / p.tiles_x) * kHaloTH + p.org_y;
        for (int c = 0; c < p.n_wide; ++c) {
          mbar_wait(&wempty[ws], wph ^ 1, p.err, 21);
          if (p.dbg & 1) {
            mbar_arrive(&wfull[ws]);
          } else {
            mbar_arrive_expect_tx(&wfull[ws], p.halo_w * p.halo_h * 128);
            tma_load_5d(&p.a64, &wfull[ws], a_wide + ws * kHaloWideSlot, c * 64, 0, x0, 0, y0);
          }
          if (++ws == p.wide_slots) { ws = 0; wph ^= 1; }
        }
        if (p.n_narrow) {
          mbar_wait(&nempty[ns], nph ^ 1, p.err, 22);
          if (p.dbg & 1) {
            mbar_arrive(&nfull[ns]);
          } else {
            mbar_arrive_expect_tx(&nfull[ns], p.halo_w * p.halo_h * 32);
            tma_load_5d(&p.a16, &nfull[ns], a_narrow + ns * kHaloNarrowSlot, p.n_wide * 64, 0, x0, 0, y0);
          }
          if (++ns == 2) { ns = 0; nph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop with warp-uniform values (so that ptxas keeps descriptors in uniform
    // registers); only the tcgen05 instructions themselves are issued by one elected lane.
    {
      mbar_wait(bres_bar, 0, p.err, 23);
      int ws = 0, ns = 0, it = 0;
      uint32_t wph = 0, nph = 0;
      const uint32_t bw_addr = smem_u32(bres_w), bn_addr = smem_u32(bres_n);
      // descriptor = {hi word (constant), lo word = flags | (address >> 4)}; all per-tap / per-K-step changes are
      // plain adds on the lo word (addresses are < 256 KB, so they never carry into the LBO field)
      const uint64_t dw = make_smem_desc(0, 16, p.halo_w * 128, SWZ_128B);   // A, wide
      const uint64_t db_ = make_smem_desc(0, 16, 1024, SWZ_128B);                 // B, wide
      const uint64_t dan = make_smem_desc(0, 16, p.halo_w * 32, SWZ_32B);    // A, narrow
      const uint64_t dbn = make_smem_desc(0, 16, 256, SWZ_32B);                   // B, narrow
      const uint32_t aw_hi = static_cast<uint32_t>(dw >> 32), aw_lo0 = static_cast<uint32_t>(dw);
      const uint32_t bw_hi = static_cast<uint32_t>(db_ >> 32), bw_lo0 = static_cast<uint32_t>(db_);
      const uint32_t an_hi = static_cast<uint32_t>(dan >> 32), an_lo0 = static_cast<uint32_t>(dan);
      const uint32_t bn_hi = static_cast<uint32_t>(dbn >> 32), bn_lo0 = static_cast<uint32_t>(dbn);
      uint32_t tap_row[9];                                  // (oy * 10 + ox): halo row of the tap's view origin
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_row[t] = static_cast<uint32_t>(p.taps[t].oy * p.halo_w + p.taps[t].ox);
      const uint32_t b_tap_step = static_cast<uint32_t>(p.n_wide * wb) >> 4, bn_tap_step = static_cast<uint32_t>(nb) >> 4;
      const uint32_t idw = p.idesc_wide, idn = p.idesc_narrow;
      for (int tile = seq0; tile < ntiles; tile += seq_stride, ++it) {
        const int as = it % kHaloAccStages;
        const uint32_t aph = (it / kHaloAccStages) & 1;
        mbar_wait(&tempty[as], aph ^ 1, p.err, 24);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 64);
        uint32_t accum = 0;
        for (int c = 0; c < p.n_wide; ++c) {
          mbar_wait(&wfull[ws], wph, p.err, 25);
          tc_fence_after();
          const uint32_t a_lo = aw_lo0 | ((smem_u32(a_wide + ws * kHaloWideSlot) & 0x3FFFF) >> 4);
          const uint32_t b_lo = bw_lo0 | (((bw_addr + static_cast<uint32_t>(c * wb)) & 0x3FFFF) >> 4);
          if (elect_one()) {          // one elected region per chunk: 36 MMAs + the commit, descriptors by adds
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t >= p.ntaps) break;
              const uint32_t da_lo = a_lo + tap_row[t] * 8u;          // 128 B per halo row = 8 x 16 B
              const uint32_t db_lo = b_lo + static_cast<uint32_t>(t) * b_tap_step;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(tmem_d, (static_cast<uint64_t>(aw_hi) << 32) | (da_lo + 2u * j),
                         (static_cast<uint64_t>(bw_hi) << 32) | (db_lo + 2u * j), idw, (t | j) ? 1u : accum);
            }
            umma_commit(&wempty[ws]);
          }
          __syncwarp();
          accum = 1;
          if (++ws == p.wide_slots) { ws = 0; wph ^= 1; }
        }
        if (p.n_narrow) {
          mbar_wait(&nfull[ns], nph, p.err, 26);
          tc_fence_after();
          const uint32_t a_lo = an_lo0 | ((smem_u32(a_narrow + ns * kHaloNarrowSlot) & 0x3FFFF) >> 4);
          const uint32_t b_lo = bn_lo0 | ((bn_addr & 0x3FFFF) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t >= p.ntaps) break;
              const uint32_t da_lo = a_lo + tap_row[t] * 2u;          // 32 B per halo row
              const uint32_t db_lo = b_lo + static_cast<uint32_t>(t) * bn_tap_step;
              umma_f16(tmem_d, (static_cast<uint64_t>(an_hi) << 32) | da_lo, (static_cast<uint64_t>(bn_hi) << 32) | db_lo,
                       idn, 1u);
            }
            umma_commit(&nempty[ns]);
          }
          __syncwarp();
          if (++ns == 2) { ns = 0; nph ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull[as]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int nchunks = p.n_part >> 4;
    const int col0 = part * p.n_part;
    float acc_s[4] = {0.f, 0.f, 0.f, 0.f}, acc_q[4] = {0.f, 0.f, 0.f, 0.f};
    int it = 0;
    for (int tile = seq0; tile < ntiles; tile += seq_stride, ++it) {
      const int as = it % kHaloAccStages;
      const uint32_t aph = (it / kHaloAccStages) & 1;
      const int x = (tile % p.tiles_x) * kHaloTW + (row & (kHaloTW - 1));
      const int y = (tile / p.tiles_x) * kHaloTH + (row / kHaloTW);
      const bool valid = (x < p.out_w) && (y < p.out_h);
      mbar_wait(&tfull[as], aph, p.err, 27);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 64);
      const long long obase = static_cast<long long>(y) * p.out_sy + static_cast<long long>(x) * p.out_sx;
      if (!(p.dbg & 2)) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nchunks) halo_epilogue_chunk(p, taddr, c, col0, valid, obase, lane, acc_s[c], acc_q[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (p.stats != nullptr && (lane & 1) == 0) {
      const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunks) {
          acc_add_f(&p.stats[col0 + c * 16 + col], acc_s[c]);
          acc_add_f(&p.stats[p.stats_stride + col0 + c * 16 + col], acc_q[c]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kHaloAccStages * 64);
}

// =============================================================================================
// conv_halo2_kernel: the halo-tile kernel on CTA PAIRS (tcgen05 cta_group::2).  Each CTA of a pair owns one
// 8 x 16 pixel tile (its own halo ring) and keeps HALF of the weight rows resident; CTA 0 issues M = 256
// (both tiles) x N = all output channels MMAs, so every shared-memory operand byte feeds twice the math of the
// single-CTA kernel (SS-mode MMAs are operand-fetch bound: measured 76 clk per 128x64x16 MMA vs the 32 clk pipe
// floor).  Both CTAs stream A with pair-TMA loads accounted on CTA 0's barriers; MMA completion is multicast to
// both CTAs' barriers; both CTAs' epilogue warps release the accumulator stage on CTA 0's barrier.
// =============================================================================================
// Generator epilogue (EP = 1): out = [residual +] prelu(acc + bias), 16 channels of one pixel, fp16, one 32-byte store.
// The bias vector sits in shared memory (copied once per CTA) and the residual values were fetched before the
// accumulator wait, so nothing here waits on a global load.
__device__ __forceinline__ void halo_epilogue_fused(const HaloParams& p, const uint32_t (&v)[16], int c, bool valid,
                                                    long long obase, float slope, const float* sbias,
                                                    const uint4 (&res)[2]) {
  const int ch = c * 16;
  if (!valid || ch >= p.n_store) return;
  float f[16];
  const float4* b4 = reinterpret_cast<const float4*>(sbias + ch);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b = b4[i];
    f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b.x;
    f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b.y;
    f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b.z;
    f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b.w;
  }
  if (p.ep_slope != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = f[i] > 0.f ? f[i] : slope * f[i];
  }
  if (p.ep_res != nullptr) {
    const uint32_t rw[8] = {res[0].x, res[0].y, res[0].z, res[0].w, res[1].x, res[1].y, res[1].z, res[1].w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&rw[i]);
      f[2 * i] += __low2float(h);
      f[2 * i + 1] += __high2float(h);
    }
  }
  uint32_t packed[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    packed[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  uint16_t* dst = reinterpret_cast<uint16_t*>(p.out) + obase + ch;
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(packed[0]), "r"(packed[1]),
               "r"(packed[2]), "r"(packed[3]), "r"(packed[4]), "r"(packed[5]), "r"(packed[6]), "r"(packed[7])
               : "memory");
}

// EP: 0 = DIP step (raw fp16 output + BatchNorm sums), 1 = generator 3x3 / 1x1 (fused bias, PReLU, residual; tall
// batch grid), 2 = generator 9x9 output conv (81 shifted views of one 16 x 24 halo tile, tanh, fp32 NCHW planes).
#ifndef DSR_EARLY_RELEASE0
#define DSR_EARLY_RELEASE0 1
#endif
constexpr bool kEarlyRelease0 = DSR_EARLY_RELEASE0 != 0;   // DIP epilogue: drain the stage into registers first
constexpr bool kRelaxedRelease = true;    // accumulator hand-back without release semantics (see dsr_ptx.cuh)
template <int EP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kHalo2Threads, 1)
    conv_halo2_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int wb = p.n_part * 128, nb = p.n_part * 32;                    // bytes of one resident weight block
  const int slot_bytes = (EP == 0) ? kHaloWideSlot : p.wide_slot_bytes;
  uint8_t* bres_w = smem;
  uint8_t* bres_n = bres_w + p.ntaps * p.n_wide * wb;
  uint8_t* a_wide = bres_n + ((p.ntaps * p.n_narrow * nb + 1023) & ~1023);
  uint8_t* a_narrow = a_wide + p.wide_slots * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_narrow + (p.n_narrow ? 2 * kHaloNarrowSlot : 0));
  uint64_t* wfull = bars;                 // [8]  (CTA 0's copy is the live one); wide_slots <= 8 are used
  uint64_t* wempty = bars + 8;            // [8]  (each CTA waits on its own copy; multicast commit)
  uint64_t* nfull = bars + 16;            // [2]
  uint64_t* nempty = bars + 18;           // [2]
  uint64_t* tfull = bars + 20;            // [2]  (multicast commit)
  uint64_t* tempty = bars + 24;           // [2]  (CTA 0's copy: 8 arrivals = 4 epilogue warps x 2 CTAs)
  uint64_t* bres_bar = bars + 28;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // EP 1: [N] bias (<= 128 floats)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (EP == 1 && static_cast<int>(threadIdx.x) < 2 * p.n_part) sbias[threadIdx.x] = __ldg(p.ep_bias + threadIdx.x);   // load-time constant
  const int nsl = (EP == 0 && p.nsplit > 1) ? p.nsplit : 1;
  const int cluster_id = blockIdx.x >> 1;
  const int slice = cluster_id % nsl;                        // this cluster's slice of the output channels
  const int pair0 = cluster_id / nsl, pair_stride = (gridDim.x >> 1) / nsl;
  const int ch0 = slice * 2 * p.n_part;
  const int ntiles = p.tiles_x * p.tiles_y;
  const int npairs = (ntiles + 1) >> 1;
  constexpr int kAcc = 2;                 // accumulator stages, 256 TMEM columns apart

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a64);
    tma_prefetch_desc(&p.b64);
    if (p.n_narrow) { tma_prefetch_desc(&p.a16); tma_prefetch_desc(&p.b16); }
    for (int i = 0; i < 8; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&nfull[i], 1); mbar_init(&nempty[i], 1); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }   // 8 epilogue warps x 2 CTAs
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  cluster_sync_all();                     // barriers of both CTAs exist before anyone signals across the pair
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {          // resident weights: packed >= 2 launches ago, fetched before the PDL wait
    const uint32_t bbytes = static_cast<uint32_t>(p.ntaps * (p.n_wide * wb + p.n_narrow * nb));
    if (rank == 0) mbar_arrive_expect_tx(bres_bar, 2 * bbytes);
    for (int t = 0; t < p.ntaps; ++t) {
      const int row = (EP == 2 ? t * 2 * p.n_part : p.taps[t].b_row) + ch0 + static_cast<int>(rank) * p.n_part;
      for (int c = 0; c < p.n_wide; ++c)
        tma_load_2d_pair(&p.b64, bres_bar, bres_w + (t * p.n_wide + c) * wb, c * 64, row);
      if (p.n_narrow) tma_load_2d_pair(&p.b16, bres_bar, bres_n + t * nb, p.n_wide * 64, row);
    }
  }
  pdl_sync();
  prof_begin(p.prof);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int ws = 0, ns = 0;
      uint32_t wph = 0, nph = 0;
      for (int pair = pair0; pair < npairs; pair += pair_stride) {
        const int tile = 2 * pair + static_cast<int>(rank);      // may be == ntiles (odd count): coordinates still valid to fetch
        const int x0 = (tile % p.tiles_x) * kHaloTW + p.org_x;
        const int y0 = (tile / p.tiles_x) * kHaloTH + p.org_y;
        for (int c = 0; c < p.n_wide; ++c) {
          mbar_wait(&wempty[ws], wph ^ 1, p.err, 31);
          if (rank == 0) mbar_arrive_expect_tx(&wfull[ws], 2 * p.halo_w * p.halo_h * 128);
          tma_load_5d_pair(&p.a64, &wfull[ws], a_wide + ws * slot_bytes, c * 64, 0, x0, 0, y0);
          if (++ws == p.wide_slots) { ws = 0; wph ^= 1; }
        }
        if (p.n_narrow) {
          mbar_wait(&nempty[ns], nph ^ 1, p.err, 32);
          if (rank == 0) mbar_arrive_expect_tx(&nfull[ns], 2 * p.halo_w * p.halo_h * 32);
          tma_load_5d_pair(&p.a16, &nfull[ns], a_narrow + ns * kHaloNarrowSlot, p.n_wide * 64, 0, x0, 0, y0);
          if (++ns == 2) { ns = 0; nph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA 0 only; warp-uniform loop, one elected lane issues) =====================
    if (rank == 0) {
      mbar_wait(bres_bar, 0, p.err, 33);
      if (lane == 0) ks_mark_here(0);
      int ws = 0, ns = 0, it = 0;
      uint32_t wph = 0, nph = 0;
      const uint32_t bw_addr = smem_u32(bres_w), bn_addr = smem_u32(bres_n);
      const uint64_t dw = make_smem_desc(0, 16, p.halo_w * 128, SWZ_128B);
      const uint64_t db_ = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint64_t dan = make_smem_desc(0, 16, p.halo_w * 32, SWZ_32B);
      const uint64_t dbn = make_smem_desc(0, 16, 256, SWZ_32B);
      const uint32_t aw_hi = static_cast<uint32_t>(dw >> 32), aw_lo0 = static_cast<uint32_t>(dw);
      const uint32_t bw_hi = static_cast<uint32_t>(db_ >> 32), bw_lo0 = static_cast<uint32_t>(db_);
      const uint32_t an_hi = static_cast<uint32_t>(dan >> 32), an_lo0 = static_cast<uint32_t>(dan);
      const uint32_t bn_hi = static_cast<uint32_t>(dbn >> 32), bn_lo0 = static_cast<uint32_t>(dbn);
      uint32_t tap_row[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_row[t] = static_cast<uint32_t>(p.taps[t].oy * p.halo_w + p.taps[t].ox);
      const uint32_t b_tap_step = static_cast<uint32_t>(p.n_wide * wb) >> 4, bn_tap_step = static_cast<uint32_t>(nb) >> 4;
      const uint32_t idw = p.idesc_wide, idn = p.idesc_narrow;
      for (int pair = pair0; pair < npairs; pair += pair_stride, ++it) {
        const int as = it % kAcc;
        const uint32_t aph = (it / kAcc) & 1;
        mbar_wait(&tempty[as], aph ^ 1, p.err, 34);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 256);
        uint32_t accum = 0;
        for (int c = 0; c < p.n_wide; ++c) {
          mbar_wait(&wfull[ws], wph, p.err, 35);
          if (lane == 0 && it == 0 && c == 0) ks_mark_here(1);
          tc_fence_after();
          const uint32_t a_lo = aw_lo0 | ((smem_u32(a_wide + ws * slot_bytes) & 0x3FFFF) >> 4);
          const uint32_t b_lo = bw_lo0 | (((bw_addr + static_cast<uint32_t>(c * wb)) & 0x3FFFF) >> 4);
          if (elect_one()) {
            if (EP == 2) {                 // 9 x 9: tap (ky, kx) is the view shifted by ky * halo_w + kx pixel rows
              uint32_t db_lo = b_lo;
              for (int ky = 0; ky < 9; ++ky) {
                uint32_t da_lo = a_lo + static_cast<uint32_t>(ky * p.halo_w) * 8u;
#pragma unroll
                for (int kx = 0; kx < 9; ++kx, da_lo += 8u, db_lo += b_tap_step) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    umma_f16_pair(tmem_d, (static_cast<uint64_t>(aw_hi) << 32) | (da_lo + 2u * j),
                                  (static_cast<uint64_t>(bw_hi) << 32) | (db_lo + 2u * j), idw,
                                  (ky | kx | j) ? 1u : accum);
                }
              }
            } else {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                if (t >= p.ntaps) break;
                const uint32_t da_lo = a_lo + tap_row[t] * 8u;
                const uint32_t db_lo = b_lo + static_cast<uint32_t>(t) * b_tap_step;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma_f16_pair(tmem_d, (static_cast<uint64_t>(aw_hi) << 32) | (da_lo + 2u * j),
                                (static_cast<uint64_t>(bw_hi) << 32) | (db_lo + 2u * j), idw, (t | j) ? 1u : accum);
              }
            }
            umma_commit_pair(&wempty[ws]);
          }
          __syncwarp();
          accum = 1;
          if (++ws == p.wide_slots) { ws = 0; wph ^= 1; }
        }
        if (p.n_narrow) {
          mbar_wait(&nfull[ns], nph, p.err, 36);
          tc_fence_after();
          const uint32_t a_lo = an_lo0 | ((smem_u32(a_narrow + ns * kHaloNarrowSlot) & 0x3FFFF) >> 4);
          const uint32_t b_lo = bn_lo0 | ((bn_addr & 0x3FFFF) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; ++t)
              if (t < p.ntaps)
              umma_f16_pair(tmem_d, (static_cast<uint64_t>(an_hi) << 32) | (a_lo + tap_row[t] * 2u),
                            (static_cast<uint64_t>(bn_hi) << 32) | (b_lo + static_cast<uint32_t>(t) * bn_tap_step), idn,
                            1u);
            umma_commit_pair(&nempty[ns]);
          }
          __syncwarp();
          if (++ns == 2) { ns = 0; nph ^= 1; }
        }
        if (elect_one()) umma_commit_pair(&tfull[as]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs: own tile, all output channels) =====================
    // 8 warps: warp w reads TMEM lanes 32 (w & 3) .. +31 (the quarter its id allows) and one half of the columns,
    // so two warps per scheduler hide each other's TMEM-load / shuffle latencies; the next chunk's tcgen05.ld is
    // in flight while the current one is converted, stored and reduced.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int nchunks = (2 * p.n_part) >> 4;              // 8 (N = 128) or 9 (N = 144)
    const int c_begin = half ? (nchunks + 1) / 2 : 0, c_end = half ? nchunks : (nchunks + 1) / 2;
    float acc_s[5], acc_q[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) { acc_s[c] = 0.f; acc_q[c] = 0.f; }
    const float slope = (EP == 1 && p.ep_slope != nullptr) ? __ldg(p.ep_slope) : 0.f;
    int it = 0;
    for (int pair = pair0; pair < npairs; pair += pair_stride, ++it) {
      const int as = it % kAcc;
      const uint32_t aph = (it / kAcc) & 1;
      const int tile = 2 * pair + static_cast<int>(rank);
      const int x = (tile % p.tiles_x) * kHaloTW + (row & (kHaloTW - 1));
      int y = (tile / p.tiles_x) * kHaloTH + (row / kHaloTW);
      bool valid = (tile < ntiles) && (x < p.out_w) && (y < p.out_h);
      long long obase = 0;
      if (EP != 0) {                       // tall batch grid: rows between images are padding, not outputs
        const int img = y / p.img_rows;
        y -= img * p.img_rows;
        valid = valid && (y < p.img_h);
        obase = static_cast<long long>(img) * p.out_img_stride;
      }
      obase += static_cast<long long>(y) * p.out_sy + static_cast<long long>(x) * p.out_sx;
      uint4 res[4][2];
      if (EP == 1 && p.ep_res != nullptr && valid) {      // residual values: in flight while the MMAs of this tile run
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c_begin + i < c_end) {
            const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.ep_res) + obase +
                                                            (c_begin + i) * 16);
            res[i][0] = __ldg(r);
            res[i][1] = __ldg(r + 1);
          }
      }
      mbar_wait(&tfull[as], aph, p.err, 37);
      if (warp == 2 && lane == 0 && it == 0) ks_mark_here(2);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 256);
      if (EP == 1) {
        // all of this warp's accumulator columns (<= 64) go to registers in one TMEM round trip, the stage is handed
        // back to the MMA warp at once, and only then the bias / PReLU / residual / store work starts
        uint32_t v[4][16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c_begin + i < c_end) tmem_ld16(taddr + static_cast<uint32_t>((c_begin + i) * 16), v[i]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(&tempty[as], 0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c_begin + i < c_end) halo_epilogue_fused(p, v[i], c_begin + i, valid, obase, slope, sbias, res[i]);
        continue;
      } else if (EP == 2) {
        if (half == 0) {                   // N = 16: one chunk; the first n_store columns are the image planes
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tmem_ld_wait();
          if (valid) {
            float* o = reinterpret_cast<float*>(p.out) + obase;
#pragma unroll
            for (int n = 0; n < 4; ++n)
              if (n < p.n_store) o[n * p.ep_plane] = tanhf(__uint_as_float(v[n]) + __ldg(p.ep_bias + n));
          }
        }
      } else if (kEarlyRelease0 && !(p.dbg & 2)) {
        // DIP epilogue: the warp's <= 80 accumulator columns in one TMEM round trip, stage handed back at once
        uint32_t v[5][16];
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (c_begin + i < c_end) tmem_ld16(taddr + static_cast<uint32_t>((c_begin + i) * 16), v[i]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(&tempty[as], 0);
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (c_begin + i < c_end) halo_epilogue_process(p, v[i], c_begin + i, ch0, valid, obase, lane, acc_s[i], acc_q[i]);
        continue;
      } else if (!(p.dbg & 2)) {
        uint32_t v[2][16];
        tmem_ld16(taddr + static_cast<uint32_t>(c_begin * 16), v[0]);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int c = c_begin + i;
          if (c < c_end) {
            tmem_ld_wait();
            if (c + 1 < c_end) tmem_ld16(taddr + static_cast<uint32_t>((c + 1) * 16), v[(i + 1) & 1]);
            halo_epilogue_process(p, v[i & 1], c, ch0, valid, obase, lane, acc_s[i], acc_q[i]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kRelaxedRelease) mbar_arrive_cluster_relaxed(&tempty[as], 0);
        else mbar_arrive_cluster(&tempty[as], 0);
      }
    }
    if (p.stats != nullptr && (lane & 1) == 0) {
      const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c = c_begin + i;
        if (c < c_end) {
          acc_add_f(&p.stats[ch0 + c * 16 + col], acc_s[i]);
          acc_add_f(&p.stats[p.stats_stride + ch0 + c * 16 + col], acc_q[i]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ks_mark_here(3);
  cluster_sync_all();                     // the peer may still be reading our shared memory / signalling our barriers
  prof_end(p.prof);
  ks_end();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// =============================================================================================
// conv9_out_kernel: the generator's 9 x 9, 64 -> 3 output convolution + tanh (generator.py:63,80-82) on CTA pairs.
//
// With N = 3 real output channels a tap-by-tap implicit GEMM (conv_halo2_kernel<2>) issues 81 x 4 MMAs of N = 16 per
// 256 pixels and is bound by the A-operand fetch of each MMA (measured 43 clk / MMA, 83 TFLOP/s).  Here the kx taps
// are folded into N:   S[p'][kx * 3 + co] = sum_ky sum_ci X[p' + (ky, 0)][ci] * W[ky][kx][co][ci]      (N = 27 -> 32)
// is accumulated over the 9 ky taps (row-shifted views of one halo tile) in TMEM -- 9 x 4 MMAs -- and the epilogue
// finishes  out[y][x][co] = sum_kx S[(y, x + kx)][kx * 3 + co]  through shared memory (an x-shift within one row).
// Geometry per CTA: halo tile 32 x 16 pixels x 64 ch (one TMA box, 64 KB, rows contiguous so a 128-row MMA operand
// is 4 whole halo rows: SBO = 1024); two M = 128 units (4 rows x 32 S-positions each) give 24 x 8 output pixels.
// 72 MMAs (pair: M = 256, N = 32, K = 16) per 384 output pixels instead of 486.
// =============================================================================================
constexpr int kC9Threads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kC9OutW = 24, kC9OutH = 8;   // output pixels per CTA tile
constexpr int kC9HaloW = 32, kC9HaloH = 16;
constexpr int kC9Slot = kC9HaloW * kC9HaloH * 128;          // 65536
constexpr int kC9Slots = 2;
constexpr int kC9WBlock = 16 * 128;                         // one ky block of this CTA's 16 weight rows
constexpr int kC9SStride = 27;                              // floats per S row in the staging buffer (odd: no bank conflicts)
constexpr int kC9Stage = 128 * kC9SStride * 4;              // 13824
constexpr int kC9Smem = 9 * kC9WBlock + kC9Slots * kC9Slot + 2 * kC9Stage + 256 + 1024 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC9Threads, 1)
    conv9_out_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_slots = smem;                                  // [kC9Slots][64 KB]
  uint8_t* bres = a_slots + kC9Slots * kC9Slot;             // [9][16 rows][128 B]
  float* stage = reinterpret_cast<float*>(bres + ((9 * kC9WBlock + 1023) & ~1023));   // [2][128][27]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage) + 2 * kC9Stage);
  uint64_t* wfull = bars;                 // [2]  (CTA 0's copy is the live one)
  uint64_t* wempty = bars + 2;            // [2]
  uint64_t* tfull = bars + 4;             // [2]
  uint64_t* tempty = bars + 6;            // [2]  (CTA 0's copy: 4 epilogue warps x 2 CTAs)
  uint64_t* bres_bar = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
  const int ntiles = p.tiles_x * p.tiles_y;
  const int npairs = (ntiles + 1) >> 1;
  constexpr int kAcc = 2;                 // accumulator stages, 64 TMEM columns each (2 units x N = 32)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a64);
    tma_prefetch_desc(&p.b64);
    for (int i = 0; i < kC9Slots; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 128);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0 && lane == 0) {          // resident weights (packed at load time), fetched before the PDL wait
    if (rank == 0) mbar_arrive_expect_tx(bres_bar, 2 * 9 * kC9WBlock);
    for (int ky = 0; ky < 9; ++ky)
      tma_load_2d_pair(&p.b64, bres_bar, bres + ky * kC9WBlock, 0, ky * 32 + static_cast<int>(rank) * 16);
  }
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      for (int pair = pair0; pair < npairs; pair += pair_stride) {
        const int tile = 2 * pair + static_cast<int>(rank);       // == ntiles for an odd count: coordinates still valid
        const int x0 = (tile % p.tiles_x) * kC9OutW - 4;
        const int y0 = (tile / p.tiles_x) * kC9OutH - 4;
        mbar_wait(&wempty[ws], wph ^ 1, p.err, 41);
        if (rank == 0) mbar_arrive_expect_tx(&wfull[ws], 2 * kC9Slot);
        tma_load_5d_pair(&p.a64, &wfull[ws], a_slots + ws * kC9Slot, 0, 0, x0, 0, y0);
        if (++ws == kC9Slots) { ws = 0; wph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (CTA 0 only) =====================
    if (rank == 0) {
      mbar_wait(bres_bar, 0, p.err, 43);
      int ws = 0, it = 0;
      uint32_t wph = 0;
      const uint64_t dsc = make_smem_desc(0, 16, 1024, SWZ_128B);
      const uint32_t d_hi = static_cast<uint32_t>(dsc >> 32), d_lo0 = static_cast<uint32_t>(dsc);
      const uint32_t b_lo = d_lo0 | ((smem_u32(bres) & 0x3FFFF) >> 4);
      const uint32_t idesc = p.idesc_wide;
      for (int pair = pair0; pair < npairs; pair += pair_stride, ++it) {
        const int as = it % kAcc;
        const uint32_t aph = (it / kAcc) & 1;
        mbar_wait(&tempty[as], aph ^ 1, p.err, 44);
        mbar_wait(&wfull[ws], wph, p.err, 45);
        tc_fence_after();
        const uint32_t a_lo = d_lo0 | ((smem_u32(a_slots + ws * kC9Slot) & 0x3FFFF) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 64 + u * 32);
#pragma unroll
            for (int ky = 0; ky < 9; ++ky) {
              // rows (4 u + ky) * 32 .. + 127 of the halo tile: 128 B per row, descriptor units of 16 B
              const uint32_t da_lo = a_lo + static_cast<uint32_t>((4 * u + ky) * 32 * 8);
              const uint32_t db_lo = b_lo + static_cast<uint32_t>(ky * (kC9WBlock >> 4));
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16_pair(tmem_d, (static_cast<uint64_t>(d_hi) << 32) | (da_lo + 2u * j),
                              (static_cast<uint64_t>(d_hi) << 32) | (db_lo + 2u * j), idesc, (ky | j) ? 1u : 0u);
            }
          }
          umma_commit_pair(&wempty[ws]);
          umma_commit_pair(&tfull[as]);
        }
        __syncwarp();
        if (++ws == kC9Slots) { ws = 0; wph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs: 4 warps = 128 S-positions per unit) =====================
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;            // TMEM lane = S-position: row m / 32 (0..3), x' = m % 32
    const int et = threadIdx.x - 64;              // 0..127 among the epilogue warps
    const float b0 = __ldg(p.ep_bias), b1 = __ldg(p.ep_bias + 1), b2 = __ldg(p.ep_bias + 2);
    int it = 0;
    for (int pair = pair0; pair < npairs; pair += pair_stride, ++it) {
      const int as = it % kAcc;
      const uint32_t aph = (it / kAcc) & 1;
      const int tile = 2 * pair + static_cast<int>(rank);
      const int tx0 = (tile % p.tiles_x) * kC9OutW, ty0 = (tile / p.tiles_x) * kC9OutH;
      mbar_wait(&tfull[as], aph, p.err, 47);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 64 + u * 32);
        uint32_t v0[16], v1[16];
        tmem_ld16(taddr, v0);
        tmem_ld16(taddr + 16, v1);
        tmem_ld_wait();
        float* srow = stage + u * (128 * kC9SStride) + m * kC9SStride;
#pragma unroll
        for (int i = 0; i < 16; ++i) srow[i] = __uint_as_float(v0[i]);
#pragma unroll
        for (int i = 0; i < 11; ++i) srow[16 + i] = __uint_as_float(v1[i]);
        if (u == 1) {                              // both units are out of TMEM: release the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // 96 output pixels of this unit (4 rows x 24), one thread each
        if (et < 4 * kC9OutW) {
          const int yy = et / kC9OutW, xo = et % kC9OutW;
          const float* sp = stage + u * (128 * kC9SStride) + (yy * 32 + xo) * kC9SStride;
          float o0 = b0, o1 = b1, o2 = b2;
#pragma unroll
          for (int kx = 0; kx < 9; ++kx) {
            o0 += sp[kx * (kC9SStride + 3)];
            o1 += sp[kx * (kC9SStride + 3) + 1];
            o2 += sp[kx * (kC9SStride + 3) + 2];
          }
          int y = ty0 + 4 * u + yy;
          const int x = tx0 + xo;
          bool valid = (tile < ntiles) && (x < p.out_w) && (y < p.out_h);
          const int img = y / p.img_rows;
          y -= img * p.img_rows;
          valid = valid && (y < p.img_h);
          if (valid) {
            float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(img) * p.out_img_stride +
                       static_cast<long long>(y) * p.out_sy + x;
            o[0] = tanhf(o0);
            o[p.ep_plane] = tanhf(o1);
            o[2 * p.ep_plane] = tanhf(o2);
          }
        }
      }
      // staging buffer u is rewritten one tile later, after the next bar.sync of the other unit: every thread has
      // finished reading it by then (its reads precede its arrival at that barrier)
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 128);
}

// =============================================================================================
// wgrad_kernel
// =============================================================================================
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// split-K epilogue: fp32 atomics into the gradient (default), or -- deterministic mode -- a plain store into this
// split's private partial, summed in split order by wgrad_reduce_kernel
__device__ __forceinline__ void wg_emit_v4(float* addr, bool plain, float a, float b, float c, float d) {
  if (plain) *reinterpret_cast<float4*>(addr) = make_float4(a, b, c, d);
  else red_add_v4(addr, a, b, c, d);
}

// dw[i] = sum over splits (in split order) of part[s * stride + i]: the fixed-order second stage of the split-K
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, long long stride, int nsplit, float* __restrict__ dw,
                                    long long n4) {
  pdl_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < nsplit; ++sp) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + sp * stride) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dw)[i] = acc;
  }
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgStages;
  uint64_t* tfull_bar = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x / p.nsplit;
  const int split = blockIdx.x % p.nsplit;
  const int npb = p.pb_x * p.pb_y;
  const int pb_begin = static_cast<int>((static_cast<long long>(npb) * split) / p.nsplit);
  const int pb_end = static_cast<int>((static_cast<long long>(npb) * (split + 1)) / p.nsplit);
  const int nkb = pb_end - pb_begin;
  const int ncols = p.n64 * 64 + p.n16 * 16;       // ci handled (<= 144)
  const uint32_t tapB = static_cast<uint32_t>(p.n64 * kWgPix * 128 + p.n16 * kWgPix * 32);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a64);
    tma_prefetch_desc(&p.b64);
    tma_prefetch_desc(&p.b16);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
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
  prof_begin(p.prof);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        const int x0 = (pb % p.pb_x) * p.pw;
        const int y0 = (pb / p.pb_x) * p.ph;
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 11);
        uint8_t* sa = smem + stage * kWgStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(kWgStageA) + static_cast<uint32_t>(p.ntaps) * tapB);
        // dR lives on the padded output grid: interior pixel (y, x) is at (y + 1, x + 1).
        tma_load_5d(&p.a64, &full_bar[stage], sa, 0, 0, x0 + 1, 0, y0 + 1);
        tma_load_5d(&p.a64, &full_bar[stage], sa + kWgPix * 128, 64, 0, x0 + 1, 0, y0 + 1);
        for (int t = 0; t < p.ntaps; ++t) {
          const WgTap tp = p.taps[group][t];
          uint8_t* sb = sa + kWgStageA + t * tapB;
          for (int c = 0; c < p.n64; ++c)
            tma_load_5d(&p.b64, &full_bar[stage], sb + c * (kWgPix * 128), c * 64, tp.px, x0 + tp.dx, tp.py, y0 + tp.dy);
          for (int c = 0; c < p.n16; ++c)
            tma_load_5d(&p.b16, &full_bar[stage], sb + p.n64 * (kWgPix * 128) + c * (kWgPix * 32), p.c16_base + c * 16,
                        tp.px, x0 + tp.dx, tp.py, y0 + tp.dy);
        }
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop; one elected lane issues the tcgen05 instructions
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t hi_a = make_smem_desc(0, kWgPix * 128, 1024, SWZ_128B);
      const uint64_t hi_b16 = make_smem_desc(0, kWgPix * 32, 256, SWZ_32B);
      for (int k = 0; k < nkb; ++k) {
        mbar_wait(&full_bar[stage], phase, p.err, 12);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kWgStageBytes);
        for (int t = 0; t < p.ntaps; ++t) {
          const uint32_t sb = sa + kWgStageA + t * tapB;
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(t * kWgTapCols);
          const uint32_t sb16 = sb + p.n64 * (kWgPix * 128);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < kWgPix / 16; ++ks) {
              // A: [pixels][co] MN-major, 2 column groups of 64 co (LBO), 8-pixel row groups (SBO).
              const uint64_t da = hi_a | static_cast<uint64_t>(((sa + ks * 2048) & 0x3FFFF) >> 4);
              if (p.n64) {
                const uint64_t db = hi_a | static_cast<uint64_t>(((sb + ks * 2048) & 0x3FFFF) >> 4);
                umma_f16(tmem_d, da, db, p.idesc64, (k | ks) != 0);
              }
              if (p.n16) {
                const uint64_t db = hi_b16 | static_cast<uint64_t>(((sb16 + ks * 512) & 0x3FFFF) >> 4);
                umma_f16(tmem_d + static_cast<uint32_t>(p.n64 * 64), da, db, p.idesc16, (k | ks) != 0);
              }
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(&empty_bar[stage]);
          if (k == nkb - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (nkb > 0) {
    const int quarter = warp & 3;
    const int co = quarter * 32 + lane;
    mbar_wait(tfull_bar, 0, p.err, 13);
    tc_fence_after();
    for (int t = 0; t < p.ntaps; ++t) {
      const int w_tap = p.taps[group][t].w_tap;
      float* drow = (p.part ? p.part + split * p.part_stride : p.dw) + (static_cast<long long>(w_tap) * 128 + co) * p.ldw;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(t * kWgTapCols);
      for (int c = 0; c * 16 < ncols; ++c) {
        uint32_t v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          wg_emit_v4(drow + c * 16 + i * 4, p.part != nullptr, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                     __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  prof_end(p.prof);
  ks_end();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// wgrad_halo_kernel  (see dsr_conv.cuh)
// =============================================================================================
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgHStages * kWgHStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgHStages;
  uint64_t* tfull_bar = bars + 2 * kWgHStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgHStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cl3 = p.cl3;
  const int group = cl3 ? static_cast<int>(cluster_ctarank()) : static_cast<int>(blockIdx.x) / p.nsplit;
  const int split = cl3 ? static_cast<int>(blockIdx.x) / 3 : static_cast<int>(blockIdx.x) % p.nsplit;
  const int npb = p.pb_x * p.pb_y;
  const int pb_begin = static_cast<int>((static_cast<long long>(npb) * split) / p.nsplit);
  const int pb_end = static_cast<int>((static_cast<long long>(npb) * (split + 1)) / p.nsplit);
  const int nkb = pb_end - pb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a64);
    tma_prefetch_desc(&p.b64[cl3 ? 1 : 0]);
    for (int s = 0; s < kWgHStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], cl3 ? 3 : 1);       // cl3: a stage is rewritten in all three CTAs at once
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (cl3) cluster_sync_all();                      // every CTA's barriers exist before anyone multicasts into them
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();
  prof_begin(p.prof);

  if (warp == 0 && cl3) {
    if (lane == 0) {
      // every CTA of the cluster receives the whole stage; this CTA issues its third of the (multicast) loads
      const uint32_t stage_bytes = kWgHStageA + static_cast<uint32_t>(kWgC3Rows * (p.n64 * 128 + p.n16 * 32));
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        const int x0 = (pb % p.pb_x) * 8;
        const int y0 = (pb / p.pb_x) * 8;
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 44);
        uint8_t* sa = smem + stage * kWgHStageBytes;
        uint8_t* sb = sa + kWgHStageA;
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        if (group == 0) {
          tma_load_5d_mc(&p.a64, &full_bar[stage], sa, 0, 0, x0 + 1, 0, y0 + 1, 7);
          tma_load_5d_mc(&p.a64, &full_bar[stage], sa + 64 * 128, 64, 0, x0 + 1, 0, y0 + 1, 7);
        } else if (group == 1) {
          tma_load_5d_mc(&p.b64[1], &full_bar[stage], sb, 0, 0, x0, 0, y0, 7);
          if (p.n16)
            tma_load_5d_mc(&p.b16[1], &full_bar[stage], sb + 2 * kWgC3Chunk, p.c16_base, 0, x0, 0, y0, 7);
        } else {
          tma_load_5d_mc(&p.b64[1], &full_bar[stage], sb + kWgC3Chunk, 64, 0, x0, 0, y0, 7);
        }
        if (++stage == kWgHStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      uint32_t stage_bytes = kWgHStageA;
      for (int b = 0; b < p.nbox; ++b)
        stage_bytes += static_cast<uint32_t>(p.box[group][b].width * 8 * (p.n64 * 128 + p.n16 * 32));
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        const int x0 = (pb % p.pb_x) * 8;
        const int y0 = (pb / p.pb_x) * 8;
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 41);
        uint8_t* sa = smem + stage * kWgHStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        tma_load_5d(&p.a64, &full_bar[stage], sa, 0, 0, x0 + 1, 0, y0 + 1);
        tma_load_5d(&p.a64, &full_bar[stage], sa + 64 * 128, 64, 0, x0 + 1, 0, y0 + 1);
        for (int b = 0; b < p.nbox; ++b) {
          const WgBox bx = p.box[group][b];
          uint8_t* sb = sa + kWgHStageA + bx.off16 * 16;
          const int rows = bx.width * 8;
          for (int c = 0; c < p.n64; ++c)
            tma_load_5d(&p.b64[b], &full_bar[stage], sb + c * rows * 128, c * 64, bx.px, x0 + bx.dx, bx.py, y0 + bx.dy);
          for (int c = 0; c < p.n16; ++c)
            tma_load_5d(&p.b16[b], &full_bar[stage], sb + p.n64 * rows * 128 + c * rows * 32, p.c16_base + c * 16, bx.px,
                        x0 + bx.dx, bx.py, y0 + bx.dy);
        }
        if (++stage == kWgHStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop; one elected lane issues the tcgen05 instructions
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t hi_a = make_smem_desc(0, 64 * 128, 1024, SWZ_128B);     // dR: chunk stride 8 KB, 8-pixel groups 1 KB
    for (int k = 0; k < nkb; ++k) {
      mbar_wait(&full_bar[stage], phase, p.err, 42);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * kWgHStageBytes);
      for (int r = 0; r < p.nruns; ++r) {
        const WgRun rn = p.runs[group][r];
        const WgBox bx = p.box[group][rn.box];
        const uint32_t rows = static_cast<uint32_t>(bx.width * 8);
        const uint32_t cstride = cl3 ? static_cast<uint32_t>(kWgC3Chunk) : rows * 128;     // between 64-channel chunks
        // cl3: one 10 x 10 halo for all tap rows, this CTA's views start `group` halo rows down
        const uint32_t row0 = cl3 ? static_cast<uint32_t>(group * bx.width) : 0u;
        const uint32_t sb = sa + kWgHStageA + static_cast<uint32_t>(bx.off16 * 16) + row0 * 128;
        const uint32_t sb16 = sa + kWgHStageA + static_cast<uint32_t>(bx.off16 * 16) + static_cast<uint32_t>(p.n64) * cstride +
                              row0 * 32;
        // wide: N-chunks = taps (LBO = 1 pixel row); 8-pixel groups one box row apart
        const uint64_t hi_b = make_smem_desc(0, 128, static_cast<uint32_t>(bx.width) * 128, SWZ_128B);
        const uint64_t hi_b16 = p.merge_narrow ? make_smem_desc(0, 32, static_cast<uint32_t>(bx.width) * 32, SWZ_32B)
                                               : make_smem_desc(0, rows * 32, static_cast<uint32_t>(bx.width) * 32, SWZ_32B);
        const uint32_t b0 = sb + static_cast<uint32_t>(rn.shift0) * 128, b16 = sb16 + static_cast<uint32_t>(rn.shift0) * 32;
        const uint32_t kstep = 2u * static_cast<uint32_t>(bx.width);       // K = 16 pixels = 2 rows of the box
        const uint32_t id_w = p.idesc_base | ((static_cast<uint32_t>(rn.r) * 64u >> 3) << 17);
        const uint32_t id_n = p.idesc_base | ((static_cast<uint32_t>(p.merge_narrow ? rn.r : p.n16) * 16u >> 3) << 17);
        const uint32_t cw = tmem_base + static_cast<uint32_t>(rn.col_wide), cn = tmem_base + static_cast<uint32_t>(rn.col_narrow);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = hi_a | static_cast<uint64_t>(((sa + ks * 2048) & 0x3FFFF) >> 4);
            for (int c = 0; c < p.n64; ++c) {
              const uint64_t db = hi_b | static_cast<uint64_t>(((b0 + c * cstride + ks * kstep * 128) & 0x3FFFF) >> 4);
              umma_f16(cw + static_cast<uint32_t>(c * rn.r * 64), da, db, id_w, (k | ks) != 0);
            }
            if (p.n16) {
              const uint64_t db = hi_b16 | static_cast<uint64_t>(((b16 + ks * kstep * 32) & 0x3FFFF) >> 4);
              umma_f16(cn, da, db, id_n, (k | ks) != 0);
            }
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (cl3) umma_commit_mc(&empty_bar[stage], 7);
        else umma_commit(&empty_bar[stage]);
        if (k == nkb - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++stage == kWgHStages) { stage = 0; phase ^= 1; }
    }
  } else if (nkb > 0) {
    const int quarter = warp & 3;
    const int co = quarter * 32 + lane;
    mbar_wait(tfull_bar, 0, p.err, 43);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int j = 0; j < p.ncolchunks; ++j) {
      const WgCol cc = p.cols[group][j];
      float* drow = (p.part ? p.part + split * p.part_stride : p.dw) + (static_cast<long long>(cc.w_tap) * 128 + co) * p.ldw + cc.ci0;
      uint32_t v[16];
      tmem_ld16(taddr + static_cast<uint32_t>(j * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        wg_emit_v4(drow + i * 4, p.part != nullptr, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cl3) cluster_sync_all();            // peers may still multicast into this CTA's barriers until their last commit
  prof_end(p.prof);
  ks_end();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// Host side: tensor-map construction and launches
// =============================================================================================
static PFN_encodeTiled g_encode = nullptr;

int ensure_driver_api() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) return -100;
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return 0;
}

// 5-D map over a padded NHWC tensor.  `base` points at padded pixel (0,0), channel 0.
//   C: channels per pixel (pitch), Wp/Hp: padded extents, step: 1 or 2 (parity split),
//   box_c: 64 (128B swizzle) or 16 (32B swizzle); box_w x box_h pixels.
int make_act_map(CUtensorMap* m, const void* base, int elem_is_16bit, int C, int Wp, int Hp, int step, int box_c,
                 int box_w, int box_h) {
  (void)elem_is_16bit;
  if (ensure_driver_api()) return -100;
  const cuuint64_t es = 2;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  dims[0] = static_cast<cuuint64_t>(C);
  dims[1] = static_cast<cuuint64_t>(step);
  dims[2] = static_cast<cuuint64_t>((Wp + step - 1) / step);
  dims[3] = static_cast<cuuint64_t>(step);
  dims[4] = static_cast<cuuint64_t>((Hp + step - 1) / step);
  strides[0] = static_cast<cuuint64_t>(C) * es;                 // px
  strides[1] = static_cast<cuuint64_t>(C) * es * step;          // x
  strides[2] = static_cast<cuuint64_t>(C) * es * Wp;            // py
  strides[3] = static_cast<cuuint64_t>(C) * es * Wp * step;     // y
  box[0] = static_cast<cuuint32_t>(box_c);
  box[1] = 1;
  box[2] = static_cast<cuuint32_t>(box_w);
  box[3] = 1;
  box[4] = static_cast<cuuint32_t>(box_h);
  const CUtensorMapSwizzle sw = (box_c == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(200 + static_cast<int>(r));
}

// 2-D map over a packed weight matrix [rows][K] (K contiguous).
int make_wgt_map(CUtensorMap* m, const void* base, int K, int rows, int box_k, int box_rows) {
  if (ensure_driver_api()) return -100;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (box_k == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(300 + static_cast<int>(r));
}

static int set_attrs() {
  static bool done_dev[kMaxDevices] = {};
  bool& g_attr_done = done_dev[device_slot()];
  if (g_attr_done) return 0;
  cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  g_attr_done = true;
  return 0;
}

int launch_conv_gemm(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  int rc = set_attrs();
  if (rc) return rc;
  const int ntiles = p.ncls ? p.cls_tile0[p.ncls] : p.tiles_x * p.tiles_y;
  if (ntiles <= 0) return 0;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  static const int threads = getenv("DSR_CONV_EPI4") ? 192 : kConvThreads;
  launch_k(conv_gemm_kernel, dim3(grid), dim3(threads), kConvSmemBytes, stream, p);
  return static_cast<int>(cudaGetLastError());
}

int halo_smem_bytes(int n_part, int n_wide, int n_narrow, int wide_slots, int ntaps, int slot_bytes) {
  const int bres = ntaps * n_wide * n_part * 128 + ((ntaps * n_narrow * n_part * 32 + 1023) & ~1023);
  return bres + wide_slots * (slot_bytes ? slot_bytes : kHaloWideSlot) + (n_narrow ? 2 * kHaloNarrowSlot : 0) + 256 + 512 +
         1024;                                  // barriers (256) + epilogue bias vector (512) + alignment slack
}

int launch_conv_halo(const HaloParams& p, int num_sms, cudaStream_t stream) {
  static int configured_dev[kMaxDevices] = {};
  int& configured = configured_dev[device_slot()];
  if (p.smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaFuncSetAttribute(conv_halo2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaFuncSetAttribute(conv_halo2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaFuncSetAttribute(conv_halo2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = p.smem_bytes;
  }
  const int ntiles = p.tiles_x * p.tiles_y;
  if (ntiles <= 0) return 0;
  if (p.pair) {                          // CTA pairs: one pair per two tiles, every CTA keeps half of the weights
    const int npairs = (ntiles + 1) / 2;
    int clusters = num_sms / 2;
    if (p.ep_mode == 0 && p.nsplit > 1) {          // N split: every group of nsplit clusters walks the same tile pairs
      int groups = clusters / p.nsplit;
      if (groups > npairs) groups = npairs;
      clusters = groups * p.nsplit;
    } else if (clusters > npairs) {
      clusters = npairs;
    }
    if (p.ep_mode == 3) {                 // generator output conv: its own tile geometry (tiles_x/y count 24 x 8 tiles)
      static bool c9_done_dev[kMaxDevices] = {};
      bool& c9_done = c9_done_dev[device_slot()];
      if (!c9_done) {
        cudaError_t e = cudaFuncSetAttribute(conv9_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC9Smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        c9_done = true;
      }
      launch_k(conv9_out_kernel, dim3(2 * clusters), dim3(kC9Threads), kC9Smem, stream, p);
      return static_cast<int>(cudaGetLastError());
    }
    if (p.ep_mode == 0) launch_k(conv_halo2_kernel<0>, dim3(2 * clusters), dim3(kHalo2Threads), p.smem_bytes, stream, p);
    else if (p.ep_mode == 1) launch_k(conv_halo2_kernel<1>, dim3(2 * clusters), dim3(kHalo2Threads), p.smem_bytes, stream, p);
    else launch_k(conv_halo2_kernel<2>, dim3(2 * clusters), dim3(kHalo2Threads), p.smem_bytes, stream, p);
    return static_cast<int>(cudaGetLastError());
  }
  int grid = (num_sms / p.parts) * p.parts;
  if (grid > ntiles * p.parts) grid = ntiles * p.parts;
  launch_k(conv_halo_kernel, dim3(grid), dim3(kHaloThreads), p.smem_bytes, stream, p);
  return static_cast<int>(cudaGetLastError());
}

static void launch_wgrad_reduce(const float* part, long long stride, int nsplit, float* dw, cudaStream_t stream) {
  const long long n4 = stride / 4;                  // the layer's packed gradient, whole float4s (stride % 4 == 0)
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  launch_k(wgrad_reduce_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, part, stride, nsplit, dw, n4);
}

// how many 3-CTA clusters of wgrad_halo_kernel the device can hold at once (clusters live inside one GPC: 148 SMs do not
// hold 49 of them, and a 49th cluster would run as a second wave)
int wgrad_cluster_capacity() {
  static int cap_dev[kMaxDevices] = {};
  int& cap = cap_dev[device_slot()];
  if (cap > 0) return cap;
  if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgHSmemBytes) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(3 * 64);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = kWgHSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 3;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, wgrad_halo_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  cap = n;
  return cap;
}

int launch_wgrad_halo(const WgHaloParams& p, cudaStream_t stream) {
  static bool done_dev[kMaxDevices] = {};
  bool& done = done_dev[device_slot()];
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgHSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    done = true;
  }
  const int grid = p.ngroups * p.nsplit;
  if (grid <= 0) return 0;
  if (p.cl3) launch_k_cluster(wgrad_halo_kernel, dim3(grid), dim3(kWgThreads), kWgHSmemBytes, stream, 3, p);
  else launch_k(wgrad_halo_kernel, dim3(grid), dim3(kWgThreads), kWgHSmemBytes, stream, p);
  if (p.part != nullptr) launch_wgrad_reduce(p.part, p.part_stride, p.nsplit, p.dw, stream);
  return static_cast<int>(cudaGetLastError());
}

int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
  int rc = set_attrs();
  if (rc) return rc;
  const int grid = p.ngroups * p.nsplit;
  if (grid <= 0) return 0;
  launch_k(wgrad_kernel, dim3(grid), dim3(kWgThreads), kWgSmemBytes, stream, p);
  if (p.part != nullptr) launch_wgrad_reduce(p.part, p.part_stride, p.nsplit, p.dw, stream);
  return static_cast<int>(cudaGetLastError());
}

DSR_KSTAMP_SETTER(kstamp_set_conv)

}  // namespace dsr
