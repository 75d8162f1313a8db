// dsr_conv.cuh -- parameter blocks of the two tcgen05 kernels (implicit-GEMM conv and wgrad).
//
// Data layout (DESIGN.md "HBM layout"): every activation / gradient tensor is NHWC with one image,
// 16-bit elements, stored in a PADDED pixel grid [H+2][W+2][C] (1-pixel halo: reflected values for
// forward activations, zeros for gradients).  TMA sees it as the 5-D tensor
//     (c, px, x, py, y)   sizes (C, PX, Wp/PX, PY, Hp/PY),  PX = PY = 1 (unit stride) or 2 (stride-2 gather)
// so that one box (kc, 1, TW, 1, TH) is a [TW*TH pixels][kc channels] K-major operand tile.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "dsr_acc.cuh"

namespace dsr {

constexpr int kConvThreads = 320;       // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (2 per lane quarter)
constexpr int kConvStages = 5;
constexpr int kConvStageA = 128 * 128;  // 128 pixels x 64 ch x 2 B
constexpr int kConvStageB = 144 * 128;  // up to 144 rows x 64 ch x 2 B
constexpr int kConvStageBytes = kConvStageA + kConvStageB;   // 34816 = 34 * 1024
constexpr int kConvSmemBytes = kConvStages * kConvStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kMaxKBlocks = 28;

// One K-block of the implicit GEMM: which channels of which (shifted) input pixels, which weight rows.
struct KBlk {
  int16_t a_c;       // first channel
  int8_t a_px, a_dx; // coordinate in the px dim, offset added to the tile's x origin
  int8_t a_py, a_dy;
  int8_t wide;       // 1: 64 channels (128B swizzle), 0: 16 channels (32B swizzle)
  int8_t pad_;
  int16_t b_k;       // column (K coordinate) in the packed weight matrix
  int16_t b_row;     // first row in the packed weight matrix
};

struct alignas(64) ConvGemmParams {
  CUtensorMap a64, a16;   // input tensor, boxes of 64 / 16 channels
  CUtensorMap b64, b16;   // packed weights [rows][K], boxes (64|16, n_mma)
  KBlk kb[kMaxKBlocks];
  int nkb;
  // Up to 4 CLASSES share one launch (the 4 output-parity classes of a stride-2 dgrad): class c uses K-blocks
  // [cls_kb0[c], cls_kb0[c] + cls_nkb[c]), writes at out + cls_out_off[c] and owns tiles [cls_tile0[c], cls_tile0[c+1]).
  // ncls == 0 means a single class described by the plain fields below.
  int ncls;
  int cls_kb0[4], cls_nkb[4], cls_tile0[5], cls_out_h[4], cls_out_w[4];
  long long cls_out_off[4];
  int tiles_x, tiles_y;   // tile grid over the output pixel grid
  int tw, th;             // tile = tw x th pixels, tw*th == 128
  int out_h, out_w;       // valid extent of the output pixel grid
  long long out_sy, out_sx;  // output strides in elements
  void* out;              // fp16 (raw conv output) or bf16 (gradient)
  int n_mma;              // UMMA N (multiple of 16, <= 144)
  int n_store;            // channels stored per pixel (multiple of 8, <= n_mma)
  int out_bf16;           // 0: convert accumulators to fp16, 1: to bf16
  acc_t* stats;           // optional [2][n_mma] (fixed point, dsr_acc.cuh): per-channel sum and sum of squares of the STORED values
  uint32_t idesc;
  int* err;
  unsigned long long* prof;   // profiling (bench.py roofline): {min start, max end} in %globaltimer ns, or nullptr
};

// ---------------------------------------------------------------------------------------------
// conv_halo_kernel: stride-1 3x3 convolution (fprop and dgrad) with WEIGHT-STATIONARY CTAs and HALO-TILE reuse.
//   * a CTA owns one slice of N_part output channels for the whole launch and keeps the 9 x N_part x K weight
//     slice resident in shared memory (loaded once by TMA);
//   * per 8 x 16 pixel tile it streams the (8+2) x (16+2) input halo ONCE per 64-channel chunk (one 5-D TMA box,
//     128B-swizzled) and issues the 9 taps as 9 shifted descriptor views of that single tile (start address
//     + (oy*10 + ox) rows, stride between 8-row groups = 10 rows; the swizzle is a function of the absolute
//     shared-memory address, verified on B200 with tools/umma_probe.cu);
//   => shared-memory fill traffic drops from 864 KB to ~50 KB per tile, so the kernel is MMA-bound, not L2-bound.
// ---------------------------------------------------------------------------------------------
constexpr int kHaloThreads = 192;        // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kHalo2Threads = 320;       // pair kernel: warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kHaloTW = 8, kHaloTH = 16; // output tile (pixels); halo = 10 x 18
constexpr int kHaloRows = (kHaloTW + 2) * (kHaloTH + 2);          // 180
constexpr int kHaloWideSlot = 23552;     // 180 rows x 128 B = 23040, rounded up to 1024
constexpr int kHaloNarrowSlot = 6144;    // 180 rows x 32 B = 5760, rounded up to 1024
constexpr int kHaloAccStages = 4;        // TMEM accumulator ring (4 x 64 columns)

struct HaloTap { int8_t oy, ox; int16_t b_row; };   // halo-relative view offset, first row of the weight block

struct alignas(64) HaloParams {
  CUtensorMap a64, a16;      // input (padded grid): boxes (64 | 16, 1, 10, 1, 18)
  CUtensorMap b64, b16;      // packed weights [rows][K]: boxes (64 | 16, n_part)
  HaloTap taps[9];
  int ntaps;                 // 9 (3x3) or 1 (1x1)
  int halo_w, halo_h;        // TMA box extent in pixels: (8+2) x (16+2) for 3x3, 8 x 16 for 1x1
  int n_wide, n_narrow;      // 64-channel chunks (2) and 16-channel chunks (0 or 1) of K
  int n_part, parts;         // output channels per CTA slice (64 or 48), number of slices (2 or 3)
  int org_x, org_y;          // halo origin relative to the tile origin in input coordinates (0 fprop, -1 dgrad)
  int tiles_x, tiles_y;
  int out_h, out_w;
  long long out_sy, out_sx;  // output strides (elements)
  void* out;                 // fp16
  int n_store;               // channels actually stored per pixel (<= n_part * parts)
  acc_t* stats;              // optional [2][stats_stride] (fixed point, dsr_acc.cuh)
  int stats_stride;
  int wide_slots;            // A ring depth for 64-channel chunks (2 or 3)
  uint32_t idesc_wide, idesc_narrow;
  int smem_bytes;
  int pair;                  // 1: CTA-pair kernel (cta_group::2; parts must be 2, n_part = N / 2)
  int dbg;                   // experiments only (DSR_HALO_DBG): bit 0 skip the A loads, bit 1 skip the epilogue work
  int* err;
  // ---- SRResNet generator (dsr_gan.cu; pair kernel only).  ep_mode 0: DIP epilogue (raw fp16 + BatchNorm sums).
  //      1: out = [residual +] prelu(acc + bias[n]) as fp16;  2: 9 x 9 taps (ntaps = 81, tap (ky, kx) = view
  //      ky * halo_w + kx, weight block ky * 9 + kx), out_f32[plane n][pixel] = tanh(acc + bias[n]) for n < n_store.
  //      Batches are TALL grids: image b occupies rows [b * img_rows, b * img_rows + img_h) of the pixel grid, the
  //      rows between images stay zero (the convolution's zero padding; left / right / top / bottom come from
  //      TMA's out-of-bounds zero fill), and output rows that fall into a gap are not stored.
  //      3: conv9_out_kernel (the 9 x 9 output conv with the kx taps folded into N; 24 x 8 pixel tiles).
  int ep_mode;
  int wide_slot_bytes;       // A ring slot size for ep_mode != 0 (halo_w * halo_h * 128 rounded up to 1024)
  int img_rows, img_h;       // tall-grid period (H + gap) and image height
  long long out_img_stride;  // output elements between consecutive images
  long long ep_plane;        // ep_mode 2: elements between output channel planes (NCHW fp32)
  const float* ep_bias;      // [N] fp32 (folded conv bias + eval-mode BatchNorm shift)
  const float* ep_slope;     // device scalar: PReLU slope, or nullptr for no activation
  const void* ep_res;        // fp16 residual tensor with the output's addressing, or nullptr
  // ---- N split across clusters (pair kernel, DIP epilogue; latency-bound small levels): cluster k owns output
  //      channels [slice * 2 n_part, (slice + 1) * 2 n_part) with slice = k % nsplit for the whole launch (its
  //      resident weights are that slice only) and walks tile pairs k / nsplit, k / nsplit + clusters / nsplit, ...
  //      A level with a handful of tiles then spreads over nsplit x more SMs, each with a 1 / nsplit as long MMA
  //      chain per tile.  0 or 1: no split.
  int nsplit;
  unsigned long long* prof;  // profiling: {min start, max end} in %globaltimer ns, or nullptr
};

// ---------------------------------------------------------------------------------------------
// wgrad:  dW[tap][co][ci] = sum over pixels  dR[p][co] * X[p (+) tap][ci]
// GEMM M = co (128), N = ci, K = pixels; both operands are MN-major tiles [pixels][channels].
// ---------------------------------------------------------------------------------------------
constexpr int kWgThreads = 192;
constexpr int kWgStages = 3;
constexpr int kWgPix = 64;                      // pixels per K-block
constexpr int kWgStageA = 2 * kWgPix * 128;     // 2 chunks of 64 co          = 16 KB
constexpr int kWgTapB = 2 * kWgPix * 128 + kWgPix * 32;  // 2x64 ci + 1x16 ci (or 2x16 ci alone) = 18 KB
constexpr int kWgStageBytes = kWgStageA + 3 * kWgTapB;   // 70 KB  (multiple of 1024)
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 1024 + 256;
constexpr int kWgTapCols = 160;                 // TMEM column spacing between taps

struct WgTap {
  int8_t px, dx, py, dy;   // X coordinate: (c, px, x0 + dx, py, y0 + dy)
  int16_t w_tap;           // tap index in the packed gradient [tap][128][ldw]
  int16_t pad_;
};

struct alignas(64) WgradParams {
  CUtensorMap a64;           // dR (bf16, padded grid, unit stride): box (64, 1, pw, 1, ph)
  CUtensorMap b64, b16;      // X (fp16): boxes (64|16, 1, pw, 1, ph)
  WgTap taps[3][3];          // [group][tap in group]
  int ngroups, ntaps;        // groups (CTA-level split of taps), taps per group (<= 3)
  int nsplit;                // pixel-range splits; grid = ngroups * nsplit
  int pb_x, pb_y;            // pixel-block grid: blocks of pw x ph over the conv OUTPUT pixel grid
  int pw, ph;                // pw * ph == 64
  int n64, n16;              // X channel chunks: n64 in {0, 2} of 64 ch, n16 in {0, 1, 2} of 16 ch
  int c16_base;              // first channel of the 16-wide chunks
  int ldw;                   // row pitch (floats) of the packed gradient = padded ci count
  float* dw;                 // [ntaps_total][128][ldw] fp32, accumulated with red.add
  float* part;               // deterministic mode: [nsplit][part_stride] private partials (plain stores), else nullptr
  long long part_stride;     // = ntaps_total * 128 * ldw
  uint32_t idesc64, idesc16;
  int* err;
  unsigned long long* prof;  // profiling: {min start, max end} in %globaltimer ns, or nullptr
};

// ---------------------------------------------------------------------------------------------
// wgrad_halo_kernel: weight gradient with HALO-ROW reuse.  A CTA owns one tap row (ky) and a range of 8 x 8 pixel
// blocks; per block it loads the dR tile once and, instead of one X tile per tap, ONE X box per parity plane that
// is (8 + shift) pixels wide -- the taps of the row are column-shifted descriptor views of that box (MN-major
// operands: a shift is a whole number of 128-byte pixel rows, the swizzle follows the absolute address).
// Shared-memory fill per 64 pixels drops from 70 KB to 39 KB (stride 1, 144 channels).  Layers narrower than 8
// pixels keep the per-tap kernel above.
// ---------------------------------------------------------------------------------------------
#ifndef DSR_WGH_STAGES
#define DSR_WGH_STAGES 3
#endif
// 3 stages = 157 KB: leaves ~70 KB of the SM's shared memory to the MAIN stream's kernels that run beside the
// weight-gradient stream (with 4 stages = 209 KB, upcat_bwd_a's 46 KB CTAs could not co-reside and waited for whole
// wgrad CTAs to retire: 136 us in the replayed graph instead of 48 us alone)
constexpr int kWgHStages = DSR_WGH_STAGES;
constexpr int kWgHStageA = 64 * 128 * 2;          // dR: 2 chunks of [64 px][64 co]            = 16 KB
constexpr int kWgHStageBox = 36 * 1024;           // X boxes of one stage (max: stride 2, 128 ch = 34 KB)
constexpr int kWgHStageBytes = kWgHStageA + kWgHStageBox;
constexpr int kWgHSmemBytes = kWgHStages * kWgHStageBytes + 1024 + 256;
constexpr int kWgC3Rows = 100;                    // cl3: the X halo of a pixel block, 10 x 10 pixels
constexpr int kWgC3Chunk = 13 * 1024;             // 100 rows x 128 B rounded up to the swizzle period
static_assert(2 * kWgC3Chunk + kWgC3Rows * 32 <= kWgHStageBox, "cluster halo fits the X area of a stage");

struct WgBox { int8_t px, py, dx, dy; int16_t width; int16_t off16; };   // off16: offset in the box area / 16
// A RUN = taps of one tap row that are consecutive 1-pixel shifts of the same box.  One MMA per 64-channel chunk
// covers the whole run: N = r x 64 with LBO = one pixel row (128 B), i.e. the N-chunks of the B operand ARE the
// taps, so the dR operand is fetched once per run instead of once per tap (wgrad is operand-fetch bound).
struct WgRun { int8_t box, shift0, r, pad_; int16_t col_wide, col_narrow; };   // first TMEM column of wide / narrow part
struct WgCol { int16_t w_tap, ci0; };              // destination of one 16-column TMEM chunk

struct alignas(64) WgHaloParams {
  CUtensorMap a64;             // dR (padded grid): box (64, 1, 8, 1, 8)
  CUtensorMap b64[2], b16[2];  // X source boxes: (64 | 16, 1, width, 1, 8)
  WgBox box[3][2];             // [group][box]
  WgRun runs[3][3];            // [group][run]
  WgCol cols[3][28];           // [group][TMEM 16-column chunk]
  int nbox, nruns, ncolchunks, ngroups;
  // cl3 = 1 (3 x 3, stride 1): the three tap-row CTAs of one pixel range form a CLUSTER.  Per 8 x 8 pixel block the dR
  // tile and ONE 10 x 10-pixel X halo (maps b64[1] / b16[1]) are fetched once and multicast to all three CTAs (each
  // issues a third of the loads); tap row ky uses the views that start ky halo rows down.  L2 -> SM traffic per block:
  // 45 KB for the cluster instead of 3 x 39 KB -- the un-clustered kernel is bound by exactly that fill rate
  // (tensor pipe 49 % at 512 x 512).
  int cl3;
  int merge_narrow;            // 1: narrow part of a run is one MMA (N = r x 16, LBO = 32 B); 0: n16 chunks, LBO = chunk
  int nsplit, pb_x, pb_y;
  int n64, n16, c16_base, ldw;
  float* dw;
  float* part;                 // deterministic mode: [nsplit][part_stride] private partials, else nullptr
  long long part_stride;
  uint32_t idesc_base;         // kind::f16, both operands MN-major, M = 128, N field left 0
  int* err;
  unsigned long long* prof;    // profiling: {min start, max end} in %globaltimer ns, or nullptr
};

}  // namespace dsr
