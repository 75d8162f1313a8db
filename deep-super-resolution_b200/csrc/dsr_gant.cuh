// dsr_gant.cuh -- SRGAN TRAINING step (train_GAN.py:38-71 of the reference): parameter blocks of the generic
// tensor-core kernels and the host-side tensor / layer records shared by dsr_gant_conv.cu, dsr_gant_elem.cu and
// dsr_gant_plan.cu.
//
// Data layout: every activation / gradient tensor of the training step is bf16 NHWC on a TALL GRID -- the B images of
// the batch are stacked vertically, image b at rows [b * P, b * P + H) of an [B * P][W][C] array; the P - H rows between
// images are zero for the whole life of the workspace (zeroed at bind time, never written) and ARE the zero padding of
// the convolutions (nn.Conv2d(padding=1) / (padding=4) everywhere in models/GAN/*.py and torchvision's VGG19); the
// left / right / top / bottom borders come from TMA's out-of-bounds zero fill.  TMA sees a tensor as
//     (c, px, x, py, y)   sizes (C, S, W / S, S, R / S),  S = 1 (unit stride) or 2 (stride-2 gather, W and R even)
// 3-channel tensors (images, image gradients) use a 16-channel pitch with channels 3..15 zero.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsr {

typedef __nv_bfloat16 bf16_t;

// ---------------------------------------------------------------------------------------------
// gconv_kernel: implicit-GEMM convolution, fprop and dgrad, any (Cin, Cout) in multiples of 64 (or 16 on one side),
// 3 x 3 stride 1 / 2 and 9 x 9 stride 1.  D[128 pixels][nt] = sum over taps, 64-channel chunks  A[128][64] B[nt][64]^T.
// ---------------------------------------------------------------------------------------------
constexpr int kGcThreads = 192;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kGcStages = 5;
constexpr int kGcStageA = 128 * 128;     // 128 pixels x 64 ch x 2 B
constexpr int kGcStageB = 128 * 128;     // up to 128 weight rows x 64 ch x 2 B
constexpr int kGcStageBytes = kGcStageA + kGcStageB;
// ROW-HALO mode (stride-1 convolutions, 8 x 16 pixel tiles): per tap ROW ky and 64-channel chunk ONE input box of
// (8 + ks - 1) x 16 pixels is loaded; the ks taps of the row are column-shifted descriptor views of it (start address
// + shift pixel rows, stride between 8-row groups = the box width; the swizzle follows the absolute shared-memory
// address, tools/umma_probe.cu) -- L2 -> SM traffic for the input drops 2.4 x (3 x 3) / 4.5 x (9 x 9).
constexpr int kGrStages = 3;
constexpr int kGrMaxStage = 68 * 1024;   // 3 x 3, N = 128: 20 KB input box + 3 x 16 KB weights
constexpr int kGcRingBytes = kGrStages * kGrMaxStage;        // 204 KB >= kGcStages * kGcStageBytes (160 KB)
constexpr int kGcSmemBytes = kGcRingBytes + 1024 + 256;

struct GTap { int8_t px, dx, py, dy; int32_t b_row; };

struct alignas(64) GConvParams {
  CUtensorMap a;           // input: box (64 | 16, 1, tw, 1, th)
  CUtensorMap b;           // packed weights [rows][K]: box (64 | 16, nt)
  CUtensorMap b2;          // row-halo mode with one N tile: box (64 | 16, b2_rows) -- the ks taps of a row in b2_loads loads
  int b2_rows, b2_loads;
  GTap taps[9];            // explicit taps (ntaps <= 9) ...
  int ntaps;
  int proc_ks;             // ... or, > 0, generated: tap t = (ky, kx) = (t / ks, t % ks), offsets sign * (k - pad),
  int proc_pad, proc_sign; //     weight rows t * rows_per_tap
  int rows_per_tap;
  int rowhalo;             // 1: row-halo mode (proc_ks taps, tw = 8, th = 16); a_slot / b_slot = bytes of a stage's parts
  int a_slot, b_slot;
  int narrow;              // 1: K of a tap = ONE 16-channel chunk (32 B swizzle); 0: kchunks chunks of 64 (128 B swizzle)
  int kchunks;
  int nt, ntiles_n;        // UMMA N (16, 64 or 128) and number of N tiles
  int tw, th;              // pixel tile, tw * th == 128
  int tiles_x, tiles_y;    // tiles over the launch's pixel grid
  // a launch pixel (j, i) is the output pixel (j * ox_mul + ox_add, i * oy_mul + oy_add) (parity classes of a
  // stride-2 data gradient); it is stored when it lies inside an image: x < out_w, y < out_rows, y % out_period < out_h
  int ox_mul, ox_add, oy_mul, oy_add;
  int out_w, out_h, out_period, out_rows;
  int out_c;               // channel pitch of the output tensor
  void* out;               // bf16, or fp32 when out_f32, or fp16 when out_f16 (the VGG19 forward pass)
  int out_f32, out_f16;
  const float* bias;       // [ntiles_n * nt] or nullptr
  const bf16_t* addend;    // tensor with the output's addressing added before the activation / mask, or nullptr
  const bf16_t* mask;      // data gradient through (Leaky)ReLU: value *= (mask > 0 ? 1 : slope); or nullptr
  float slope;
  int act;                 // forward: value = value > 0 ? value : slope * value
  double* stats;           // optional [2][ntiles_n * nt]: per-channel sum / sum of squares of the STORED values
  uint32_t idesc;
  int* err;
};

// ---------------------------------------------------------------------------------------------
// gwgrad_kernel: dW[tap][co][ci] += sum over pixels dY[p][co] * X[p (+) tap][ci];  M = co tile (128), N = ci tile
// (64 | 128), K = pixels (8 x 8 blocks), both operands MN-major straight from NHWC; up to 3 taps per CTA (TMEM columns
// t * 128); split over pixel ranges with fp32 red.global.add into the packed gradient.
// ---------------------------------------------------------------------------------------------
constexpr int kGwThreads = 192;
constexpr int kGwStages = 3;
constexpr int kGwChunk = 64 * 128;                    // [64 pixels][64 channels] = 8 KB
constexpr int kGwStageBytes = 2 * kGwChunk + 3 * 2 * kGwChunk;   // dY (2 chunks) + 3 taps x X (2 chunks) = 64 KB
constexpr int kGwSmemBytes = kGwStages * kGwStageBytes + 1024 + 256;

struct alignas(64) GWgradParams {
  CUtensorMap a;           // dY (unit view): box (64, 1, 8, 1, 8)
  CUtensorMap b;           // X (view with the conv's stride): box (64, 1, 8, 1, 8)
  GTap taps[9];            // b_row = tap index in the packed gradient
  int ngroups, tpg;        // tap groups (CTA level), taps per group (<= 3)
  int cout, cin;
  int co_tiles, ci_tiles;  // 128-channel tiles (the last one may hold 64)
  int nsplit, pb_x, pb_y;  // pixel-range splits; 8 x 8 pixel blocks over the OUTPUT grid
  float* dw;               // [taps][cout][cin] fp32
  uint32_t idesc64, idesc128;
  int* err;
};

// ---------------------------------------------------------------------------------------------
// host-side records
// ---------------------------------------------------------------------------------------------
struct GT {                // tall-grid tensor
  void* ptr = nullptr;
  int C = 0, W = 0, H = 0, P = 0, B = 0;
  int f32 = 0;
  int f16 = 0;             // 16-bit elements are IEEE half instead of bf16 (VGG19 forward activations)
  int rows() const { return B * P; }
  size_t bytes() const { return static_cast<size_t>(rows()) * W * C * (f32 ? 4 : 2); }
};

int make_gconv_fprop(GConvParams* g, const GT& in, const GT& out, const void* w_pack, int cin_pad, int cout_pad, int ks,
                     int stride, int* err);   // operand format (bf16 / fp16) follows in.f16
int make_gconv_dgrad(GConvParams* g, int* nlaunch, const GT& dy, const GT& dx, const void* w_pack_d, int cin_pad,
                     int cout_pad, int ks, int stride, int* err);
int make_gwgrad(GWgradParams* g, const GT& dy, const GT& x, float* dw, int cin, int cout, int stride, int num_sms,
                int* err);
int make_gconv_taps(GConvParams* g, const GT& in, const GT& out, const void* w, int k_pad, int n_pad, const GTap* taps,
                    int ntaps, int* err);
int make_gwgrad_taps(GWgradParams* g, const GT& dy, const GT& x, float* dw, int cin, int cout, const GTap* taps, int ntaps,
                     int num_sms, int* err);
int launch_gconv(const GConvParams& p, int num_sms, cudaStream_t s);
int launch_gwgrad(const GWgradParams& p, cudaStream_t s);

}  // namespace dsr
