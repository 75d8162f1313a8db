// dsr_elem.cuh -- host-callable launchers of the bandwidth kernels of the DIP step (dsr_elem.cu).
//
// Conventions: "act" tensors are fp16 NHWC on a padded pixel grid [H+2][W+2][C]; pointers passed
// here always point at PADDED pixel (0,0) unless the argument is called *_interior or "plain"
// (= unpadded [H][W][C]).  Gradient tensors are fp16 and hold S * gradient (S = gs[0], see launch_grad_scale_finish).  Statistics blocks are acc_t [2][C]
// (sum, sum of squares) in 64-bit fixed point (dsr_acc.cuh: order-independent, hence run-to-run deterministic),
// accumulated with integer atomics and zeroed by the caller.  Forward sums use the forward scale, everything the
// backward pass accumulates (BN-backward sums, small-layer weight gradients in S * gradient units) the backward scale.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsr_acc.cuh"

namespace dsr {

struct BnRef {            // one train-mode BatchNorm over n pixels
  const acc_t* stats;     // [2][cstride] forward sums of the BN input (fixed point, forward scale)
  const float* gamma;     // [C]
  const float* beta;      // [C]
  float inv_n;            // 1 / (H*W)
  int cstride;            // distance between the sum block and the sum-of-squares block
};

// z fp32 NCHW [C][H][W] (C multiple of 8) -> fp16 padded NHWC with reflected halo.
// z fp32 NCHW -> fp16 padded NHWC with the reflected halo.  Fast path (input_pack_fast: C == 32, W % 4 == 0) can
// fuse level 0's skip-branch 1x1 conv (skip_w -> sraw fp32 [H][W][4], skip_stats [2][4]) and the input perturbation
// (z = perturb_zs + sigma N(0,1), Philox counters of perturb_kernel; z is then the OUTPUT); otherwise both must be null.
int launch_input_pack(float* z, void* xpad, int C, int H, int W, cudaStream_t s, const float* skip_w = nullptr,
                      float* sraw = nullptr, acc_t* skip_stats = nullptr, const float* perturb_zs = nullptr,
                      float sigma = 0.f, unsigned long long seed = 0, const float* state = nullptr, int pad_zero = 0);
int input_pack_fast(int C, int W);

// raw fp16 plain [H][W][128] -> LeakyReLU(BN(raw)) fp16 padded (interior + reflected halo if halo != 0)
// optional fusion of the NEXT level's skip-branch 1x1 conv (128 -> 4) on the activation being written:
// skip_w fp32 [4][128] -> skip_sraw fp32 plain [H][W][4], skip_stats [2][4] (accumulated)
int launch_bn_act(const void* raw, BnRef bn, void* act_pad, int H, int W, int halo, cudaStream_t s,
                  const float* skip_w = nullptr, float* skip_sraw = nullptr, acc_t* skip_stats = nullptr);

// skip branch 1x1 conv: x padded [H+2][W+2][Cin] (fp16) * w fp32 [4][Cin] -> sraw fp32 plain [H][W][4], stats [2][4]
int launch_skip_conv(const void* xpad, int Cin, const float* w, float* sraw, acc_t* stats, int H, int W,
                     cudaStream_t s);

// Upsample x2 (bilinear, align_corners=False) of `deep` + skip activation -> concat statistics / apply.
struct UpcatArgs {
  const void* deep;        // fp16, pixel (0,0) of the low-res activation
  long long deep_sy;       // row pitch in elements (pixel pitch is 128)
  int h, w;                // low-res extent
  int H, W;                // concat extent (H <= 2h, W <= 2w; centre-crop offset is 0)
  const float* sraw;       // fp32 plain [H][W][4]
  BnRef bn_skip;           // BN(4) of the skip branch
  acc_t* cat_stats;        // [2][144]: packed channel order (0..127 upsampled, 128..131 skip)
  const float* cat_gamma;  // BN(132) affine, REFERENCE channel order (0..3 skip, 4..131 upsampled)
  const float* cat_beta;
  void* cat_pad;           // fp16 padded [H+2][W+2][144], reflected halo
  void* qd;                // fp16 plain [h][w][128]: (U^T U d), written by the forward statistics, read by the backward
  int nearest;             // 1: nn.Upsample(mode='nearest') instead of bilinear (skip.py:77)
  int pad_zero;            // 1: zero padding (models/DIP/utils.py:96-102): no reflected halo is written or folded back
};
int launch_upcat_stats(const UpcatArgs& a, cudaStream_t s);
int launch_upcat_apply(const UpcatArgs& a, cudaStream_t s);

// final 1x1 conv 128 -> 3 + bias + sigmoid: act padded [H+2][W+2][128] -> out fp32 NCHW [3][H][W]
int launch_final_conv(const void* act_pad, const float* w, const float* b, float* out, int H, int W, cudaStream_t s);

// ---- backward ------------------------------------------------------------------------------
// d(out) fp32 NCHW, out fp32 NCHW -> dA fp16 padded interior [H+2][W+2][128]; S * dW [3][128] and S * db [3]
// accumulated in fixed point (converted into the gradient buffer by launch_small_grads_finish).
int launch_final_bwd(const float* gout, const float* out, const void* act_pad, const float* w, void* dact_pad,
                     acc_t* dw, acc_t* db, const float* gs, int H, int W, cudaStream_t s);

// fused top of the network (level 0): see dsr_elem.cu
int launch_bn_act_final(const void* raw, BnRef bn, const float* w, const float* b, float* out, int H, int W,
                        cudaStream_t s);
struct TopBwdArgs {
  const float* gout;       // fp32 NCHW [3][H][W]: dL/d(out)
  const float* out;        // fp32 NCHW: sigmoid output of the forward pass
  const void* raw;         // fp16 plain [H][W][128]: BN input of the last decoder conv
  BnRef bn;
  const float* w;          // final conv weight [3][128]
  acc_t* bstats;           // [2][128]
  void* dr_pad;            // fp16 padded [H+2][W+2][128]
  float* dgamma; float* dbeta;       // BN parameter gradients
  acc_t* dw; acc_t* db;              // S * final conv weight / bias gradients (fixed point, accumulated)
  float* gs;
  int H, W;
};
int launch_bn_bwd_top_stats(const TopBwdArgs& a, cudaStream_t s);
int launch_bn_bwd_top_apply(const TopBwdArgs& a, cudaStream_t s);

struct BnBwdArgs {
  const void* g;           // fp16 gradient w.r.t. the activation, PADDED grid [H+2][W+2][gC]
  int gC;                  // channel pitch of g (128 or 144; first 128 channels are used)
  int fold;                // 1: g holds padded-grid data-gradients whose halo must be folded back (reflection)
  // optional (HAS_DS): this activation also feeds the next level's skip branch (1x1 conv 128 -> 4, BN(4)); the whole
  // backward of that branch is done here: dsraw = BN(4)' (dsy), act-gradient += W^T dsraw, dW += dsraw act^T
  const float* dsy;        // fp32 plain [H][W][4]: gradient w.r.t. the skip BN(4) output (null: no skip branch)
  const float* sraw;       // fp32 plain [H][W][4]: skip conv output saved by the forward
  BnRef bn_skip;           // BN(4)
  const acc_t* sbstats;    // [2][4]: sum dsy, sum dsy * xhat
  const float* wskip;      // fp32 [4][128]
  acc_t* dwskip;           // [4][128] S * gradient, fixed point, accumulated
  float* dskip_gamma;      // [4]
  float* dskip_beta;       // [4]
  float* dsraw;            // fp32 plain [H][W][4]: written for inspection (diagnostics)
  const void* raw;         // fp16 plain [H][W][128]: BN input saved by the forward
  BnRef bn;
  acc_t* bstats;           // [2][128]: sum dy, sum dy*xhat (bstats_raw: sum dy, sum dy*r -- see below)
  int bstats_raw;          // 1: the sums were accumulated by the kernel that PRODUCED g (no separate statistics pass),
                           //    in raw form: [1] = sum dy * r; sum dy*xhat = rstd ([1] - mean [0])
  void* dr_pad;            // fp16 padded [H+2][W+2][128] (interior written; halo stays zero)
  float* dgamma;           // [128]
  float* dbeta;            // [128]
  float* gs;               // gradient-scale block {S, 1/S, amax bits, non-finite flag, ...}
  int H, W;
};
int launch_bn_bwd_stats(const BnBwdArgs& a, cudaStream_t s);
int launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t s);

struct UpcatBwdArgs {
  UpcatArgs f;             // forward description (deep, sraw, stats ...)
  const void* gcat;        // fp16 padded grid [H+2][W+2][144]: data-gradient of the 3x3 conv reading cat (to fold)
  acc_t* cbstats;          // [2][144] packed order: sum dc, sum dc*xhat
  void* dup_pad;           // fp16 padded [H+2][W+2][128]: gradient w.r.t. the upsampled tensor (interior)
  float* dsy;              // fp32 plain [H][W][4]: gradient w.r.t. BN(4) output (after LeakyReLU')
  acc_t* sbstats;          // [2][4]
  float* dcat_gamma;       // [132] reference channel order
  float* dcat_beta;
  const float* gs;
  // optional: BatchNorm-backward sums of the layer that CONSUMES ddeep (the deeper level's last decoder conv, or the
  // last level's second encoder conv), accumulated by pass C while ddeep is in registers: sum dy, sum dy * r
  const void* cons_raw;    // fp16 plain [h][w][128]: that layer's BN input (null: no fusion)
  BnRef cons_bn;
  acc_t* cons_bstats;      // [2][128]
};
// source-domain formulation (see dsr_elem.cu): forward statistics of the concat tensor, and the whole backward of
// upsample + concat + BN(132) (a.dup_pad is used as the [h][w][128] scratch tensor t = U^T dc)
int launch_upcat_stats_lowres(const UpcatArgs& a, cudaStream_t s);
int launch_upcat_bwd_gather(const UpcatBwdArgs& a, cudaStream_t s);
int launch_upcat_bwd_apply_lowres(const UpcatBwdArgs& a, void* ddeep_pad, cudaStream_t s);
int launch_upcat_bwd_stats(const UpcatBwdArgs& a, cudaStream_t s);
int launch_upcat_bwd_apply(const UpcatBwdArgs& a, cudaStream_t s);

// BN(4)+skip conv backward: dsy -> dsraw fp32 [H][W][4]; dWskip [4][Cin] (accumulated), dgamma4/dbeta4.
int launch_skip_bwd(const float* dsy, const float* sraw, BnRef bn_skip, const acc_t* sbstats, const void* xpad, int Cin,
                    float* dsraw, acc_t* dw, float* dgamma, float* dbeta, const float* gs, int H, int W,
                    cudaStream_t s);

// transpose of the bilinear x2 upsample: dup padded-interior [H][W][128] (fp16) -> ddeep fp16 padded interior [h][w][128]
int launch_upsample_bwd(const void* dup_pad, int H, int W, void* ddeep_pad, int h, int w, cudaStream_t s);

// ---- parameters ----------------------------------------------------------------------------
struct PackDesc {           // one conv layer's weight (OIHW fp32 in the flat parameter buffer)
  long long w_off;          // offset (floats) into the flat parameter / gradient buffer
  int cout, cin, k;         // k = 1 or 3
  int cin_pad;              // packed K (16-multiple): fprop matrix is [k*k][128][cin_pad]
  int n_rows;               // packed dgrad rows per tap: matrix is [k*k][n_rows][128]  (0: no dgrad needed)
  int perm;                 // 1: packed ci j <-> reference ci (j < 128 ? j + 4 : j - 128)   (concat layer)
  long long f_off, d_off;   // offsets (16-bit elements) into the packed-weight arena
  long long g_off;          // offset (floats) into the packed wgrad arena [k*k][128][cin_pad]
};
int launch_pack_weights(const float* params, void* arena, const PackDesc* table_dev, int nlayers, cudaStream_t s);
int launch_unpack_wgrad(const float* garena, float* grads, const PackDesc* table_dev, int nlayers, const float* gs,
                        cudaStream_t s);
// Gradients of the small 1x1 layers (skip-branch convs, final conv) are accumulated by the element-wise kernels as
// S * gradient in fixed point; this converts them into the flat gradient buffer: grads[g_off + i] = acc[acc_off + i] / S.
struct SmallGradDesc { long long acc_off; long long g_off; int n; int pad_; };   // acc_off: acc_t index from the workspace base
int launch_small_grads_finish(const SmallGradDesc* table_dev, int n, const acc_t* ws_acc, float* grads, const float* gs,
                              cudaStream_t s);
// zero `grads` if the pass saw a non-finite gradient, then adapt the scale for the next pass
int launch_grad_scale_finish(float* grads, long long n, float* gs, cudaStream_t s);

struct BnRunDesc {          // running-statistics update of one BatchNorm (all offsets in floats)
  long long stats_off;      // into the plan workspace viewed as acc_t*: [2][cstride] forward sums
  int cstride; int C; float n;
  long long rm_off, rv_off; // into the flat BatchNorm buffer: running_mean[C], running_var[C]
  long long bias_off;       // into the flat parameter buffer: bias of the conv feeding this BN (the kernels drop
                            // it because BN cancels it, but torch's running_mean contains it), or -1
  int perm;                 // 1: stats are in packed concat order
};
int launch_bn_running(const BnRunDesc* table_dev, int nbn, const acc_t* ws_acc, const float* params, float* bn_buffers,
                      float momentum, cudaStream_t s);

// gs (optional): gradient-scale block of the plan; the update is skipped when the last backward pass overflowed (gs[7])
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                int t, cudaStream_t s, const float* state = nullptr, const float* gs = nullptr);
// t_set > 0: state.t = t_set; t_set == 0: state.t += 1.  Also zeroes loss_base[t - 1] (if given).
int launch_step_begin(float* state, float* loss_base, int t_set, float lr, float b1, float b2, cudaStream_t s,
                      const float* gs = nullptr);

// busy-waits `ns` nanoseconds on the stream (profiling: lets the host run ahead of the GPU)
int launch_spin(unsigned long long ns, cudaStream_t s);

// ---- Lanczos downsampler (fp32 NCHW) -------------------------------------------------------
struct DsTables {           // device tables built by the plan (dsr_downsampler.cu)
  const float* taps;        // [k] normalised 1-D taps (fp32)
  int k, factor, pad;
  // backward gather tables: for every input row y (col x): first output index and NW weights
  const int* by0; const float* bwy;   // [H], [H][nw]
  const int* bx0; const float* bwx;   // [W], [W][nw]
  int nw;
  // fused-MSE launches: fixed-point loss accumulator + arrival ticket (both zero between launches; ONE launch at a
  // time per table set -- launches that share a downsampler must be stream-ordered)
  acc_t* loss_acc; unsigned int* loss_ticket;
};
int launch_downsample_fwd(const float* x, float* y, int C, int H, int W, int oh, int ow, DsTables t, cudaStream_t s);
int launch_downsample_bwd(const float* gy, float* gx, int C, int H, int W, int oh, int ow, DsTables t, cudaStream_t s);
// fused: y = D(x); loss += mean((y - target)^2); gy = 2 (y - target) / N
// loss accumulates into loss[0], or into loss[t - 1] when `state` (device iteration state) is given
int launch_downsample_mse(const float* x, const float* target, float* y, float* gy, float* loss, int C, int H, int W,
                          int oh, int ow, DsTables t, cudaStream_t s, const float* state = nullptr);

// z = z_saved + sigma * N(0,1)   (Philox4x32-10 counter RNG, fp32, elementwise)
int launch_perturb(const float* z_saved, float* z, long long n, float sigma, unsigned long long seed,
                   unsigned long long offset, cudaStream_t s, const float* state = nullptr);

}  // namespace dsr
