// dsr_gant_elem.cuh -- launch wrappers of the bandwidth / CUDA-core kernels of the SRGAN training step
// (dsr_gant_elem.cu).  Tensor convention: dsr_gant.cuh (bf16 NHWC tall grid).
#pragma once
#include "dsr_gant.cuh"

namespace dsr {

struct TG { int C, W, H, P, B; };   // tall-grid geometry passed to kernels by value
inline TG tg_of(const GT& t) { return TG{t.C, t.W, t.H, t.P, t.B}; }

enum { GACT_NONE = 0, GACT_LRELU = 1, GACT_PRELU = 2 };

// images
int gl_pack_image(const float* nchw, const GT& out, cudaStream_t s);                       // fp32 [B][3][H][W] -> NHWC16
int gl_tanh_bwd(const float* dout_nchw, const float* out_nchw, const GT& dz, float* dbias3, cudaStream_t s);
// BatchNorm (batch statistics), activation, residual
// rm / rv (may be nullptr): running statistics, updated run_times times by block 0 (conv_bias: the bias the conv dropped)
int gl_bn_apply(const GT& raw, const GT& out, const bf16_t* res, const double* stats, const float* gamma, const float* beta,
                int act, const float* slope, const float* conv_bias, float* rm, float* rv, int run_times, cudaStream_t s);
// backward of out = act(bn(raw)): dy -> draw, dgamma / dbeta / dslope accumulated (+=).  sums2 = ping-pong scratch of two
// halves of kBnSumsHalf doubles (zero at bind time; the apply kernel clears the half the NEXT call uses, so no memset node
// sits between the kernels of a backward chain), *parity = the half of this call (host state of the owning chain)
constexpr int kBnSumsHalf = 2 * 512 + 8;
int gl_bn_bwd(const GT& dy, const GT& raw, const GT& draw, const double* stats, const float* gamma, const float* beta,
              int act, const float* slope, double* sums2, int* parity, float* dgamma, float* dbeta, float* dslope,
              cudaStream_t s);
// PixelShuffle(2) + PReLU
int gl_shuffle_fwd(const GT& sraw, const GT& u, const float* slope, cudaStream_t s);
int gl_shuffle_bwd(const GT& du, const GT& sraw, const GT& ds, const float* slope, float* dbias256, float* dslope,
                   cudaStream_t s);
// out = prelu(z) and its backward with up to two incoming gradients: dz = (d1 + d2) * prelu'(z)
int gl_prelu_fwd(const GT& z, const GT& out, const float* slope, cudaStream_t s);
int gl_prelu_bwd(const GT& d1, const bf16_t* d2, const GT& z, const GT& dz, const float* slope, float* dbias, float* dslope,
                 cudaStream_t s);
int gl_chan_sum(const GT& t, float* out_c, cudaStream_t s);                                // out_c[c] += sum over pixels
// VGG side
int gl_vgg_pre_fwd(const float* img_nchw, int Hi, int Wi, int Hr, int Wr, int top, int left, const GT& out, cudaStream_t s);
int gl_vgg_pre_bwd(const GT& dpre_f32, int Hi, int Wi, int Hr, int Wr, int top, int left, float* dimg_nchw, int accumulate,
                   cudaStream_t s);
int gl_maxpool_fwd(const GT& in, const GT& out, cudaStream_t s);
int gl_maxpool_bwd(const GT& dout, const GT& y_in, const GT& dz_in, cudaStream_t s);
int gl_feat_mse(const GT& f_fake, const GT& f_real, const GT& dz, double* loss_acc, cudaStream_t s);
// discriminator head
int gl_flatten(const GT& h, float* flat, cudaStream_t s);                                  // -> [B][C*H*W] NCHW order
int gl_unflatten(const float* dflat, const GT& dh, cudaStream_t s);
int gl_dense1_fwd(const float* W, const float* bias, const float* x, float* z1, int B, int K, int J, cudaStream_t s);
int gl_dense2_fwd(const float* z1, const float* w2, const float* b2, float* prob, int B, int J, cudaStream_t s);
// dlogit[b] = dprob[b] * p (1 - p) when dprob != nullptr, else (p - target) / B (BCE, utils/GAN.py:96-107)
int gl_dense2_bwd(const float* prob, const float* dprob, float target, const float* z1, const float* w2, float* dz1,
                  float* dw2, float* db2, int B, int J, cudaStream_t s);
int gl_dense1_bwd(const float* W, const float* x, const float* dz1, float* dW, float* db1, float* dx, int B, int K, int J,
                  cudaStream_t s);
// the same for two groups of B samples (two forward passes) in one sweep over W / dW
int gl_dense1_bwd2(const float* W, const float* x0, const float* dz0, float* dx0, const float* x1, const float* dz1,
                   float* dx1, float* dW, float* db1, int B, int K, int J, int accumulate, cudaStream_t s);
int gl_bce(const float* prob, float target, int B, float* loss_out, int accumulate, cudaStream_t s);
// weights
// bf16 GEMM layouts [tap][Cout][Cin] / [tap][Cin][Cout] (+ padded bias) of ALL layers of a network from its flat fp32
// parameters, and its packed weight gradients added back in OIHW order: one launch each, item tables in device memory
// (built at bind time)
struct GPackItem {
  long long w_off, b_off;        // offsets into the flat parameter array; b_off < 0: bias_pad is written as zero
  bf16_t* w_f; bf16_t* w_d; float* bias_pad;
  int cout, cin, ks, cout_pad, cin_pad, f16_fwd, blk0, nblk;
};
struct GUnpackItem { const float* dw; long long g_off; int cout, cin, ks, blk0, nblk, pad_; };
int gl_group_blocks(long long items);
int gl_pack_group(const float* params, const GPackItem* items_dev, int n, int total_blocks, cudaStream_t s);
int gl_unpack_group(float* grads, const GUnpackItem* items_dev, int n, int total_blocks, cudaStream_t s);
// CUDA-core weight gradients of the layers with 3 INPUT channels
int gl_wgrad_in3(const GT& dy64, const GT& x16, float* g_oihw, int ks, cudaStream_t s);     // dW[64][3][ks][ks] +=
// backward of the 9 x 9 output convolution with the kx taps folded into the channels (dsr_gant_elem.cu)
int gl_expand9(const GT& dz16, const GT& dz9, cudaStream_t s);
int gl_pack9(const float* w_oihw, bf16_t* w9, cudaStream_t s);
// the same fold for a 3 x 3, 3 -> 64 layer's weight gradient: x9[(y, x)][kx * 3 + ci] = img[(y, x + kx - 1)][ci]
int gl_expand3(const GT& img16, const GT& x9, cudaStream_t s);
int gl_unpack3(const float* dw3, float* g_oihw, cudaStream_t s);
int gl_unpack9(const float* dw9, float* g_oihw, cudaStream_t s);
int gl_pack9f(const float* w_oihw, bf16_t* w9f, cudaStream_t s);
int gl_fold9_tanh(const GT& S_f32_32, const float* bias3, float* z16_f32, float* out_nchw, cudaStream_t s);
int gl_finish_double(const double* acc, float* out, float scale, int accumulate, cudaStream_t s);

}  // namespace dsr
