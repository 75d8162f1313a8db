// dsr_gant_plan.cu -- host orchestration and C ABI of the SRGAN TRAINING step.
//
// Reference path (train_GAN.py:38-71, do_epoch): D(HR), G(LR).detach(), D(fake), BCE real/fake, D backward, Adam(D);
// G(LR), D(fake.detach()), PerceptualLoss = MSE(VGG19 relu5_4 features) + BCE(D(fake), 1) (utils/GAN.py:62-123),
// G backward, Adam(G).  Modules: Generator / ResidualBlock / PixelShuffleBlock (models/GAN/generator.py:4-81) in
// TRAIN mode (batch-statistics BatchNorm), Discriminator / DiscriminatorConvBlock (models/GAN/discriminator.py:4-74),
// torchvision VGG19 features[:36] behind VGG19_Weights.IMAGENET1K_V1.transforms() (utils/GAN.py:62-88).
//
// One trainer object per (batch, LR size, factor): tensors on tall bf16 NHWC grids (dsr_gant.cuh) in a caller-owned
// workspace, parameters / gradients / BatchNorm buffers in caller-owned flat fp32 arrays laid out in
// named_parameters() order.  The entry points are the forward / backward halves of the three networks; the Python
// mirror (dsr_b200/gan_train.py) composes them into do_epoch and puts the NCCL gradient all-reduce between them.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dsr_b200.h"
#include "dsr_gant.cuh"
#include "dsr_gant_elem.cuh"
#include "dsr_host.h"

namespace dsr {

struct PEnt { std::string name; long long off, numel; };

struct ConvL {
  int cin = 0, cout = 0, ks = 3, stride = 1, cin_pad = 0, cout_pad = 0;
  long long w_off = 0, b_off = 0;          // offsets into the net's flat parameter array
  bf16_t* w_f = nullptr;                   // [tap][cout_pad][cin_pad]
  bf16_t* w_d = nullptr;                   // [tap][cin_pad][cout_pad]
  float* bias_pad = nullptr;               // [cout_pad]
  float* dw = nullptr;                     // [tap][cout][cin] fp32 (tensor-core weight gradient), or nullptr
  bool need_d = true;
};
struct BnL { int C = 0; long long g_off = 0, b_off = 0, rm_off = 0, rv_off = 0; };

// The launch sequence of an entry point is fixed once the workspace is bound (same tensors, same tensor maps, same
// epilogue pointers every call): the parameter blocks (two cuTensorMapEncodeTiled calls each) are built on the first call
// and replayed afterwards.
struct Tape {
  std::vector<GConvParams> convs;
  std::vector<GWgradParams> wgs;
  size_t ci = 0, wi = 0;
  bool built = false;
  void rewind() { ci = 0; wi = 0; }
  void clear() { convs.clear(); wgs.clear(); built = false; rewind(); }
};

struct Arena {
  uint8_t* base = nullptr;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

}  // namespace dsr

using namespace dsr;

struct dsr_gant {
  int B = 0, lh = 0, lw = 0, factor = 8, blocks = 16, nshuf = 3, with_vgg = 1;
  int H = 0, W = 0;                        // HR size
  int num_sms = 148;
  bool bound = false;
  size_t ws_bytes = 0;
  int* err = nullptr;
  // ---- parameter tables
  std::vector<PEnt> par[3], buf[3];
  long long npar[3] = {0, 0, 0}, nbuf[3] = {0, 0, 0};
  // ---- generator
  ConvL g_conv1, g_conv2, g_conv3;
  std::vector<ConvL> g_ca, g_cb, g_cs;     // block conv1 / conv2, shuffle convs
  std::vector<BnL> g_bna, g_bnb;
  BnL g_bn;
  long long g_prelu1 = 0;
  std::vector<long long> g_prelu_blk, g_prelu_sh;
  GT g_lr16, g_z1, g_x0, g_rt, g_t, g_z, g_dz16, g_dz9;
  GT g_S;                                  // conv3 forward, kx folded into N: fp32 [.][W][32]
  bf16_t* g_w9f = nullptr;                 // conv3 weights [ky][kx * 3 + co][ci] (forward)
  bf16_t* g_w9 = nullptr;                  // conv3 weights with the kx taps folded into K: [ky][ci][kx * 3 + co]
  float* g_dw9 = nullptr;                  // conv3 weight gradient in the folded layout [ky][kx * 3 + co][ci]
  std::vector<GT> g_r1, g_a1, g_r2, g_x;   // g_x[k + 1] = output of block k; g_x[0] aliases g_x0
  std::vector<GT> g_s, g_u, g_ds, g_du;    // shuffle levels
  GT g_dT, g_dX[2], g_dR, g_dA;
  double* g_stats = nullptr;               // [2 * blocks + 1][2 * 64]
  float* g_out = nullptr;                  // copy of the last forward's output image (tanh backward)
  // ---- discriminator
  ConvL d_conv0;
  std::vector<ConvL> d_c;
  std::vector<BnL> d_bn;
  long long d_w1 = 0, d_b1 = 0, d_w2 = 0, d_b2 = 0;
  int d_K = 0;
  GT d_img[2], d_h0[2];
  std::vector<GT> d_raw[2], d_h[2];
  std::vector<GT> d_ga, d_gb;              // gradient scratch per block output / input
  GT d_dz0, d_x9;                          // d_x9: the input image with the kx taps folded into channels (conv0 wgrad)
  float* d_dw3 = nullptr;                  // conv0 weight gradient in the folded layout [ky][co][kx * 3 + ci]
  double* d_stats[2] = {nullptr, nullptr}; // [7][2 * 512]
  float *d_flat[2] = {nullptr, nullptr}, *d_z1[2] = {nullptr, nullptr}, *d_prob[2] = {nullptr, nullptr};
  float *d_dz1[2] = {nullptr, nullptr}, *d_dflat[2] = {nullptr, nullptr};
  // ---- VGG19
  std::vector<ConvL> v_c;
  std::vector<GT> v_y;                     // 16 conv outputs (post-ReLU)
  std::vector<GT> v_pool, v_dpool;         // 4 pool outputs and their gradients
  std::vector<GT> v_ga, v_gb;              // gradient scratch per level
  GT v_pre, v_dpre, v_freal;
  int v_Hr = 0, v_Wr = 0, v_top = 0, v_left = 0;
  // ---- shared
  double* sums = nullptr;                  // BatchNorm-backward scratch [2 * 512 + 1] of the generator's backward pass
  int sums_par = 0, sums_d_par = 0;        // ping-pong halves of the two scratch areas (gl_bn_bwd)
  double* sums_d = nullptr;                // the discriminator's own (its backward may run on another stream, beside G's)
  double* loss_acc = nullptr;
  float* dw_arena[2] = {nullptr, nullptr}; // packed weight gradients of G / D
  size_t dw_bytes[2] = {0, 0};
  int launches = 0;
  int dense_overwrite = 0;                 // d_backward_pair writes (not adds) dense1.weight's gradient
  // one-launch weight packing / gradient unpacking per network: item tables (host copy, device copy, grid size)
  std::vector<GPackItem> h_pack[3];
  std::vector<GUnpackItem> h_unpack[2];
  GPackItem* t_pack[3] = {nullptr, nullptr, nullptr};
  GUnpackItem* t_unpack[2] = {nullptr, nullptr};
  int pack_blocks[3] = {0, 0, 0}, unpack_blocks[2] = {0, 0};
  Tape tp_gf, tp_gb, tp_df[2], tp_db[2], tp_v, tp_vl, tp_vr, tp_v2, tp_vl2;
  Tape* tape = nullptr;                    // the running entry point's tape
};

namespace {

constexpr int kVggPoolAfter[16] = {0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0};
constexpr int kVggIdx[16] = {0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28, 30, 32, 34};
constexpr int kVggCout[16] = {64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512, 512};

int pad_c(int c) { return c < 16 ? 16 : c; }

void add_par(std::vector<PEnt>& v, long long& total, const std::string& name, long long numel, long long* off) {
  v.push_back(PEnt{name, total, numel});
  if (off) *off = total;
  total += numel;
}

void add_conv(dsr_gant* p, int net, const std::string& name, ConvL& c, int cin, int cout, int ks, int stride) {
  c.cin = cin; c.cout = cout; c.ks = ks; c.stride = stride;
  c.cin_pad = pad_c(cin); c.cout_pad = pad_c(cout);
  add_par(p->par[net], p->npar[net], name + ".weight", static_cast<long long>(cout) * cin * ks * ks, &c.w_off);
  add_par(p->par[net], p->npar[net], name + ".bias", cout, &c.b_off);
}
void add_bn(dsr_gant* p, int net, const std::string& name, BnL& b, int C) {
  b.C = C;
  add_par(p->par[net], p->npar[net], name + ".weight", C, &b.g_off);
  add_par(p->par[net], p->npar[net], name + ".bias", C, &b.b_off);
  add_par(p->buf[net], p->nbuf[net], name + ".running_mean", C, &b.rm_off);
  add_par(p->buf[net], p->nbuf[net], name + ".running_var", C, &b.rv_off);
}

GT mk(Arena& a, int C, int W, int H, int P, int B, int f32 = 0) {
  GT t;
  t.C = C; t.W = W; t.H = H; t.P = P; t.B = B; t.f32 = f32;
  t.ptr = a.take(t.bytes());
  return t;
}

void layout_pack(Arena& a, ConvL& c) {
  const size_t n = static_cast<size_t>(c.ks) * c.ks * c.cout_pad * c.cin_pad;
  c.w_f = static_cast<bf16_t*>(a.take(n * 2));
  c.w_d = c.need_d ? static_cast<bf16_t*>(a.take(n * 2)) : nullptr;
  c.bias_pad = static_cast<float*>(a.take(static_cast<size_t>(c.cout_pad) * 4));
}
// packed fp32 weight gradients of the tensor-core layers of one network: one contiguous arena (a single memset clears it)
float* layout_dw(Arena& a, std::vector<ConvL*>& cs, size_t* bytes) {
  float* start = static_cast<float*>(a.take(0));
  const size_t off0 = a.off;
  for (ConvL* c : cs) {
    c->dw = reinterpret_cast<float*>(a.base ? a.base + a.off : nullptr);
    a.off += static_cast<size_t>(c->ks) * c->ks * c->cout * c->cin * 4;
  }
  *bytes = a.off - off0;
  return start;
}

// All tensors, weight packs and scratch of the trainer, in one pass (base == nullptr: size only).
size_t layout(dsr_gant* p, uint8_t* base) {
  Arena a;
  a.base = base;
  const int B = p->B;
  p->err = static_cast<int*>(a.take(1024));
  for (int n = 0; n < 3; ++n) p->t_pack[n] = static_cast<GPackItem*>(a.take(48 * sizeof(GPackItem)));
  for (int n = 0; n < 2; ++n) p->t_unpack[n] = static_cast<GUnpackItem*>(a.take(48 * sizeof(GUnpackItem)));
  p->sums = static_cast<double*>(a.take(2 * kBnSumsHalf * sizeof(double)));
  p->sums_d = static_cast<double*>(a.take(2 * kBnSumsHalf * sizeof(double)));
  p->loss_acc = static_cast<double*>(a.take(64));
  // ---------------- generator ----------------
  {
    const int nl = p->nshuf + 1;
    std::vector<int> Hs(nl), Ws(nl), Ps(nl);
    for (int l = 0; l < nl; ++l) { Hs[l] = p->lh << l; Ws[l] = p->lw << l; Ps[l] = Hs[l] + 4; }
    p->g_lr16 = mk(a, 16, Ws[0], Hs[0], Ps[0], B);
    p->g_z1 = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_x0 = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_r1.resize(p->blocks); p->g_a1.resize(p->blocks); p->g_r2.resize(p->blocks); p->g_x.resize(p->blocks + 1);
    p->g_x[0] = p->g_x0;
    for (int k = 0; k < p->blocks; ++k) {
      p->g_r1[k] = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
      p->g_a1[k] = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
      p->g_r2[k] = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
      p->g_x[k + 1] = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    }
    p->g_rt = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_t = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_s.resize(p->nshuf); p->g_u.resize(p->nshuf); p->g_ds.resize(p->nshuf); p->g_du.resize(p->nshuf);
    for (int j = 0; j < p->nshuf; ++j) {
      p->g_s[j] = mk(a, 256, Ws[j], Hs[j], Ps[j], B);
      p->g_ds[j] = mk(a, 256, Ws[j], Hs[j], Ps[j], B);
      p->g_u[j] = mk(a, 64, Ws[j + 1], Hs[j + 1], Ps[j + 1], B);
      p->g_du[j] = mk(a, 64, Ws[j + 1], Hs[j + 1], Ps[j + 1], B);
    }
    p->g_z = mk(a, 16, Ws[nl - 1], Hs[nl - 1], Ps[nl - 1], B, 1);
    p->g_dz16 = mk(a, 16, Ws[nl - 1], Hs[nl - 1], Ps[nl - 1], B);
    p->g_dz9 = mk(a, 64, Ws[nl - 1], Hs[nl - 1], Ps[nl - 1], B);
    p->g_S = mk(a, 32, Ws[nl - 1], Hs[nl - 1], Ps[nl - 1], B, 1);
    p->g_w9f = static_cast<bf16_t*>(a.take(9 * 32 * 64 * 2));
    p->g_w9 = static_cast<bf16_t*>(a.take(9 * 64 * 64 * 2));
    p->g_dw9 = static_cast<float*>(a.take(9 * 64 * 64 * 4));
    p->g_dT = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_dX[0] = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_dX[1] = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_dR = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_dA = mk(a, 64, Ws[0], Hs[0], Ps[0], B);
    p->g_stats = static_cast<double*>(a.take(static_cast<size_t>(2 * p->blocks + 1) * 128 * sizeof(double)));
    p->g_out = static_cast<float*>(a.take(static_cast<size_t>(B) * 3 * p->H * p->W * 4));
    std::vector<ConvL*> tc;
    for (auto& c : p->g_ca) tc.push_back(&c);
    for (auto& c : p->g_cb) tc.push_back(&c);
    for (auto& c : p->g_cs) tc.push_back(&c);
    tc.push_back(&p->g_conv2);
    p->dw_arena[0] = layout_dw(a, tc, &p->dw_bytes[0]);
    p->g_conv1.need_d = false;
    layout_pack(a, p->g_conv1);
    for (ConvL* c : tc) layout_pack(a, *c);
  }
  // ---------------- discriminator ----------------
  {
    const int P0 = 16 * (p->H / 16 + 1);
    for (int s = 0; s < 2; ++s) {
      p->d_img[s] = mk(a, 16, p->W, p->H, P0, B);
      p->d_h0[s] = mk(a, 64, p->W, p->H, P0, B);
      p->d_raw[s].resize(7); p->d_h[s].resize(7);
      int h = p->H, w = p->W, P = P0;
      for (int k = 0; k < 7; ++k) {
        if (p->d_c[k].stride == 2) { h /= 2; w /= 2; P /= 2; }
        p->d_raw[s][k] = mk(a, p->d_c[k].cout, w, h, P, B);
        p->d_h[s][k] = mk(a, p->d_c[k].cout, w, h, P, B);
      }
      p->d_stats[s] = static_cast<double*>(a.take(7 * 1024 * sizeof(double)));
      p->d_flat[s] = static_cast<float*>(a.take(static_cast<size_t>(B) * p->d_K * 4));
      p->d_z1[s] = static_cast<float*>(a.take(static_cast<size_t>(B) * 1024 * 4));
      p->d_prob[s] = static_cast<float*>(a.take(64));
    }
    p->d_ga.resize(7); p->d_gb.resize(7);
    for (int k = 0; k < 7; ++k) {
      const GT& o = p->d_raw[0][k];
      p->d_ga[k] = mk(a, o.C, o.W, o.H, o.P, B);     // gradient w.r.t. h[k] (block output)
      p->d_gb[k] = mk(a, o.C, o.W, o.H, o.P, B);     // gradient w.r.t. raw[k]
    }
    p->d_dz0 = mk(a, 64, p->W, p->H, P0, B);
    p->d_x9 = mk(a, 64, p->W, p->H, P0, B);
    p->d_dw3 = static_cast<float*>(a.take(3 * 64 * 64 * 4));
    for (int s2 = 0; s2 < 2; ++s2) {
      p->d_dz1[s2] = static_cast<float*>(a.take(static_cast<size_t>(B) * 1024 * 4));
      p->d_dflat[s2] = static_cast<float*>(a.take(static_cast<size_t>(B) * p->d_K * 4));
    }
    std::vector<ConvL*> tc;
    for (auto& c : p->d_c) tc.push_back(&c);
    p->dw_arena[1] = layout_dw(a, tc, &p->dw_bytes[1]);
    p->d_conv0.need_d = false;
    layout_pack(a, p->d_conv0);
    for (ConvL* c : tc) layout_pack(a, *c);
  }
  // ---------------- VGG19 ----------------
  if (p->with_vgg) {
    int h = 224, w = 224;
    // The VGG19 FORWARD pass runs in fp16 (tensors marked f16, fp16 weight copies): the content loss is a difference
    // of two nearly equal feature maps, so the 3 extra mantissa bits matter (gradient w.r.t. the image: cosine 0.79
    // with bf16 features at 192 x 192); the frozen network has no weight gradient, so no kernel
    // ever mixes the fp16 activations with the bf16 gradients of the backward chain.
    p->v_pre = mk(a, 16, w, h, h + 2, B);
    p->v_pre.f16 = 1;
    p->v_dpre = mk(a, 16, w, h, h + 2, B, 1);
    p->v_y.resize(16); p->v_pool.clear(); p->v_dpool.clear(); p->v_ga.clear(); p->v_gb.clear();
    int level_c = 64;
    for (int i = 0; i < 16; ++i) {
      p->v_y[i] = mk(a, kVggCout[i], w, h, h + 2, B);
      p->v_y[i].f16 = 1;
      level_c = kVggCout[i];
      if (kVggPoolAfter[i]) {
        p->v_ga.push_back(mk(a, level_c, w, h, h + 2, B));
        p->v_gb.push_back(mk(a, level_c, w, h, h + 2, B));
        h /= 2; w /= 2;
        p->v_pool.push_back(mk(a, level_c, w, h, h + 2, B));
        p->v_pool.back().f16 = 1;
        p->v_dpool.push_back(mk(a, level_c, w, h, h + 2, B));
      }
    }
    p->v_ga.push_back(mk(a, 512, w, h, h + 2, B));
    p->v_gb.push_back(mk(a, 512, w, h, h + 2, B));
    p->v_freal = mk(a, 512, w, h, h + 2, B);
    p->v_freal.f16 = 1;
    for (int i = 0; i < 16; ++i) {
      p->v_c[i].need_d = true;
      layout_pack(a, p->v_c[i]);
    }
  }
  return a.off + 1024;
}

#define GCHK(call)                 \
  do {                             \
    const int rc_ = (call);        \
    if (rc_) return rc_;           \
    ++p->launches;                 \
  } while (0)

int run_fprop(dsr_gant* p, const ConvL& c, const GT& in, const GT& out, bool bias, int act, float slope, double* stats,
              cudaStream_t s) {
  Tape& t = *p->tape;
  if (!t.built) {
    GConvParams g;
    int rc = make_gconv_fprop(&g, in, out, c.w_f, c.cin_pad, c.cout_pad, c.ks, c.stride, p->err);
    if (rc) return rc;
    g.bias = bias ? c.bias_pad : nullptr;
    g.act = act; g.slope = slope; g.stats = stats;
    t.convs.push_back(g);
  }
  return launch_gconv(t.convs[t.ci++], p->num_sms, s);
}
int run_dgrad(dsr_gant* p, const ConvL& c, const GT& dy, const GT& dx, const bf16_t* addend, const bf16_t* mask, float slope,
              cudaStream_t s) {
  Tape& t = *p->tape;
  const int n = (c.stride == 1) ? 1 : 4;
  if (!t.built) {
    GConvParams g[4];
    int nn = 0;
    int rc = make_gconv_dgrad(g, &nn, dy, dx, c.w_d, c.cin_pad, c.cout_pad, c.ks, c.stride, p->err);
    if (rc) return rc;
    for (int i = 0; i < nn; ++i) {
      g[i].addend = addend; g[i].mask = mask; g[i].slope = slope;
      t.convs.push_back(g[i]);
    }
  }
  for (int i = 0; i < n; ++i) {
    const int rc = launch_gconv(t.convs[t.ci++], p->num_sms, s);
    if (rc) return rc;
  }
  return 0;
}
int run_wgrad(dsr_gant* p, const ConvL& c, const GT& dy, const GT& x, cudaStream_t s) {
  Tape& t = *p->tape;
  if (!t.built) {
    GWgradParams g;
    int rc = make_gwgrad(&g, dy, x, c.dw, c.cin, c.cout, c.stride, p->num_sms, p->err);
    if (rc) return rc;
    t.wgs.push_back(g);
  }
  return launch_gwgrad(t.wgs[t.wi++], s);
}
struct TapeScope {                       // selects an entry point's tape; marks it built when the call got to its end
  dsr_gant* p;
  Tape* t;
  bool ok = false;
  TapeScope(dsr_gant* pp, Tape* tt) : p(pp), t(tt) { t->rewind(); p->tape = t; }
  ~TapeScope() {
    if (ok) t->built = true;
    else if (!t->built) t->clear();
    p->tape = nullptr;
  }
};

}  // namespace

extern "C" {

int dsr_gant_create(dsr_gant_t** out, int batch, int lr_h, int lr_w, int factor, int residual_blocks, int with_vgg) {
  if (!out || batch < 1 || batch > 8 || lr_h < 8 || lr_w < 8 || (factor != 8 && factor != 16) || residual_blocks < 1)
    return -1;
  if ((lr_w % 2) || (lr_h % 2)) return -5;
  dsr_gant* p = new dsr_gant();
  p->B = batch; p->lh = lr_h; p->lw = lr_w; p->factor = factor; p->blocks = residual_blocks;
  p->nshuf = factor == 8 ? 3 : 4;
  p->H = lr_h * factor; p->W = lr_w * factor;
  p->with_vgg = with_vgg;
  if ((p->H % 16) || (p->W % 16)) { delete p; return -5; }
  if (with_vgg && (p->H > 256 || p->W > 256)) { delete p; return -5; }   // the transform only ENLARGES here (resize to 256)
  int dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); p->num_sms = 148; }
  // ---- generator parameters, named_parameters() order of models/GAN/generator.py:44-66
  add_conv(p, 0, "conv1", p->g_conv1, 3, 64, 9, 1);
  add_par(p->par[0], p->npar[0], "prelu1.weight", 1, &p->g_prelu1);
  p->g_ca.resize(p->blocks); p->g_cb.resize(p->blocks); p->g_bna.resize(p->blocks); p->g_bnb.resize(p->blocks);
  p->g_prelu_blk.resize(p->blocks);
  for (int k = 0; k < p->blocks; ++k) {
    const std::string pre = "residual_blocks." + std::to_string(k) + ".";
    add_conv(p, 0, pre + "conv1", p->g_ca[k], 64, 64, 3, 1);
    add_bn(p, 0, pre + "bn1", p->g_bna[k], 64);
    add_par(p->par[0], p->npar[0], pre + "prelu1.weight", 1, &p->g_prelu_blk[k]);
    add_conv(p, 0, pre + "conv2", p->g_cb[k], 64, 64, 3, 1);
    add_bn(p, 0, pre + "bn2", p->g_bnb[k], 64);
  }
  add_conv(p, 0, "conv2", p->g_conv2, 64, 64, 3, 1);
  add_bn(p, 0, "bn1", p->g_bn, 64);
  p->g_cs.resize(p->nshuf); p->g_prelu_sh.resize(p->nshuf);
  for (int j = 0; j < p->nshuf; ++j) {
    const std::string pre = "pixel_shuffle_blocks." + std::to_string(j) + ".";
    add_conv(p, 0, pre + "conv1", p->g_cs[j], 64, 256, 3, 1);
    add_par(p->par[0], p->npar[0], pre + "prelu1.weight", 1, &p->g_prelu_sh[j]);
  }
  add_conv(p, 0, "conv3", p->g_conv3, 64, 3, 9, 1);
  // ---- discriminator (models/GAN/discriminator.py:21-47)
  add_conv(p, 1, "conv", p->d_conv0, 3, 64, 3, 1);
  const int dc[7][3] = {{64, 64, 2}, {64, 128, 1}, {128, 128, 2}, {128, 256, 1}, {256, 256, 2}, {256, 512, 1}, {512, 512, 2}};
  p->d_c.resize(7); p->d_bn.resize(7);
  for (int k = 0; k < 7; ++k) {
    const std::string pre = "convblocks." + std::to_string(k) + ".";
    add_conv(p, 1, pre + "conv1", p->d_c[k], dc[k][0], dc[k][1], 3, dc[k][2]);
    add_bn(p, 1, pre + "bn1", p->d_bn[k], dc[k][1]);
  }
  p->d_K = 512 * (p->H / 16) * (p->W / 16);
  add_par(p->par[1], p->npar[1], "dense1.weight", 1024LL * p->d_K, &p->d_w1);
  add_par(p->par[1], p->npar[1], "dense1.bias", 1024, &p->d_b1);
  add_par(p->par[1], p->npar[1], "dense2.weight", 1024, &p->d_w2);
  add_par(p->par[1], p->npar[1], "dense2.bias", 1, &p->d_b2);
  // ---- VGG19 features[:36] (utils/GAN.py:66-69): keys net.0.<idx>.weight / .bias
  p->v_c.resize(16);
  int cin = 3;
  for (int i = 0; i < 16; ++i) {
    add_conv(p, 2, "net.0." + std::to_string(kVggIdx[i]), p->v_c[i], cin, kVggCout[i], 3, 1);
    cin = kVggCout[i];
  }
  // torchvision: resize the smaller edge to 256 (the other edge int(256 * long / short)), centre crop 224
  if (p->H <= p->W) { p->v_Hr = 256; p->v_Wr = static_cast<int>(256LL * p->W / p->H); }
  else { p->v_Wr = 256; p->v_Hr = static_cast<int>(256LL * p->H / p->W); }
  p->v_top = static_cast<int>(lround((p->v_Hr - 224) / 2.0));
  p->v_left = static_cast<int>(lround((p->v_Wr - 224) / 2.0));
  p->ws_bytes = layout(p, nullptr);
  *out = p;
  return 0;
}

void dsr_gant_destroy(dsr_gant_t* p) { delete p; }

long long dsr_gant_param_numel(const dsr_gant_t* p, int net) { return (p && net >= 0 && net < 3) ? p->npar[net] : -1; }
long long dsr_gant_buffer_numel(const dsr_gant_t* p, int net) { return (p && net >= 0 && net < 3) ? p->nbuf[net] : -1; }
int dsr_gant_num_params(const dsr_gant_t* p, int net) { return (p && net >= 0 && net < 3) ? static_cast<int>(p->par[net].size()) : -1; }
int dsr_gant_num_buffers(const dsr_gant_t* p, int net) { return (p && net >= 0 && net < 3) ? static_cast<int>(p->buf[net].size()) : -1; }
static int ent_info(const std::vector<PEnt>& v, int idx, char* name, int cap, long long* off, long long* numel) {
  if (idx < 0 || idx >= static_cast<int>(v.size()) || !name || cap < 1) return -1;
  snprintf(name, static_cast<size_t>(cap), "%s", v[idx].name.c_str());
  if (off) *off = v[idx].off;
  if (numel) *numel = v[idx].numel;
  return 0;
}
int dsr_gant_param_info(const dsr_gant_t* p, int net, int idx, char* name, int cap, long long* off, long long* numel) {
  if (!p || net < 0 || net > 2) return -1;
  return ent_info(p->par[net], idx, name, cap, off, numel);
}
int dsr_gant_buffer_info(const dsr_gant_t* p, int net, int idx, char* name, int cap, long long* off, long long* numel) {
  if (!p || net < 0 || net > 2) return -1;
  return ent_info(p->buf[net], idx, name, cap, off, numel);
}
size_t dsr_gant_workspace_bytes(const dsr_gant_t* p) { return p ? p->ws_bytes : 0; }

static void add_pack_item(dsr_gant* p, int net, const ConvL& c, bool with_bias, int f16_fwd) {
  GPackItem q{};
  q.w_off = c.w_off; q.b_off = with_bias ? c.b_off : -1;
  q.w_f = c.w_f; q.w_d = c.w_d; q.bias_pad = c.bias_pad;
  q.cout = c.cout; q.cin = c.cin; q.ks = c.ks; q.cout_pad = c.cout_pad; q.cin_pad = c.cin_pad; q.f16_fwd = f16_fwd;
  q.blk0 = p->pack_blocks[net];
  q.nblk = gl_group_blocks(static_cast<long long>(c.ks) * c.ks * c.cout_pad * c.cin_pad);
  p->pack_blocks[net] += q.nblk;
  p->h_pack[net].push_back(q);
}
static void add_unpack_item(dsr_gant* p, int net, const ConvL& c) {
  GUnpackItem q{};
  q.dw = c.dw; q.g_off = c.w_off; q.cout = c.cout; q.cin = c.cin; q.ks = c.ks;
  q.blk0 = p->unpack_blocks[net];
  q.nblk = gl_group_blocks(static_cast<long long>(c.ks) * c.ks * c.cout * c.cin);
  p->unpack_blocks[net] += q.nblk;
  p->h_unpack[net].push_back(q);
}
// the item tables of the grouped pack / unpack launches (pointers known once the workspace is laid out)
static int build_group_tables(dsr_gant* p, cudaStream_t s) {
  for (int n = 0; n < 3; ++n) { p->h_pack[n].clear(); p->pack_blocks[n] = 0; }
  for (int n = 0; n < 2; ++n) { p->h_unpack[n].clear(); p->unpack_blocks[n] = 0; }
  add_pack_item(p, 0, p->g_conv1, true, 0);
  for (int k = 0; k < p->blocks; ++k) {
    add_pack_item(p, 0, p->g_ca[k], false, 0);
    add_pack_item(p, 0, p->g_cb[k], false, 0);
    add_unpack_item(p, 0, p->g_ca[k]);
    add_unpack_item(p, 0, p->g_cb[k]);
  }
  add_pack_item(p, 0, p->g_conv2, false, 0);
  add_unpack_item(p, 0, p->g_conv2);
  for (int j = 0; j < p->nshuf; ++j) {
    add_pack_item(p, 0, p->g_cs[j], true, 0);
    add_unpack_item(p, 0, p->g_cs[j]);
  }
  add_pack_item(p, 1, p->d_conv0, true, 0);
  for (int k = 0; k < 7; ++k) {
    add_pack_item(p, 1, p->d_c[k], false, 0);
    add_unpack_item(p, 1, p->d_c[k]);
  }
  if (p->with_vgg)
    for (int i = 0; i < 16; ++i) add_pack_item(p, 2, p->v_c[i], true, 1);     // forward copies in fp16 (see layout)
  for (int n = 0; n < 3; ++n) {
    if (p->h_pack[n].size() > 48) return -57;
    if (!p->h_pack[n].empty() &&
        cudaMemcpyAsync(p->t_pack[n], p->h_pack[n].data(), p->h_pack[n].size() * sizeof(GPackItem), cudaMemcpyHostToDevice, s) != cudaSuccess)
      return -58;
  }
  for (int n = 0; n < 2; ++n) {
    if (p->h_unpack[n].size() > 48) return -57;
    if (cudaMemcpyAsync(p->t_unpack[n], p->h_unpack[n].data(), p->h_unpack[n].size() * sizeof(GUnpackItem), cudaMemcpyHostToDevice, s) != cudaSuccess)
      return -58;
  }
  return 0;
}

int dsr_gant_bind(dsr_gant_t* p, void* workspace, size_t bytes, void* stream) {
  if (!p || !workspace) return -1;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return -3;
  if (bytes < p->ws_bytes) return -8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(workspace, 0, p->ws_bytes, s);      // gap rows and padded channels stay zero from here on
  if (e != cudaSuccess) return static_cast<int>(e);
  layout(p, static_cast<uint8_t*>(workspace));
  for (Tape* t : {&p->tp_gf, &p->tp_gb, &p->tp_df[0], &p->tp_df[1], &p->tp_db[0], &p->tp_db[1], &p->tp_v, &p->tp_vl, &p->tp_vr,
                  &p->tp_v2, &p->tp_vl2})
    t->clear();
  p->sums_par = p->sums_d_par = 0;
  const int rc = build_group_tables(p, s);
  if (rc) return rc;
  p->bound = true;
  return ensure_driver_api();
}

int dsr_gant_device_error(dsr_gant_t* p, int* host_code) {
  if (!p || !p->bound || !host_code) return -1;
  cudaError_t e = cudaMemcpy(host_code, p->err, sizeof(int), cudaMemcpyDeviceToHost);
  return static_cast<int>(e);
}
int dsr_gant_last_launches(const dsr_gant_t* p) { return p ? p->launches : -1; }

// bf16 GEMM layouts of one network's convolution weights from its flat fp32 parameters; call after every update.
int dsr_gant_pack(dsr_gant_t* p, int net, const float* params, void* stream) {
  if (!p || !p->bound || !params || net < 0 || net > 2) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  if (net == 2 && !p->with_vgg) return -5;
  GCHK(gl_pack_group(params, p->t_pack[net], static_cast<int>(p->h_pack[net].size()), p->pack_blocks[net], s));
  if (net == 0) {
    GCHK(gl_pack9(params + p->g_conv3.w_off, p->g_w9, s));
    GCHK(gl_pack9f(params + p->g_conv3.w_off, p->g_w9f, s));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Generator, train mode (generator.py:68-81).  bn_updates: how many times the running statistics take this batch
// (do_epoch runs the generator twice on the same weights and batch, train_GAN.py:46,56 -- the second run is identical).
// ---------------------------------------------------------------------------------------------
int dsr_gant_g_forward(dsr_gant_t* p, const float* params, float* buffers, const float* lr_nchw, float* out_nchw,
                       int bn_updates, void* stream) {
  if (!p || !p->bound || !params || !lr_nchw || !out_nchw) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  TapeScope scope(p, &p->tp_gf);
  cudaMemsetAsync(p->g_stats, 0, static_cast<size_t>(2 * p->blocks + 1) * 128 * sizeof(double), s);
  GCHK(gl_pack_image(lr_nchw, p->g_lr16, s));
  GCHK(run_fprop(p, p->g_conv1, p->g_lr16, p->g_z1, true, 0, 0.f, nullptr, s));
  GCHK(gl_prelu_fwd(p->g_z1, p->g_x0, params + p->g_prelu1, s));
  for (int k = 0; k < p->blocks; ++k) {
    double* sa = p->g_stats + (2 * k) * 128;
    double* sb = p->g_stats + (2 * k + 1) * 128;
    const BnL& ba = p->g_bna[k];
    const BnL& bb = p->g_bnb[k];
    GCHK(run_fprop(p, p->g_ca[k], p->g_x[k], p->g_r1[k], false, 0, 0.f, sa, s));
    const bool upd = buffers && bn_updates > 0;
    GCHK(gl_bn_apply(p->g_r1[k], p->g_a1[k], nullptr, sa, params + ba.g_off, params + ba.b_off, GACT_PRELU,
                     params + p->g_prelu_blk[k], params + p->g_ca[k].b_off, upd ? buffers + ba.rm_off : nullptr,
                     upd ? buffers + ba.rv_off : nullptr, bn_updates, s));
    GCHK(run_fprop(p, p->g_cb[k], p->g_a1[k], p->g_r2[k], false, 0, 0.f, sb, s));
    GCHK(gl_bn_apply(p->g_r2[k], p->g_x[k + 1], static_cast<const bf16_t*>(p->g_x[k].ptr), sb, params + bb.g_off,
                     params + bb.b_off, GACT_NONE, nullptr, params + p->g_cb[k].b_off, upd ? buffers + bb.rm_off : nullptr,
                     upd ? buffers + bb.rv_off : nullptr, bn_updates, s));
  }
  double* st = p->g_stats + (2 * p->blocks) * 128;
  GCHK(run_fprop(p, p->g_conv2, p->g_x[p->blocks], p->g_rt, false, 0, 0.f, st, s));
  const bool updt = buffers && bn_updates > 0;
  GCHK(gl_bn_apply(p->g_rt, p->g_t, static_cast<const bf16_t*>(p->g_x0.ptr), st, params + p->g_bn.g_off,
                   params + p->g_bn.b_off, GACT_NONE, nullptr, params + p->g_conv2.b_off,
                   updt ? buffers + p->g_bn.rm_off : nullptr, updt ? buffers + p->g_bn.rv_off : nullptr, bn_updates, s));
  const GT* cur = &p->g_t;
  for (int j = 0; j < p->nshuf; ++j) {
    GCHK(run_fprop(p, p->g_cs[j], *cur, p->g_s[j], true, 0, 0.f, nullptr, s));
    GCHK(gl_shuffle_fwd(p->g_s[j], p->g_u[j], params + p->g_prelu_sh[j], s));
    cur = &p->g_u[j];
  }
  // conv3 (9 x 9, 64 -> 3) with the kx taps folded into N: a 9-tap (ky) GEMM with N = 27 -> 32, then the 9-term fold inside
  // the tanh / NCHW output pass (dsr_gant_elem.cu) -- 9 x fewer MMAs than the tap-by-tap form
  {
    Tape& t = *p->tape;
    if (!t.built) {
      GTap tf[9];
      for (int ky = 0; ky < 9; ++ky) tf[ky] = GTap{0, 0, 0, static_cast<int8_t>(ky - 4), ky * 32};
      GConvParams gc;
      const int rc = make_gconv_taps(&gc, *cur, p->g_S, p->g_w9f, 64, 32, tf, 9, p->err);
      if (rc) return rc;
      t.convs.push_back(gc);
    }
    GCHK(launch_gconv(t.convs[t.ci++], p->num_sms, s));
  }
  GCHK(gl_fold9_tanh(p->g_S, params + p->g_conv3.b_off, static_cast<float*>(p->g_z.ptr), out_nchw, s));
  cudaMemcpyAsync(p->g_out, out_nchw, static_cast<size_t>(p->B) * 3 * p->H * p->W * 4, cudaMemcpyDeviceToDevice, s);
  scope.ok = true;
  return static_cast<int>(cudaGetLastError());
}

// grads += d(loss)/d(params) for the forward pass above, given d(loss)/d(output image) [B][3][H][W] fp32.
int dsr_gant_g_backward(dsr_gant_t* p, const float* params, const float* dout_nchw, float* grads, void* stream) {
  if (!p || !p->bound || !params || !dout_nchw || !grads) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  TapeScope scope(p, &p->tp_gb);
  cudaMemsetAsync(p->dw_arena[0], 0, p->dw_bytes[0], s);
  cudaMemsetAsync(p->g_dw9, 0, 9 * 64 * 64 * 4, s);           // here, not between the kernels below (a memset node breaks their overlap)
  const int last = p->nshuf - 1;
  GCHK(gl_tanh_bwd(dout_nchw, p->g_out, p->g_dz16, grads + p->g_conv3.b_off, s));
  // conv3 (9 x 9, 64 -> 3) backward with the kx taps folded into the channels: 9-tap tensor-core problems (dsr_gant_elem.cu)
  GCHK(gl_expand9(p->g_dz16, p->g_dz9, s));
  {
    Tape& t = *p->tape;
    if (!t.built) {
      GTap tw[9], td[9];
      for (int ky = 0; ky < 9; ++ky) {
        tw[ky] = GTap{0, 0, 0, static_cast<int8_t>(ky - 4), ky};              // x[q + (ky - 4, 0)], gradient block ky
        td[ky] = GTap{0, 0, 0, static_cast<int8_t>(4 - ky), ky * 64};          // dz9[q - (ky - 4, 0)], weight rows ky * 64
      }
      GWgradParams gw;
      int rc = make_gwgrad_taps(&gw, p->g_dz9, p->g_u[last], p->g_dw9, 64, 64, tw, 9, p->num_sms, p->err);
      if (rc) return rc;
      t.wgs.push_back(gw);
      GConvParams gc;
      if ((rc = make_gconv_taps(&gc, p->g_dz9, p->g_du[last], p->g_w9, 64, 64, td, 9, p->err))) return rc;
      t.convs.push_back(gc);
    }
    GCHK(launch_gwgrad(t.wgs[t.wi++], s));
    GCHK(launch_gconv(t.convs[t.ci++], p->num_sms, s));
  }
  GCHK(gl_unpack9(p->g_dw9, grads + p->g_conv3.w_off, s));
  for (int j = last; j >= 0; --j) {
    const GT& in = (j == 0) ? p->g_t : p->g_u[j - 1];
    const GT& din = (j == 0) ? p->g_dT : p->g_du[j - 1];
    GCHK(gl_shuffle_bwd(p->g_du[j], p->g_s[j], p->g_ds[j], params + p->g_prelu_sh[j], grads + p->g_cs[j].b_off,
                        grads + p->g_prelu_sh[j], s));
    GCHK(run_wgrad(p, p->g_cs[j], p->g_ds[j], in, s));
    GCHK(run_dgrad(p, p->g_cs[j], p->g_ds[j], din, nullptr, nullptr, 0.f, s));
  }
  // t = x0 + bn(conv2(x_last))
  double* st = p->g_stats + (2 * p->blocks) * 128;
  GCHK(gl_bn_bwd(p->g_dT, p->g_rt, p->g_dR, st, params + p->g_bn.g_off, params + p->g_bn.b_off, GACT_NONE, nullptr, p->sums, &p->sums_par,
                 grads + p->g_bn.g_off, grads + p->g_bn.b_off, nullptr, s));
  p->launches += 3;
  GCHK(run_wgrad(p, p->g_conv2, p->g_dR, p->g_x[p->blocks], s));
  int cur = 0;
  GCHK(run_dgrad(p, p->g_conv2, p->g_dR, p->g_dX[cur], nullptr, nullptr, 0.f, s));
  for (int k = p->blocks - 1; k >= 0; --k) {
    double* sa = p->g_stats + (2 * k) * 128;
    double* sb = p->g_stats + (2 * k + 1) * 128;
    const BnL& ba = p->g_bna[k];
    const BnL& bb = p->g_bnb[k];
    GCHK(gl_bn_bwd(p->g_dX[cur], p->g_r2[k], p->g_dR, sb, params + bb.g_off, params + bb.b_off, GACT_NONE, nullptr, p->sums, &p->sums_par,
                   grads + bb.g_off, grads + bb.b_off, nullptr, s));
    GCHK(run_wgrad(p, p->g_cb[k], p->g_dR, p->g_a1[k], s));
    GCHK(run_dgrad(p, p->g_cb[k], p->g_dR, p->g_dA, nullptr, nullptr, 0.f, s));
    GCHK(gl_bn_bwd(p->g_dA, p->g_r1[k], p->g_dR, sa, params + ba.g_off, params + ba.b_off, GACT_PRELU,
                   params + p->g_prelu_blk[k], p->sums, &p->sums_par, grads + ba.g_off, grads + ba.b_off, grads + p->g_prelu_blk[k], s));
    GCHK(run_wgrad(p, p->g_ca[k], p->g_dR, p->g_x[k], s));
    GCHK(run_dgrad(p, p->g_ca[k], p->g_dR, p->g_dX[cur ^ 1], static_cast<const bf16_t*>(p->g_dX[cur].ptr), nullptr, 0.f, s));
    p->launches += 6;
    cur ^= 1;
  }
  GCHK(gl_prelu_bwd(p->g_dX[cur], static_cast<const bf16_t*>(p->g_dT.ptr), p->g_z1, p->g_dR, params + p->g_prelu1,
                    grads + p->g_conv1.b_off, grads + p->g_prelu1, s));
  GCHK(gl_wgrad_in3(p->g_dR, p->g_lr16, grads + p->g_conv1.w_off, 9, s));
  GCHK(gl_unpack_group(grads, p->t_unpack[0], static_cast<int>(p->h_unpack[0].size()), p->unpack_blocks[0], s));
  scope.ok = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Discriminator, train mode (discriminator.py:57-74).  `slot` (0 / 1) selects one of two activation sets so that the
// real and the fake pass of do_epoch (train_GAN.py:44-48) can both be back-propagated afterwards.
// ---------------------------------------------------------------------------------------------
int dsr_gant_d_forward(dsr_gant_t* p, int slot, const float* params, float* buffers, const float* img_nchw, float* prob,
                       void* stream) {
  if (!p || !p->bound || !params || !img_nchw || !prob || slot < 0 || slot > 1) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  TapeScope scope(p, &p->tp_df[slot]);
  cudaMemsetAsync(p->d_stats[slot], 0, 7 * 1024 * sizeof(double), s);
  GCHK(gl_pack_image(img_nchw, p->d_img[slot], s));
  GCHK(run_fprop(p, p->d_conv0, p->d_img[slot], p->d_h0[slot], true, 1, 0.2f, nullptr, s));
  const GT* cur = &p->d_h0[slot];
  for (int k = 0; k < 7; ++k) {
    double* st = p->d_stats[slot] + k * 1024;
    const BnL& b = p->d_bn[k];
    const GT& raw = p->d_raw[slot][k];
    GCHK(run_fprop(p, p->d_c[k], *cur, raw, false, 0, 0.f, st, s));
    GCHK(gl_bn_apply(raw, p->d_h[slot][k], nullptr, st, params + b.g_off, params + b.b_off, GACT_LRELU, nullptr,
                     params + p->d_c[k].b_off, buffers ? buffers + b.rm_off : nullptr, buffers ? buffers + b.rv_off : nullptr,
                     1, s));
    cur = &p->d_h[slot][k];
  }
  GCHK(gl_flatten(*cur, p->d_flat[slot], s));
  GCHK(gl_dense1_fwd(params + p->d_w1, params + p->d_b1, p->d_flat[slot], p->d_z1[slot], p->B, p->d_K, 1024, s));
  GCHK(gl_dense2_fwd(p->d_z1[slot], params + p->d_w2, params + p->d_b2, p->d_prob[slot], p->B, 1024, s));
  cudaMemcpyAsync(prob, p->d_prob[slot], static_cast<size_t>(p->B) * 4, cudaMemcpyDeviceToDevice, s);
  scope.ok = true;
  return static_cast<int>(cudaGetLastError());
}

// the convolutional part of the discriminator backward, from the gradient of the flattened features on
// clear / unpack: the packed weight gradients (tensor-core layers and the folded conv0) are cleared before and added to
// `grads` after this pass; the paired call clears once, lets both passes accumulate, and unpacks once
static int d_convs_backward(dsr_gant* p, int slot, const float* params, float* grads, cudaStream_t s, bool clear = true,
                            bool unpack = true) {
  TapeScope scope(p, &p->tp_db[slot]);
  if (clear) {
    cudaMemsetAsync(p->dw_arena[1], 0, p->dw_bytes[1], s);
    cudaMemsetAsync(p->d_dw3, 0, 3 * 64 * 64 * 4, s);
  }
  GCHK(gl_unflatten(p->d_dflat[slot], p->d_ga[6], s));
  for (int k = 6; k >= 0; --k) {
    double* st = p->d_stats[slot] + k * 1024;
    const BnL& b = p->d_bn[k];
    const GT& in = (k == 0) ? p->d_h0[slot] : p->d_h[slot][k - 1];
    GCHK(gl_bn_bwd(p->d_ga[k], p->d_raw[slot][k], p->d_gb[k], st, params + b.g_off, params + b.b_off, GACT_LRELU, nullptr,
                   p->sums_d, &p->sums_d_par, grads + b.g_off, grads + b.b_off, nullptr, s));
    p->launches += 1;
    GCHK(run_wgrad(p, p->d_c[k], p->d_gb[k], in, s));
    if (k > 0) {
      GCHK(run_dgrad(p, p->d_c[k], p->d_gb[k], p->d_ga[k - 1], nullptr, nullptr, 0.f, s));
    } else {
      // through conv0's LeakyReLU: h0 > 0 <=> its pre-activation > 0
      GCHK(run_dgrad(p, p->d_c[0], p->d_gb[0], p->d_dz0, nullptr, static_cast<const bf16_t*>(p->d_h0[slot].ptr), 0.2f, s));
    }
    if (p->d_c[k].stride == 2) p->launches += 3;
  }
  // conv0 (3 x 3, 3 -> 64) weight gradient with the kx taps folded into the input channels: a 3-tap gwgrad_kernel
  GCHK(gl_expand3(p->d_img[slot], p->d_x9, s));
  {
    Tape& t = *p->tape;
    if (!t.built) {
      GTap tw[3];
      for (int ky = 0; ky < 3; ++ky) tw[ky] = GTap{0, 0, 0, static_cast<int8_t>(ky - 1), ky};
      GWgradParams gw;
      const int rc = make_gwgrad_taps(&gw, p->d_dz0, p->d_x9, p->d_dw3, 64, 64, tw, 3, p->num_sms, p->err);
      if (rc) return rc;
      t.wgs.push_back(gw);
    }
    GCHK(launch_gwgrad(t.wgs[t.wi++], s));
  }
  GCHK(gl_chan_sum(p->d_dz0, grads + p->d_conv0.b_off, s));
  if (unpack) {
    GCHK(gl_unpack3(p->d_dw3, grads + p->d_conv0.w_off, s));
    GCHK(gl_unpack_group(grads, p->t_unpack[1], static_cast<int>(p->h_unpack[1].size()), p->unpack_blocks[1], s));
  }
  scope.ok = true;
  return 0;
}

// grads += gradient of the pass in `slot`.  dprob: d(loss)/d(prob) [B] (autograd path), or nullptr for the fused
// BCE against the constant `target` with mean reduction (utils/GAN.py:96-107): d(loss)/d(logit) = (p - target) / B.
int dsr_gant_d_backward(dsr_gant_t* p, int slot, const float* params, const float* dprob, float target, float* grads,
                        void* stream) {
  if (!p || !p->bound || !params || !grads || slot < 0 || slot > 1) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  GCHK(gl_dense2_bwd(p->d_prob[slot], dprob, target, p->d_z1[slot], params + p->d_w2, p->d_dz1[slot], grads + p->d_w2,
                     grads + p->d_b2, p->B, 1024, s));
  GCHK(gl_dense1_bwd(params + p->d_w1, p->d_flat[slot], p->d_dz1[slot], grads + p->d_w1, grads + p->d_b1, p->d_dflat[slot],
                     p->B, p->d_K, 1024, s));
  return d_convs_backward(p, slot, params, grads, s);
}

// Both kept passes at once (loss_D.backward() of do_epoch, train_GAN.py:52): slot 0 against target0, slot 1 against
// target1, BCE fused; the dense head's weight matrix and its gradient are swept once for the two passes.
int dsr_gant_d_backward_pair(dsr_gant_t* p, const float* params, float target0, float target1, float* grads, void* stream) {
  if (!p || !p->bound || !params || !grads) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  const float tg[2] = {target0, target1};
  for (int slot = 0; slot < 2; ++slot)
    GCHK(gl_dense2_bwd(p->d_prob[slot], nullptr, tg[slot], p->d_z1[slot], params + p->d_w2, p->d_dz1[slot], grads + p->d_w2,
                       grads + p->d_b2, p->B, 1024, s));
  GCHK(gl_dense1_bwd2(params + p->d_w1, p->d_flat[0], p->d_dz1[0], p->d_dflat[0], p->d_flat[1], p->d_dz1[1], p->d_dflat[1],
                      grads + p->d_w1, grads + p->d_b1, p->B, p->d_K, 1024, p->dense_overwrite ? 0 : 1, s));
  const int n = p->launches;
  int rc = d_convs_backward(p, 0, params, grads, s, true, false);
  if (rc) return rc;
  const int n0 = p->launches;
  rc = d_convs_backward(p, 1, params, grads, s, false, true);
  p->launches += n0 - n;
  return rc;
}

int dsr_gant_dense_grad_overwrite(dsr_gant_t* p, int on) {
  if (!p) return -1;
  p->dense_overwrite = on ? 1 : 0;
  return 0;
}

int dsr_gant_bce(dsr_gant_t* p, const float* prob, float target, int n, float* loss, int accumulate, void* stream) {
  if (!p || !prob || !loss || n < 1) return -1;
  return gl_bce(prob, target, n, loss, accumulate, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// Perceptual content loss (utils/GAN.py:62-88): loss[0] (= or +=) MSE(VGG(T(fake)), VGG(T(real))), T = the
// IMAGENET1K_V1 transform; dfake (optional) = d(loss)/d(fake) [B][3][H][W] fp32.
// ---------------------------------------------------------------------------------------------
static int vgg_forward(dsr_gant* p, const float* img, cudaStream_t s) {
  GCHK(gl_vgg_pre_fwd(img, p->H, p->W, p->v_Hr, p->v_Wr, p->v_top, p->v_left, p->v_pre, s));
  const GT* cur = &p->v_pre;
  int pl = 0;
  for (int i = 0; i < 16; ++i) {
    GCHK(run_fprop(p, p->v_c[i], *cur, p->v_y[i], true, 1, 0.f, nullptr, s));
    cur = &p->v_y[i];
    if (kVggPoolAfter[i]) {
      GCHK(gl_maxpool_fwd(*cur, p->v_pool[pl], s));
      cur = &p->v_pool[pl++];
    }
  }
  return 0;
}

// The features of the REAL batch alone (they do not depend on the generator: a caller may compute them on another stream
// while the generator's forward pass runs); dsr_gant_vgg_loss(..., real_nchw = NULL, ...) then uses them.  The pass
// writes the VGG activation buffers, so it must be complete before that call starts and must not overlap another one.
int dsr_gant_vgg_real(dsr_gant_t* p, const float* real_nchw, void* stream) {
  if (!p || !p->bound || !p->with_vgg || !real_nchw) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  TapeScope scope(p, &p->tp_vr);
  int rc;
  if ((rc = vgg_forward(p, real_nchw, s))) return rc;
  cudaMemcpyAsync(p->v_freal.ptr, p->v_y[15].ptr, p->v_freal.bytes(), cudaMemcpyDeviceToDevice, s);
  scope.ok = true;
  return 0;
}

int dsr_gant_vgg_loss(dsr_gant_t* p, const float* fake_nchw, const float* real_nchw, float* loss, int accumulate,
                      float* dfake_nchw, void* stream) {
  if (!p || !p->bound || !p->with_vgg || !fake_nchw || !loss) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  TapeScope scope(p, real_nchw ? (dfake_nchw ? &p->tp_v : &p->tp_vl) : (dfake_nchw ? &p->tp_v2 : &p->tp_vl2));
  int rc;
  if (real_nchw) {
    if ((rc = vgg_forward(p, real_nchw, s))) return rc;
    cudaMemcpyAsync(p->v_freal.ptr, p->v_y[15].ptr, p->v_freal.bytes(), cudaMemcpyDeviceToDevice, s);
  }
  if ((rc = vgg_forward(p, fake_nchw, s))) return rc;
  cudaMemsetAsync(p->loss_acc, 0, sizeof(double), s);
  const int top = static_cast<int>(p->v_ga.size()) - 1;
  GCHK(gl_feat_mse(p->v_y[15], p->v_freal, p->v_gb[top], p->loss_acc, s));
  const GT& f = p->v_y[15];
  GCHK(gl_finish_double(p->loss_acc, loss, 1.f / (static_cast<float>(f.B) * f.H * f.W * f.C), accumulate, s));
  if (dfake_nchw == nullptr) { scope.ok = true; return 0; }
  // v_gb[level] holds d(loss)/d(pre-activation) of the current conv, v_ga[level] is the other buffer of the level
  int level = top;
  GT dz = p->v_gb[level];
  bool in_b = true;
  for (int i = 15; i >= 1; --i) {
    const bool pooled = kVggPoolAfter[i - 1] != 0;     // conv i reads pool(y[i-1])
    if (!pooled) {
      const GT& dst = in_b ? p->v_ga[level] : p->v_gb[level];
      GCHK(run_dgrad(p, p->v_c[i], dz, dst, nullptr, static_cast<const bf16_t*>(p->v_y[i - 1].ptr), 0.f, s));
      dz = dst;
      in_b = !in_b;
    } else {
      --level;
      const GT& dpool = p->v_dpool[level];
      GCHK(run_dgrad(p, p->v_c[i], dz, dpool, nullptr, nullptr, 0.f, s));
      GT dst = p->v_gb[level];
      GCHK(gl_maxpool_bwd(dpool, p->v_y[i - 1], dst, s));
      dz = dst;
      in_b = true;
    }
  }
  GCHK(run_dgrad(p, p->v_c[0], dz, p->v_dpre, nullptr, nullptr, 0.f, s));
  GCHK(gl_vgg_pre_bwd(p->v_dpre, p->H, p->W, p->v_Hr, p->v_Wr, p->v_top, p->v_left, dfake_nchw, 0, s));
  scope.ok = true;
  return 0;
}

/* Tests: a named activation of the last pass (tall bf16 / fp32 grid). */
int dsr_gant_tensor(const dsr_gant_t* p, const char* name, void** ptr, int* C, int* W, int* H, int* P, int* B, int* f32) {
  if (!p || !p->bound || !name || !ptr) return -1;
  const GT* t = nullptr;
  std::string n(name);
  auto idx = [&](const char* pre) { return atoi(n.c_str() + strlen(pre)); };
  if (n == "g_z1") t = &p->g_z1;
  else if (n == "g_x0") t = &p->g_x0;
  else if (n == "g_t") t = &p->g_t;
  else if (n == "g_z") t = &p->g_z;
  else if (n.rfind("g_x", 0) == 0) { const int k = idx("g_x"); if (k >= 0 && k <= p->blocks) t = &p->g_x[k]; }
  else if (n.rfind("g_u", 0) == 0) { const int k = idx("g_u"); if (k >= 0 && k < p->nshuf) t = &p->g_u[k]; }
  else if (n == "d_h0") t = &p->d_h0[0];
  else if (n.rfind("d_h", 0) == 0) { const int k = idx("d_h") - 1; if (k >= 0 && k < 7) t = &p->d_h[0][k]; }
  else if (n == "v_pre") t = &p->v_pre;
  else if (n.rfind("v_y", 0) == 0) { const int k = idx("v_y"); if (k >= 0 && k < 16 && p->with_vgg) t = &p->v_y[k]; }
  if (!t || !t->ptr) return -1;
  *ptr = t->ptr;
  if (C) *C = t->C;
  if (W) *W = t->W;
  if (H) *H = t->H;
  if (P) *P = t->P;
  if (B) *B = t->B;
  if (f32) *f32 = t->f32 ? 1 : (t->f16 ? 2 : 0);     // 0 bf16, 1 fp32, 2 fp16
  return 0;
}

}  // extern "C"
