// dsr_plan.cu -- host side of libdsr_b200.so: the execution plan of the skip network (layer table,
// workspace layout, TMA descriptors, forward / backward / whole-step launch sequences) and the
// extern "C" entry points declared in include/dsr_b200.h.
//
// Reference structure being replaced (paths relative to the upstream repo):
//   models/DIP/skip.py:41-94      per-level module graph (skip branch, two encoder convs, recursion,
//                                 upsample, Concat, BN(132), decoder 3x3 + 1x1, final 1x1 + sigmoid)
//   models/DIP/utils.py:5-8       1-based child naming (state_dict keys)
//   DIP.py:47-95, utils/DIP.py:35-38   closure + Adam loop (dsr_dip_step)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dsr_b200.h"
#include "dsr_conv.cuh"
#include "dsr_debug.h"
#include "dsr_elem.cuh"
#include "dsr_host.h"
#include "dsr_launch.cuh"
#include "dsr_ptx.cuh"

namespace dsr {

void lanczos_taps(int factor, int support, std::vector<double>& taps);
void ds_bwd_table(int n, int on, int factor, int k, int pad, const std::vector<float>& taps, int& nw,
                  std::vector<int>& o0, std::vector<float>& w);

namespace {

constexpr int kNC = 128;       // n33d = n33u (DIP.py:170-171)
constexpr int kNS = 4;         // skip_n11   (DIP.py:172)
constexpr int kCat = 144;      // packed concat channel pitch: 128 upsampled + 4 skip + 12 zero
constexpr float kMomentum = 0.1f;
constexpr int kProfSlots = 2048;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ParamInfo {
  std::string name;
  long long off;
  int ndim;
  int shape[4];
};
struct BnInfo {
  std::string name;
  long long off;   // running_mean at off, running_var at off + C
  int C;
};

// A tensor of the workspace (offset resolved to a pointer at bind time).
struct Buf {
  size_t off = 0, bytes = 0;
  void* ptr = nullptr;
};

struct TensorInfo {
  std::string name;
  const Buf* buf;     // workspace tensor, or nullptr when `acc_off` (a float index from the base) is used
  size_t acc_off;
  int kind;      // 0 fp16, 1 bf16, 2 fp32
  int padded;
  int H, W, C;
};

// One tensor-core convolution (encoder d1 / d2, decoder u1 / u2 of a level) and its BatchNorm.
struct ConvLayer {
  const char* tag = "";
  int cin = 0, cin_pad = 0, k = 3, stride = 1;
  int inH = 0, inW = 0, outH = 0, outW = 0;
  bool need_dgrad = true;
  bool bstats_ready = false;       // transient: BN-backward sums of this layer were accumulated by the kernel that produced its g
  int n_rows = 0;                  // dgrad output channels (packed): 128 or 144
  long long w_off = 0, b_off = 0, g_off = 0, be_off = 0;   // conv weight / bias, BN gamma / beta (floats)
  long long bn_off = 0;            // BN buffers
  PackDesc pack{};
  const Buf* in_pad = nullptr;     // fp16 padded input [inH+2][inW+2][cin_pad]
  Buf raw;                         // fp16 plain [outH][outW][128]
  Buf act;                         // fp16 padded [outH+2][outW+2][128]
  int act_halo = 0;                // write the reflected halo of `act`
  size_t stats_off = 0;            // acc_t index from the workspace base: [2][128] forward sums (fixed point)
  size_t bstats_off = 0;           // [2][128] backward sums
  Buf dr;                          // fp16 padded [outH+2][outW+2][128]   gradient w.r.t. raw
  Buf gin;                         // fp16 padded [inH+2][inW+2][n_rows]  data gradient (to fold)
  ConvGemmParams fprop{};
  ConvGemmParams dgrad[4];
  int ndgrad = 0;
  ConvGemmParams dgrad_merged{};   // stride 2: the 4 parity classes as ONE launch (product path)
  bool has_merged = false;
  HaloParams hfprop{}, hdgrad{};   // weight-stationary halo-tile variants (stride-1 layers)
  bool has_halo = false;
  WgHaloParams hwgrad{};           // halo-row wgrad (layers at least 8 pixels wide)
  bool has_hwgrad = false;
  WgradParams wgrad{};
  ActRef ref_in{}, ref_dr{};       // checker views
  WgtRef ref_wf{}, ref_wd{};
};

struct Level {
  int H = 0, W = 0, h = 0, w = 0, Cin = 0;
  long long skip_w = 0, skip_b = 0, skip_g = 0, skip_be = 0, skip_bn = 0;
  long long cat_g = 0, cat_be = 0, cat_bn = 0;
  ConvLayer d1, d2, u1, u2;
  Buf xin;                          // level 0 only: packed input
  const Buf* x_pad = nullptr;       // fp16 padded [H+2][W+2][Cin]
  Buf sraw;                         // fp32 [H][W][4]
  Buf cat;                          // fp16 padded [H+2][W+2][144]
  size_t skip_stats_off = 0, cat_stats_off = 0;     // forward accumulators ([2][4], [2][144]), acc_t indices
  size_t sbstats_off = 0, cbstats_off = 0;          // backward accumulators
  size_t skip_dw_acc = 0;                           // backward: S * d(skip conv weight) [4][Cin], fixed point
  Buf g_u2a;                        // fp16 padded [H+2][W+2][128]: gradient w.r.t. this level's output
  Buf dup;                          // fp16 padded [H+2][W+2][128]
  Buf dsy, dsraw;                   // fp32 [H][W][4]
  Buf g_d2a;                        // last level only: bf16 padded [h+2][w+2][128]
};

}  // namespace
}  // namespace dsr

using namespace dsr;

struct dsr_downsampler {
  int factor, support, H, W, oh, ow, k, pad, nw_y, nw_x;
  unsigned long long serial = 0;   // unique per created object: a graph captured for a destroyed downsampler is never
                                   // replayed for a new one that happens to land at the same host address
  std::vector<float> taps;
  std::vector<int> by0, bx0;
  std::vector<float> bwy, bwx;
  DsTables t{};
};

struct dsr_plan {
  int H, W, input_depth, num_scales, n_out;
  int pad_zero = 0;     // get_net(pad='zero'): Conv2d zero padding instead of ReflectionPad2d (models/DIP/utils.py:96-102)
  int nearest = 0;      // get_net(upsample_mode='nearest') (models/DIP/skip.py:77)
  std::vector<Level> lv;
  std::vector<ParamInfo> params;
  std::vector<BnInfo> bns;
  long long nparam = 0, nbn = 0;
  long long fin_w = 0, fin_b = 0;
  // workspace
  size_t ws_bytes = 0;
  size_t acc_fwd_off = 0, acc_fwd_bytes = 0;      // zeroed at the start of every forward (fixed-point accumulators)
  size_t acc_bwd_off = 0, acc_bwd_bytes = 0;      // zeroed at the start of every backward (incl. the fp32 wgrad arena)
  size_t garena_off = 0;                          // FLOAT index from the workspace base, inside the backward block
  size_t fin_dw_acc = 0;                          // acc_t index: S * d(final conv weight [3][128], bias [3]) at +384
  Buf small_table;                                // SmallGradDesc[]: fixed-point small-layer gradients -> flat gradients
  std::vector<SmallGradDesc> small_host;
  Buf warena;                                     // packed 16-bit weights
  Buf pack_table, bnrun_table, errword, gscale, stepstate;
  // DSR_DETERMINISTIC=1: the split-K weight-gradient kernels store per-split partials here (plain stores) and a
  // second kernel adds them in split order, instead of fp32 atomics whose order varies from run to run.  With the
  // fixed-point statistics (dsr_acc.cuh) two runs of one binary are then bit-identical.
  int det = 0;
  Buf wg_part;
  Buf g_final;                                    // unused placeholder (level 0 g_u2a is the final-conv gradient)
  std::vector<PackDesc> pack_host;
  std::vector<BnRunDesc> bnrun_host;
  std::vector<TensorInfo> tensors;
  char* base = nullptr;
  int num_sms = 148;
  int launches = 0;
  int debug_conv = 0;
  int use_halo = 1;              // 3x3 stride-1 layers: halo-tile weight-stationary kernel (0: generic kernel)
  // weight-gradient kernels run on a second stream (they only feed the final unpack): fork after the layer's
  // dR is written, join before unpack_wgrad.  Fills the SMs that the latency-bound low-resolution levels leave idle.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int use_side = 1;
  // CUDA graph of one whole DIP iteration (dsr_dip_run): every launch argument is iteration-independent, the
  // iteration counter lives on the device (stepstate)
  cudaGraphExec_t graph_exec = nullptr;
  dsr_step_buffers_t graph_bufs{};
  const dsr_downsampler* graph_ds = nullptr;
  unsigned long long graph_ds_serial = 0;
  float graph_lr = 0.f, graph_sigma = 0.f;
  unsigned long long graph_seed = 0;
  int use_graph = 1;
  // small cache of captured forward / backward passes for the separate-call path (closure protocol): key = the
  // pointer arguments; an entry is captured on its SECOND use (the first run is eager and sets function attributes)
  struct PassGraph { int kind; const void* a0; const void* a1; const void* a2; const void* a3; int uses; cudaGraphExec_t exec; };
  std::vector<PassGraph> pass_graphs;
  int fuse_ubstats = 1;          // BN-backward sums of the layer consuming ddeep accumulated by the upcat backward pass C
  int fuse_skip = 1;             // skip-branch 1x1 conv of level i+1 fused into level i's last BN + LeakyReLU pass
  int lowres_upcat = 1;          // upsample + concat + BN(132): statistics and backward in the low-resolution domain
  int in_step = 0;               // inside enqueue_step (whole-iteration graph): no nested pass graphs
  // set by enqueue_step around the forward pass: z = pz_saved + psigma N(0,1) is drawn by the input packing pass
  const float* pz_saved = nullptr;
  float psigma = 0.f;
  unsigned long long pseed = 0;
  int fuse_top = 1;              // level 0: BN/LeakyReLU of the last decoder conv fused with the final conv (fwd + bwd)
  bool bound = false, have_forward = false;
  // optional per-launch timing of the tensor-core kernels (bench.py roofline): class 0 = halo-tile conv (3x3
  // stride-1 fprop + dgrad), 1 = wgrad, 2 = generic implicit-GEMM conv (stride-2 layers, 32-channel input), 3 = halo-tile
  // conv on 1x1 layers (memory-bound)
  struct ProfRec { int cls; cudaEvent_t a, b; double flops; int slot; };
  Buf prof_slots;                // kProfSlots x {min start, max end} %globaltimer stamps written by the kernels themselves
  std::vector<ProfRec> prof;
  struct AllRec { const char* what; cudaEvent_t a, b; };
  std::vector<AllRec> allprof;
  int profile = 0;
};

namespace dsr {
namespace {

struct Bump {
  size_t off = 0;
  void take(Buf& b, size_t bytes) {
    off = align_up(off, 1024);
    b.off = off;
    b.bytes = bytes;
    off += bytes;
  }
  size_t take_acc(size_t n) {      // n fixed-point accumulators (acc_t); returns an acc_t index
    off = align_up(off, 1024);
    const size_t r = off / sizeof(acc_t);
    off += n * sizeof(acc_t);
    return r;
  }
};

void add_param(dsr_plan* p, const std::string& name, long long& off_out, std::initializer_list<int> shape) {
  ParamInfo pi;
  pi.name = name;
  pi.off = p->nparam;
  pi.ndim = static_cast<int>(shape.size());
  long long n = 1;
  int i = 0;
  for (int s : shape) { pi.shape[i++] = s; n *= s; }
  for (; i < 4; ++i) pi.shape[i] = 1;
  off_out = pi.off;
  p->nparam += n;
  p->params.push_back(pi);
}
void add_conv_params(dsr_plan* p, const std::string& name, long long& w, long long& b, int co, int ci, int k) {
  add_param(p, name + ".weight", w, {co, ci, k, k});
  add_param(p, name + ".bias", b, {co});
}
void add_bn(dsr_plan* p, const std::string& name, long long& g, long long& be, long long& bn, int C) {
  add_param(p, name + ".weight", g, {C});
  add_param(p, name + ".bias", be, {C});
  BnInfo bi;
  bi.name = name;
  bi.off = p->nbn;
  bi.C = C;
  bn = bi.off;
  p->nbn += 2 * C;
  p->bns.push_back(bi);
}

// state_dict naming: models/DIP/utils.py:5-8 (children numbered from 1), skip.py:41-92.
void name_level(dsr_plan* p, int i) {
  Level& L = p->lv[i];
  std::string P;
  for (int j = 0; j < i; ++j) P += "1.1.7.";
  // models/DIP/utils.py:103-104: a conv is Sequential(ReflectionPad2d, Conv2d) -> "....1.weight"; with pad='zero'
  // there is no padder module and the Conv2d is child 0
  const std::string ci = p->pad_zero ? ".0" : ".1";
  add_conv_params(p, P + "1.0.1" + ci, L.skip_w, L.skip_b, kNS, L.Cin, 1);
  add_bn(p, P + "1.0.2", L.skip_g, L.skip_be, L.skip_bn, kNS);
  add_conv_params(p, P + "1.1.1" + ci, L.d1.w_off, L.d1.b_off, kNC, L.Cin, 3);
  add_bn(p, P + "1.1.2", L.d1.g_off, L.d1.be_off, L.d1.bn_off, kNC);
  add_conv_params(p, P + "1.1.4" + ci, L.d2.w_off, L.d2.b_off, kNC, kNC, 3);
  add_bn(p, P + "1.1.5", L.d2.g_off, L.d2.be_off, L.d2.bn_off, kNC);
  if (i + 1 < p->num_scales) name_level(p, i + 1);
  add_bn(p, P + "2", L.cat_g, L.cat_be, L.cat_bn, kNS + kNC);
  add_conv_params(p, P + "3" + ci, L.u1.w_off, L.u1.b_off, kNC, kNS + kNC, 3);
  add_bn(p, P + "4", L.u1.g_off, L.u1.be_off, L.u1.bn_off, kNC);
  add_conv_params(p, P + "6" + ci, L.u2.w_off, L.u2.b_off, kNC, kNC, 1);
  add_bn(p, P + "7", L.u2.g_off, L.u2.be_off, L.u2.bn_off, kNC);
}

void choose_tile(int out_w, int& tw, int& th) {
  tw = 16;
  while (tw > 1 && tw / 2 >= out_w) tw /= 2;
  th = 128 / tw;
}

void setup_conv(dsr_plan* p, Bump& ws, Bump& accf, Bump& accb, ConvLayer& c, const char* tag, int cin, int cin_pad,
                int k, int stride, int inH, int inW, const Buf* in_pad, bool need_dgrad, int act_halo, int perm,
                size_t& warena_elems, size_t& garena_floats) {
  c.tag = tag;
  c.cin = cin;
  c.cin_pad = cin_pad;
  c.k = k;
  c.stride = stride;
  c.inH = inH;
  c.inW = inW;
  c.outH = (stride == 1) ? inH : (inH + 1) / 2;
  c.outW = (stride == 1) ? inW : (inW + 1) / 2;
  c.in_pad = in_pad;
  c.need_dgrad = need_dgrad;
  c.n_rows = need_dgrad ? cin_pad : 0;
  c.act_halo = act_halo;
  ws.take(c.raw, static_cast<size_t>(c.outH) * c.outW * kNC * 2);
  // + one padded row and one pixel of zeros: the parity-split (stride-2) TMA view of an odd-sized grid addresses them
  ws.take(c.act, (static_cast<size_t>(c.outH + 3) * (c.outW + 2) + 1) * kNC * 2);
  ws.take(c.dr, static_cast<size_t>(c.outH + 2) * (c.outW + 2) * kNC * 2);
  if (need_dgrad) ws.take(c.gin, static_cast<size_t>(inH + 2) * (inW + 2) * c.n_rows * 2);
  c.stats_off = accf.take_acc(2 * kNC);
  c.bstats_off = accb.take_acc(2 * kNC);
  const int taps = k * k;
  c.pack.w_off = c.w_off;
  c.pack.cout = kNC;
  c.pack.cin = cin;
  c.pack.k = k;
  c.pack.cin_pad = cin_pad;
  c.pack.n_rows = c.n_rows;
  c.pack.perm = perm;
  warena_elems = align_up(warena_elems, 512);
  c.pack.f_off = static_cast<long long>(warena_elems);
  warena_elems += static_cast<size_t>(taps) * kNC * cin_pad;
  warena_elems = align_up(warena_elems, 512);
  c.pack.d_off = static_cast<long long>(warena_elems);
  warena_elems += static_cast<size_t>(taps) * c.n_rows * kNC;
  garena_floats = align_up(garena_floats, 256);
  c.pack.g_off = static_cast<long long>(garena_floats);
  garena_floats += static_cast<size_t>(taps) * kNC * cin_pad;
  (void)p;
}

void push_kblocks(ConvGemmParams& g, int c_total, int px, int dx, int py, int dy, int b_row) {
  int c = 0;
  while (c < c_total) {
    KBlk kb{};
    const int wide = (c_total - c) >= 64 ? 1 : 0;
    kb.a_c = static_cast<int16_t>(c);
    kb.a_px = static_cast<int8_t>(px);
    kb.a_dx = static_cast<int8_t>(dx);
    kb.a_py = static_cast<int8_t>(py);
    kb.a_dy = static_cast<int8_t>(dy);
    kb.wide = static_cast<int8_t>(wide);
    kb.b_k = static_cast<int16_t>(c);
    kb.b_row = static_cast<int16_t>(b_row);
    g.kb[g.nkb++] = kb;
    c += wide ? 64 : 16;
  }
}

// Builds the launch descriptions of one conv layer (needs resolved pointers).
int build_conv(dsr_plan* p, ConvLayer& c) {
  __half* warena = static_cast<__half*>(p->warena.ptr);
  acc_t* acc = reinterpret_cast<acc_t*>(p->base);
  int rc;
  // ---------------- fprop ----------------
  {
    ConvGemmParams& g = c.fprop;
    memset(&g, 0, sizeof(g));
    const int Wp = c.inW + 2, Hp = c.inH + 2;
    choose_tile(c.outW, g.tw, g.th);
    if ((rc = make_act_map(&g.a16, c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, c.stride, 16, g.tw, g.th))) return rc;
    if (c.cin_pad >= 64) {
      if ((rc = make_act_map(&g.a64, c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, c.stride, 64, g.tw, g.th))) return rc;
    } else {
      g.a64 = g.a16;
    }
    const void* wf = warena + c.pack.f_off;
    const int rows = c.k * c.k * kNC;
    if ((rc = make_wgt_map(&g.b16, wf, c.cin_pad, rows, 16, kNC))) return rc;
    if (c.cin_pad >= 64) {
      if ((rc = make_wgt_map(&g.b64, wf, c.cin_pad, rows, 64, kNC))) return rc;
    } else {
      g.b64 = g.b16;
    }
    for (int ky = 0; ky < c.k; ++ky)
      for (int kx = 0; kx < c.k; ++kx) {
        int px = 0, dx = 1, py = 0, dy = 1;        // 1x1: interior pixel
        if (c.k == 3) {
          if (c.stride == 1) { dx = kx; dy = ky; } else { px = kx & 1; dx = kx >> 1; py = ky & 1; dy = ky >> 1; }
        }
        push_kblocks(g, c.cin_pad, px, dx, py, dy, (ky * c.k + kx) * kNC);
      }
    if (g.nkb > kMaxKBlocks) return -20;
    g.tiles_x = (c.outW + g.tw - 1) / g.tw;
    g.tiles_y = (c.outH + g.th - 1) / g.th;
    g.out_h = c.outH;
    g.out_w = c.outW;
    g.out_sy = static_cast<long long>(c.outW) * kNC;
    g.out_sx = kNC;
    g.out = c.raw.ptr;
    g.n_mma = kNC;
    g.n_store = kNC;
    g.out_bf16 = 0;
    g.stats = acc + c.stats_off;
    g.idesc = make_idesc_f16(128, kNC, FMT_F16, FMT_F16, 0, 0);
    g.err = static_cast<int*>(p->errword.ptr);
    c.ref_in = ActRef{static_cast<const uint16_t*>(c.in_pad->ptr), 0, c.cin_pad, Wp, Hp, c.stride};
    c.ref_wf = WgtRef{reinterpret_cast<const uint16_t*>(wf), 0, c.cin_pad};
  }
  // ---------------- dgrad ----------------
  c.ndgrad = 0;
  c.ref_dr = ActRef{static_cast<const uint16_t*>(c.dr.ptr), 0, kNC, c.outW + 2, c.outH + 2, 1};
  if (c.need_dgrad) {
    const int oWp = c.outW + 2, oHp = c.outH + 2;     // dR grid
    const int iWp = c.inW + 2, iHp = c.inH + 2;       // output (input-gradient) grid
    const void* wd = warena + c.pack.d_off;
    const int rows = c.k * c.k * c.n_rows;
    c.ref_wd = WgtRef{reinterpret_cast<const uint16_t*>(wd), 0, kNC};
    const int nclass = (c.stride == 1) ? 1 : 4;
    for (int cls = 0; cls < nclass; ++cls) {
      ConvGemmParams& g = c.dgrad[c.ndgrad++];
      memset(&g, 0, sizeof(g));
      const int ry = (c.stride == 1) ? 0 : (cls >> 1), rx = (c.stride == 1) ? 0 : (cls & 1);
      const int step = c.stride;
      g.out_h = (iHp - ry + step - 1) / step;
      g.out_w = (iWp - rx + step - 1) / step;
      choose_tile(g.out_w, g.tw, g.th);
      if ((rc = make_act_map(&g.a64, c.dr.ptr, 1, kNC, oWp, oHp, 1, 64, g.tw, g.th))) return rc;
      g.a16 = g.a64;
      if ((rc = make_wgt_map(&g.b64, wd, kNC, rows, 64, c.n_rows))) return rc;
      g.b16 = g.b64;
      for (int ky = 0; ky < c.k; ++ky)
        for (int kx = 0; kx < c.k; ++kx) {
          int dx, dy;
          if (c.k == 1) {
            dx = 0; dy = 0;
          } else if (c.stride == 1) {
            dx = 1 - kx; dy = 1 - ky;
          } else {
            // padded input position q = 2 t + r receives output o = (q - k) / 2 when q - k is even
            if (((ry - ky) & 1) || ((rx - kx) & 1)) continue;
            dy = (ry - ky) / 2 + 1;      // (q - k)/2 + 1 = t + (r - k)/2 + 1   [(r-k) even, in {-2, 0}]
            dx = (rx - kx) / 2 + 1;
          }
          push_kblocks(g, kNC, 0, dx, 0, dy, (ky * c.k + kx) * c.n_rows);
        }
      g.tiles_x = (g.out_w + g.tw - 1) / g.tw;
      g.tiles_y = (g.out_h + g.th - 1) / g.th;
      g.out_sy = static_cast<long long>(iWp) * c.n_rows * step;
      g.out_sx = static_cast<long long>(c.n_rows) * step;
      g.out = static_cast<uint16_t*>(c.gin.ptr) + (static_cast<long long>(ry) * iWp + rx) * c.n_rows;
      g.n_mma = c.n_rows;
      g.n_store = c.n_rows;
      g.out_bf16 = 0;
      g.stats = nullptr;
      g.idesc = make_idesc_f16(128, c.n_rows, FMT_F16, FMT_F16, 0, 0);
      g.err = static_cast<int*>(p->errword.ptr);
    }
  }
  // stride-2 dgrad: merge the parity classes into one launch when they share the tile shape
  c.has_merged = false;
  if (c.ndgrad == 4) {
    bool same = true;
    int nkb = 0;
    for (int i = 0; i < 4; ++i) {
      same = same && c.dgrad[i].tw == c.dgrad[0].tw && c.dgrad[i].th == c.dgrad[0].th;
      nkb += c.dgrad[i].nkb;
    }
    if (same && nkb <= kMaxKBlocks) {
      ConvGemmParams& m = c.dgrad_merged;
      m = c.dgrad[0];
      m.ncls = 4;
      m.nkb = 0;
      m.tiles_x = 0;
      m.tiles_y = 0;
      int tile0 = 0;
      for (int i = 0; i < 4; ++i) {
        const ConvGemmParams& g = c.dgrad[i];
        if (g.tiles_x > m.tiles_x) m.tiles_x = g.tiles_x;
        if (g.tiles_y > m.tiles_y) m.tiles_y = g.tiles_y;
      }
      for (int i = 0; i < 4; ++i) {
        const ConvGemmParams& g = c.dgrad[i];
        m.cls_kb0[i] = m.nkb;
        m.cls_nkb[i] = g.nkb;
        for (int k = 0; k < g.nkb; ++k) m.kb[m.nkb++] = g.kb[k];
        m.cls_tile0[i] = tile0;
        tile0 += m.tiles_x * m.tiles_y;          // every class enumerates the common (max) tile grid; extras are masked
        m.cls_out_h[i] = g.out_h;
        m.cls_out_w[i] = g.out_w;
        m.cls_out_off[i] = static_cast<long long>(static_cast<const uint16_t*>(g.out) - static_cast<const uint16_t*>(c.gin.ptr));
      }
      m.cls_tile0[4] = tile0;
      m.out = c.gin.ptr;
      c.has_merged = true;
    }
  }
  // ---------------- wgrad ----------------
  {
    WgradParams& g = c.wgrad;
    memset(&g, 0, sizeof(g));
    g.pw = 8;
    while (g.pw > 1 && g.pw / 2 >= c.outW) g.pw /= 2;
    g.ph = kWgPix / g.pw;
    if ((rc = make_act_map(&g.a64, c.dr.ptr, 1, kNC, c.outW + 2, c.outH + 2, 1, 64, g.pw, g.ph))) return rc;
    const int Wp = c.inW + 2, Hp = c.inH + 2;
    if ((rc = make_act_map(&g.b16, c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, c.stride, 16, g.pw, g.ph))) return rc;
    if (c.cin_pad >= 64) {
      if ((rc = make_act_map(&g.b64, c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, c.stride, 64, g.pw, g.ph))) return rc;
    } else {
      g.b64 = g.b16;
    }
    g.n64 = (c.cin_pad >= 64) ? 2 : 0;
    g.n16 = (c.cin_pad - g.n64 * 64) / 16;
    g.c16_base = g.n64 * 64;
    if (g.n16 > 2 || (g.n64 == 2 && g.n16 > 1)) return -21;
    if (c.k == 1) {
      g.ngroups = 1;
      g.ntaps = 1;
      g.taps[0][0] = WgTap{0, 1, 0, 1, 0, 0};
    } else {
      g.ngroups = 3;
      g.ntaps = 3;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          WgTap t{};
          if (c.stride == 1) {
            t.px = 0; t.dx = static_cast<int8_t>(kx); t.py = 0; t.dy = static_cast<int8_t>(ky);
          } else {
            t.px = static_cast<int8_t>(kx & 1); t.dx = static_cast<int8_t>(kx >> 1);
            t.py = static_cast<int8_t>(ky & 1); t.dy = static_cast<int8_t>(ky >> 1);
          }
          t.w_tap = static_cast<int16_t>(ky * 3 + kx);
          g.taps[ky][kx] = t;
        }
    }
    g.pb_x = (c.outW + g.pw - 1) / g.pw;
    g.pb_y = (c.outH + g.ph - 1) / g.ph;
    const int npb = g.pb_x * g.pb_y;
    // split-K over pixel ranges: every CTA ends with a full-size fp32 atomic epilogue (128 x up to 432 values), so
    // small layers get few CTAs (>= kWgMinBlocks pixel blocks each) -- less atomic traffic, and SMs left free for
    // the main stream's conv kernels that run beside the weight-gradient stream
    int nsplit = p->num_sms / g.ngroups;
    const int kWgMinBlocks = getenv("DSR_WG_MINBLK") ? atoi(getenv("DSR_WG_MINBLK")) : 8;   // A/B: 16 -> 635.9, 8 -> 639.6, 4 -> 633.7, 2 -> 628.2 it/s
    if (nsplit > npb / kWgMinBlocks) nsplit = npb / kWgMinBlocks;
    if (nsplit > npb) nsplit = npb;
    if (nsplit < 1) nsplit = 1;
    g.ldw = c.cin_pad;
    g.dw = reinterpret_cast<float*>(p->base) + p->garena_off + c.pack.g_off;
    g.part_stride = static_cast<long long>(c.k) * c.k * kNC * c.cin_pad;
    g.part = nullptr;
    if (p->det) {
      g.part = static_cast<float*>(p->wg_part.ptr);
      const long long cap = static_cast<long long>(p->wg_part.bytes / 4) / g.part_stride;
      if (nsplit > cap) nsplit = static_cast<int>(cap);
    }
    g.nsplit = nsplit;
    g.idesc64 = make_idesc_f16(128, 128, FMT_F16, FMT_F16, 1, 1);
    g.idesc16 = make_idesc_f16(128, 16 * (g.n16 > 0 ? g.n16 : 1), FMT_F16, FMT_F16, 1, 1);   // n16 chunks, LBO apart
    g.err = static_cast<int*>(p->errword.ptr);
  }
  // ---------------- halo-row wgrad ----------------
  c.has_hwgrad = (c.outW >= 8 && c.outH >= 8 && !getenv("DSR_WGRAD_OLD"));
  if (c.has_hwgrad) {
    WgHaloParams& g = c.hwgrad;
    memset(&g, 0, sizeof(g));
    const WgradParams& o = c.wgrad;
    if ((rc = make_act_map(&g.a64, c.dr.ptr, 1, kNC, c.outW + 2, c.outH + 2, 1, 64, 8, 8))) return rc;
    const int Wp = c.inW + 2, Hp = c.inH + 2;
    g.n64 = o.n64; g.n16 = o.n16; g.c16_base = o.c16_base; g.ldw = o.ldw; g.dw = o.dw;
    g.idesc_base = make_idesc_f16(128, 0, FMT_F16, FMT_F16, 1, 1);
    g.err = o.err;
    g.pb_x = (c.outW + 7) / 8;
    g.pb_y = (c.outH + 7) / 8;
    g.merge_narrow = (g.n16 <= 1) ? 1 : 0;
    int widths[2] = {8, 8};
    // taps of one tap row, as runs of consecutive shifts: {box, shift0, r, kx of the run's taps...}
    struct RunDef { int box, shift0, r, kx[3]; };
    RunDef defs[3];
    int ndefs = 0;
    if (c.k == 1) {
      g.ngroups = 1; g.nbox = 1;
      g.box[0][0] = WgBox{0, 0, 1, 1, 8, 0};
      defs[ndefs++] = RunDef{0, 0, 1, {0, 0, 0}};
    } else if (c.stride == 1) {
      g.ngroups = 3; g.nbox = 1;
      widths[0] = 10;
      for (int ky = 0; ky < 3; ++ky) g.box[ky][0] = WgBox{0, 0, 0, static_cast<int8_t>(ky), 10, 0};
      if (g.merge_narrow) {
        defs[ndefs++] = RunDef{0, 0, 3, {0, 1, 2}};
      } else {
        for (int kx = 0; kx < 3; ++kx) defs[ndefs++] = RunDef{0, kx, 1, {kx, 0, 0}};
      }
    } else {   // stride 2: parity-0 columns serve kx = 0 (shift 0) and kx = 2 (shift 1); parity-1 columns serve kx = 1
      g.ngroups = 3; g.nbox = 2;
      widths[0] = 9;
      widths[1] = 8;
      const int rows0 = 9 * 8;
      const int off1 = rows0 * (g.n64 * 128 + g.n16 * 32);
      for (int ky = 0; ky < 3; ++ky) {
        g.box[ky][0] = WgBox{0, static_cast<int8_t>(ky & 1), 0, static_cast<int8_t>(ky >> 1), 9, 0};
        g.box[ky][1] = WgBox{1, static_cast<int8_t>(ky & 1), 0, static_cast<int8_t>(ky >> 1), 8,
                             static_cast<int16_t>(((off1 + 1023) & ~1023) / 16)};
      }
      if (g.merge_narrow) {
        defs[ndefs++] = RunDef{0, 0, 2, {0, 2, 0}};
        defs[ndefs++] = RunDef{1, 0, 1, {1, 0, 0}};
      } else {
        defs[ndefs++] = RunDef{0, 0, 1, {0, 0, 0}};
        defs[ndefs++] = RunDef{1, 0, 1, {1, 0, 0}};
        defs[ndefs++] = RunDef{0, 1, 1, {2, 0, 0}};
      }
    }
    g.nruns = ndefs;
    // TMEM column layout: per run, per 64-channel chunk: r x 64 columns (tap-major); then the narrow parts
    for (int grp = 0; grp < g.ngroups; ++grp) {
      int col = 0;
      for (int r = 0; r < ndefs; ++r) {
        g.runs[grp][r] = WgRun{static_cast<int8_t>(defs[r].box), static_cast<int8_t>(defs[r].shift0),
                               static_cast<int8_t>(defs[r].r), 0, static_cast<int16_t>(col), 0};
        for (int ch = 0; ch < g.n64; ++ch)
          for (int t = 0; t < defs[r].r; ++t)
            for (int q = 0; q < 4; ++q) {
              g.cols[grp][col / 16] = WgCol{static_cast<int16_t>(c.k == 1 ? 0 : grp * 3 + defs[r].kx[t]),
                                            static_cast<int16_t>(ch * 64 + q * 16)};
              col += 16;
            }
      }
      for (int r = 0; r < ndefs; ++r) {
        g.runs[grp][r].col_narrow = static_cast<int16_t>(col);
        if (g.merge_narrow) {
          for (int t = 0; t < defs[r].r && g.n16; ++t) {
            g.cols[grp][col / 16] = WgCol{static_cast<int16_t>(c.k == 1 ? 0 : grp * 3 + defs[r].kx[t]),
                                          static_cast<int16_t>(g.c16_base)};
            col += 16;
          }
        } else {
          for (int q = 0; q < g.n16; ++q) {
            g.cols[grp][col / 16] = WgCol{static_cast<int16_t>(c.k == 1 ? 0 : grp * 3 + defs[r].kx[0]),
                                          static_cast<int16_t>(g.c16_base + q * 16)};
            col += 16;
          }
        }
      }
      g.ncolchunks = col / 16;
      if (col > 512) return -22;
    }
    for (int b = 0; b < g.nbox; ++b) {
      if ((rc = make_act_map(&g.b16[b], c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, c.stride, 16, widths[b], 8))) return rc;
      g.b64[b] = g.b16[b];
      if (c.cin_pad >= 64 &&
          (rc = make_act_map(&g.b64[b], c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, c.stride, 64, widths[b], 8)))
        return rc;
    }
    if (g.nbox == 1) { g.b64[1] = g.b64[0]; g.b16[1] = g.b16[0]; }
    // 3 x 3 stride 1: tap-row CTAs as a cluster around one multicast 10 x 10 halo (maps in slot 1)
    g.cl3 = (c.k == 3 && c.stride == 1 && g.merge_narrow && g.n64 == 2 && !getenv("DSR_WG_NO_CL3")) ? 1 : 0;
    if (g.cl3) {
      if ((rc = make_act_map(&g.b64[1], c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, 1, 64, 10, 10))) return rc;
      if (g.n16 && (rc = make_act_map(&g.b16[1], c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, 1, 16, 10, 10))) return rc;
    }
    const int npb = g.pb_x * g.pb_y;
    // split-K over pixel ranges: every CTA ends with a full-size fp32 atomic epilogue (128 x up to 432 values), so
    // small layers get few CTAs (>= kWgMinBlocks pixel blocks each) -- less atomic traffic, and SMs left free for
    // the main stream's conv kernels that run beside the weight-gradient stream
    int nsplit = p->num_sms / g.ngroups;
    const int kWgMinBlocks = getenv("DSR_WG_MINBLK") ? atoi(getenv("DSR_WG_MINBLK")) : 8;   // A/B: 16 -> 635.9, 8 -> 639.6, 4 -> 633.7, 2 -> 628.2 it/s
    if (nsplit > npb / kWgMinBlocks) nsplit = npb / kWgMinBlocks;
    if (nsplit > npb) nsplit = npb;
    if (nsplit < 1) nsplit = 1;
    g.part_stride = o.part_stride;
    g.part = o.part;
    if (p->det) {
      const long long cap = static_cast<long long>(p->wg_part.bytes / 4) / g.part_stride;
      if (nsplit > cap) nsplit = static_cast<int>(cap);
    }
    if (g.cl3) {
      const int cap = getenv("DSR_WG_CL3_CAP") ? atoi(getenv("DSR_WG_CL3_CAP")) : wgrad_cluster_capacity();
      if (cap <= 0) g.cl3 = 0;
      else if (nsplit > cap) nsplit = cap;
    }
    g.nsplit = nsplit;
  }
  // ---------------- halo-tile variants ----------------
  c.has_halo = (c.stride == 1 && c.cin_pad >= 128);     // 3x3 and 1x1 stride-1 layers
  if (c.has_halo) {
    {
      HaloParams& h = c.hfprop;
      memset(&h, 0, sizeof(h));
      const int Wp = c.inW + 2, Hp = c.inH + 2;
      h.n_wide = 2;
      h.n_narrow = (c.cin_pad - 128) / 16;
      h.n_part = 64;
      h.parts = 2;
      h.nsplit = 1;
      if (!getenv("DSR_HALO_1CTA") && !getenv("DSR_NO_NSPLIT") && c.k == 3) {
        // small levels: spread the output channels over 2 or 4 clusters per tile pair (see HaloParams::nsplit)
        const int np = (((c.outW + kHaloTW - 1) / kHaloTW) * ((c.outH + kHaloTH - 1) / kHaloTH) + 1) / 2;
        for (int d = 4; d >= 2; d >>= 1)
          if (np * d <= p->num_sms / 2) { h.nsplit = d; break; }
        h.n_part = 64 / h.nsplit;
      }
      h.wide_slots = h.n_narrow ? 2 : 3;
      h.ntaps = c.k * c.k;
      h.halo_w = kHaloTW + (c.k - 1);
      h.halo_h = kHaloTH + (c.k - 1);
      if ((rc = make_act_map(&h.a64, c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, 1, 64, h.halo_w, h.halo_h))) return rc;
      h.a16 = h.a64;
      if (h.n_narrow && (rc = make_act_map(&h.a16, c.in_pad->ptr, 1, c.cin_pad, Wp, Hp, 1, 16, h.halo_w, h.halo_h)))
        return rc;
      const void* wf = warena + c.pack.f_off;
      if ((rc = make_wgt_map(&h.b64, wf, c.cin_pad, h.ntaps * kNC, 64, h.n_part))) return rc;
      h.b16 = h.b64;
      if (h.n_narrow && (rc = make_wgt_map(&h.b16, wf, c.cin_pad, h.ntaps * kNC, 16, h.n_part))) return rc;
      if (c.k == 3) {
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            h.taps[ky * 3 + kx] = HaloTap{static_cast<int8_t>(ky), static_cast<int8_t>(kx),
                                          static_cast<int16_t>((ky * 3 + kx) * kNC)};
      } else {
        h.taps[0] = HaloTap{0, 0, 0};     // 1x1: the box starts at the interior pixel itself (origin +1 below)
      }
      h.org_x = (c.k == 3) ? 0 : 1;
      h.org_y = (c.k == 3) ? 0 : 1;
      h.tiles_x = (c.outW + kHaloTW - 1) / kHaloTW;
      h.tiles_y = (c.outH + kHaloTH - 1) / kHaloTH;
      h.out_h = c.outH;
      h.out_w = c.outW;
      h.out_sy = static_cast<long long>(c.outW) * kNC;
      h.out_sx = kNC;
      h.out = c.raw.ptr;
      h.n_store = kNC;
      h.stats = acc + c.stats_off;
      h.stats_stride = kNC;
      h.idesc_wide = make_idesc_f16(128, h.n_part, FMT_F16, FMT_F16, 0, 0);
      h.idesc_narrow = h.idesc_wide;
      h.smem_bytes = halo_smem_bytes(h.n_part, h.n_wide, h.n_narrow, h.wide_slots, h.ntaps);
      h.dbg = getenv("DSR_HALO_DBG") ? atoi(getenv("DSR_HALO_DBG")) : 0;
      h.pair = getenv("DSR_HALO_1CTA") ? 0 : 1;
      if (h.pair) h.idesc_wide = h.idesc_narrow = make_idesc_f16(256, 2 * h.n_part, FMT_F16, FMT_F16, 0, 0);
      h.wide_slot_bytes = kHaloWideSlot;
      h.err = static_cast<int*>(p->errword.ptr);
    }
    if (c.need_dgrad) {
      HaloParams& h = c.hdgrad;
      memset(&h, 0, sizeof(h));
      const int oWp = c.outW + 2, oHp = c.outH + 2;     // dR grid
      const int iWp = c.inW + 2, iHp = c.inH + 2;       // output (input-gradient) grid
      h.n_wide = 2;
      h.n_narrow = 0;
      h.pair = getenv("DSR_HALO_1CTA") ? 0 : 1;
      h.parts = (h.pair || c.n_rows == 128) ? 2 : 3;
      h.n_part = c.n_rows / h.parts;                    // 64 / 72 (pairs) or 64 / 48
      h.nsplit = 1;
      if (h.pair && !getenv("DSR_NO_NSPLIT") && c.k == 3) {
        const int np = (((iWp + kHaloTW - 1) / kHaloTW) * ((iHp + kHaloTH - 1) / kHaloTH) + 1) / 2;
        if (c.n_rows == 144) {
          if (np * 3 <= p->num_sms / 2) h.nsplit = 3;   // 3 x 48 channels
        } else {
          for (int d = 4; d >= 2; d >>= 1)
            if (np * d <= p->num_sms / 2) { h.nsplit = d; break; }
        }
        h.n_part = c.n_rows / (2 * h.nsplit);
      }
      h.wide_slots = (h.n_part > 64) ? 2 : 3;           // 72-row weight slices leave room for two halo slots only
      h.ntaps = c.k * c.k;
      h.halo_w = kHaloTW + (c.k - 1);
      h.halo_h = kHaloTH + (c.k - 1);
      if ((rc = make_act_map(&h.a64, c.dr.ptr, 1, kNC, oWp, oHp, 1, 64, h.halo_w, h.halo_h))) return rc;
      h.a16 = h.a64;
      const void* wd = warena + c.pack.d_off;
      if ((rc = make_wgt_map(&h.b64, wd, kNC, h.ntaps * c.n_rows, 64, h.n_part))) return rc;
      h.b16 = h.b64;
      if (c.k == 3) {
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            h.taps[ky * 3 + kx] = HaloTap{static_cast<int8_t>(2 - ky), static_cast<int8_t>(2 - kx),
                                          static_cast<int16_t>((ky * 3 + kx) * c.n_rows)};
      } else {
        h.taps[0] = HaloTap{0, 0, 0};
      }
      h.org_x = (c.k == 3) ? -1 : 0;                    // 3x3: padded position q reads dR (padded) at q - k + 1
      h.org_y = (c.k == 3) ? -1 : 0;
      h.tiles_x = (iWp + kHaloTW - 1) / kHaloTW;
      h.tiles_y = (iHp + kHaloTH - 1) / kHaloTH;
      h.out_h = iHp;
      h.out_w = iWp;
      h.out_sy = static_cast<long long>(iWp) * c.n_rows;
      h.out_sx = c.n_rows;
      h.out = c.gin.ptr;
      h.n_store = c.n_rows;
      h.stats = nullptr;
      h.stats_stride = 0;
      h.idesc_wide = h.pair ? make_idesc_f16(256, 2 * h.n_part, FMT_F16, FMT_F16, 0, 0)
                            : make_idesc_f16(128, h.n_part, FMT_F16, FMT_F16, 0, 0);
      h.idesc_narrow = h.idesc_wide;
      h.smem_bytes = halo_smem_bytes(h.n_part, h.n_wide, h.n_narrow, h.wide_slots, h.ntaps);
      h.err = static_cast<int*>(p->errword.ptr);
    }
  }
  return 0;
}

BnRef conv_bn(const dsr_plan* p, const ConvLayer& c, const float* params) {
  BnRef b;
  b.stats = reinterpret_cast<const acc_t*>(p->base) + c.stats_off;
  b.gamma = params + c.g_off;
  b.beta = params + c.be_off;
  b.inv_n = 1.f / (static_cast<float>(c.outH) * static_cast<float>(c.outW));
  b.cstride = kNC;
  return b;
}
BnRef skip_bn(const dsr_plan* p, const Level& L, const float* params) {
  BnRef b;
  b.stats = reinterpret_cast<const acc_t*>(p->base) + L.skip_stats_off;
  b.gamma = params + L.skip_g;
  b.beta = params + L.skip_be;
  b.inv_n = 1.f / (static_cast<float>(L.H) * static_cast<float>(L.W));
  b.cstride = kNS;
  return b;
}

// With profile == 2 every launch is bracketed by CUDA events on the main stream and accumulated per call-site
// name (tools/step_table.py): warm, in-order timings of ALL kernels, unlike ncu's cold serialised replays.
#define DSR_TRY(expr)                                                          \
  do {                                                                         \
    cudaEvent_t ea__ = nullptr, eb__ = nullptr;                                \
    if (p->profile == 2) {                                                     \
      cudaEventCreate(&ea__);                                                  \
      cudaEventCreate(&eb__);                                                  \
      cudaEventRecord(ea__, s);                                                \
    }                                                                          \
    int rc__ = (expr);                                                         \
    ++p->launches;                                                             \
    if (p->profile == 2) {                                                     \
      cudaEventRecord(eb__, s);                                                \
      p->allprof.push_back(dsr_plan::AllRec{#expr, ea__, eb__});               \
    }                                                                          \
    if (rc__ != 0) return rc__;                                                \
  } while (0)

// Algorithmic FLOPs of one conv pass (fprop, dgrad or wgrad): 2 * M * N * K with the true channel count.
double conv_flops(const ConvLayer& c) {
  return 2.0 * c.outH * c.outW * kNC * static_cast<double>(c.cin) * c.k * c.k;
}
struct ProfScope {
  dsr_plan* p;
  cudaStream_t s;
  dsr_plan::ProfRec r;
  unsigned long long* slot = nullptr;      // device {min start, max end} pair the kernel stamps, or nullptr
  ProfScope(dsr_plan* p_, int cls, double flops, cudaStream_t s_) : p(p_), s(s_) {
    if (!p->profile) return;
    r.cls = cls;
    r.flops = flops;
    r.slot = static_cast<int>(p->prof.size()) < kProfSlots ? static_cast<int>(p->prof.size()) : -1;
    if (r.slot >= 0 && p->profile == 1) slot = static_cast<unsigned long long*>(p->prof_slots.ptr) + 2 * r.slot;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
  }
  ~ProfScope() {
    if (!p->profile) return;
    cudaEventRecord(r.b, s);
    p->prof.push_back(r);
  }
};
int run_fprop(dsr_plan* p, ConvLayer& c, cudaStream_t s) {
  if (p->debug_conv) return launch_conv_ref(c.fprop, c.ref_in, c.ref_wf, s);
  ProfScope ps(p, (c.has_halo && p->use_halo) ? (c.k == 1 ? 3 : 0) : 2, conv_flops(c), s);
  if (c.has_halo && p->use_halo) {
    if (ps.slot == nullptr) return launch_conv_halo(c.hfprop, p->num_sms, s);
    HaloParams hp = c.hfprop;
    hp.prof = ps.slot;
    return launch_conv_halo(hp, p->num_sms, s);
  }
  if (ps.slot == nullptr) return launch_conv_gemm(c.fprop, p->num_sms, s);
  ConvGemmParams gp = c.fprop;
  gp.prof = ps.slot;
  return launch_conv_gemm(gp, p->num_sms, s);
}
int run_dgrad(dsr_plan* p, ConvLayer& c, int i, cudaStream_t s) {
  if (p->debug_conv) return launch_conv_ref(c.dgrad[i], c.ref_dr, c.ref_wd, s);
  if (c.has_merged && !getenv("DSR_NO_MERGE")) {        // one launch covers all parity classes
    if (i > 0) return 0;
    ProfScope ps(p, 2, conv_flops(c), s);
    if (ps.slot == nullptr) return launch_conv_gemm(c.dgrad_merged, p->num_sms, s);
    ConvGemmParams gp = c.dgrad_merged;
    gp.prof = ps.slot;
    return launch_conv_gemm(gp, p->num_sms, s);
  }
  ProfScope ps(p, (c.has_halo && p->use_halo) ? (c.k == 1 ? 3 : 0) : 2, conv_flops(c) / c.ndgrad, s);
  if (c.has_halo && p->use_halo) {
    if (ps.slot == nullptr) return launch_conv_halo(c.hdgrad, p->num_sms, s);
    HaloParams hp = c.hdgrad;
    hp.prof = ps.slot;
    return launch_conv_halo(hp, p->num_sms, s);
  }
  if (ps.slot == nullptr) return launch_conv_gemm(c.dgrad[i], p->num_sms, s);
  ConvGemmParams gp = c.dgrad[i];
  gp.prof = ps.slot;
  return launch_conv_gemm(gp, p->num_sms, s);
}
int run_wgrad(dsr_plan* p, ConvLayer& c, cudaStream_t s) {
  if (p->debug_conv) return launch_wgrad_ref(c.wgrad, c.ref_dr, c.ref_in, s);
  {                  // diagnostics only (WRONG gradients): the main stream's timeline without the weight-gradient stream
    static const bool skip = getenv("DSR_EXP_SKIP_WGRAD") != nullptr;
    if (skip) return 0;
  }
  ProfScope ps(p, 1, conv_flops(c), s);
  if (c.has_hwgrad) {
    if (ps.slot == nullptr) return launch_wgrad_halo(c.hwgrad, s);
    WgHaloParams wp = c.hwgrad;
    wp.prof = ps.slot;
    return launch_wgrad_halo(wp, s);
  }
  if (ps.slot == nullptr) return launch_wgrad(c.wgrad, s);
  WgradParams wp = c.wgrad;
  wp.prof = ps.slot;
  return launch_wgrad(wp, s);
}

// `next` != nullptr: c is the last encoder conv of a level whose activation feeds level `next`'s skip branch; the
// skip 1x1 conv and its statistics are fused into the BN + LeakyReLU pass
int conv_bn_act(dsr_plan* p, ConvLayer& c, const float* params, cudaStream_t s, Level* next = nullptr) {
  DSR_TRY(run_fprop(p, c, s));
  if (next != nullptr && p->fuse_skip) {
    DSR_TRY(launch_bn_act(c.raw.ptr, conv_bn(p, c, params), c.act.ptr, c.outH, c.outW, p->pad_zero ? 0 : c.act_halo, s,
                          params + next->skip_w, static_cast<float*>(next->sraw.ptr),
                          reinterpret_cast<acc_t*>(p->base) + next->skip_stats_off));
  } else {
    DSR_TRY(launch_bn_act(c.raw.ptr, conv_bn(p, c, params), c.act.ptr, c.outH, c.outW, p->pad_zero ? 0 : c.act_halo, s));
  }
  return 0;
}

UpcatArgs upcat_args(dsr_plan* p, int i, const float* params) {
  Level& L = p->lv[i];
  const bool last = (i + 1 == p->num_scales);
  const Buf& deep = last ? L.d2.act : p->lv[i + 1].u2.act;
  UpcatArgs a{};
  a.deep = static_cast<const __half*>(deep.ptr) + (static_cast<long long>(L.w + 2) + 1) * kNC;
  a.deep_sy = static_cast<long long>(L.w + 2) * kNC;
  a.h = L.h;
  a.w = L.w;
  a.H = L.H;
  a.W = L.W;
  a.sraw = static_cast<const float*>(L.sraw.ptr);
  a.bn_skip = skip_bn(p, L, params);
  a.cat_stats = reinterpret_cast<acc_t*>(p->base) + L.cat_stats_off;
  a.cat_gamma = params + L.cat_g;
  a.cat_beta = params + L.cat_be;
  a.cat_pad = L.cat.ptr;
  // second quarter of the (otherwise backward-only) scratch tensor dup: [0, h w 128) holds t = U^T dc, then Q d
  a.qd = static_cast<__half*>(L.dup.ptr) + static_cast<size_t>(L.h) * L.w * kNC;
  a.nearest = p->nearest;
  a.pad_zero = p->pad_zero;
  return a;
}

int forward_level(dsr_plan* p, int i, const float* params, cudaStream_t s) {
  Level& L = p->lv[i];
  acc_t* acc = reinterpret_cast<acc_t*>(p->base);
  if (!p->fuse_skip || (i == 0 && !input_pack_fast(L.Cin, L.W)))   // otherwise produced by the pass that wrote this level's input
    DSR_TRY(launch_skip_conv(L.x_pad->ptr, L.Cin, params + L.skip_w, static_cast<float*>(L.sraw.ptr),
                             acc + L.skip_stats_off, L.H, L.W, s));
  int rc;
  if ((rc = conv_bn_act(p, L.d1, params, s))) return rc;
  if ((rc = conv_bn_act(p, L.d2, params, s, i + 1 < p->num_scales ? &p->lv[i + 1] : nullptr))) return rc;
  if (i + 1 < p->num_scales && (rc = forward_level(p, i + 1, params, s))) return rc;
  const UpcatArgs a = upcat_args(p, i, params);
  if (p->lowres_upcat) DSR_TRY(launch_upcat_stats_lowres(a, s));
  else DSR_TRY(launch_upcat_stats(a, s));
  DSR_TRY(launch_upcat_apply(a, s));
  if ((rc = conv_bn_act(p, L.u1, params, s))) return rc;
  if (i == 0 && p->fuse_top) {           // BN + LeakyReLU of u2 is fused with the final conv (dsr_net_forward)
    DSR_TRY(run_fprop(p, L.u2, s));
    return 0;
  }
  if ((rc = conv_bn_act(p, L.u2, params, s))) return rc;
  return 0;
}

// `next` != nullptr: c's activation also feeds level `next`'s skip branch, whose backward is done in the same passes
int bn_backward(dsr_plan* p, ConvLayer& c, const float* params, float* grads, const void* g, int gC, int fold,
                Level* next, cudaStream_t s) {
  BnBwdArgs a{};
  a.g = g;
  a.gC = gC;
  a.fold = p->pad_zero ? 0 : fold;
  a.bstats_raw = c.bstats_ready ? 1 : 0;
  if (next != nullptr) {
    a.dsy = static_cast<const float*>(next->dsy.ptr);
    a.sraw = static_cast<const float*>(next->sraw.ptr);
    a.bn_skip = skip_bn(p, *next, params);
    a.sbstats = reinterpret_cast<acc_t*>(p->base) + next->sbstats_off;
    a.wskip = params + next->skip_w;
    a.dwskip = reinterpret_cast<acc_t*>(p->base) + next->skip_dw_acc;
    a.dskip_gamma = grads + next->skip_g;
    a.dskip_beta = grads + next->skip_be;
    a.dsraw = static_cast<float*>(next->dsraw.ptr);
  }
  a.raw = c.raw.ptr;
  a.bn = conv_bn(p, c, params);
  a.bstats = reinterpret_cast<acc_t*>(p->base) + c.bstats_off;
  a.dr_pad = c.dr.ptr;
  a.dgamma = grads + c.g_off;
  a.dbeta = grads + c.be_off;
  a.gs = static_cast<float*>(p->gscale.ptr);
  a.H = c.outH;
  a.W = c.outW;
  if (!c.bstats_ready) DSR_TRY(launch_bn_bwd_stats(a, s));     // else: accumulated by the kernel that produced g
  c.bstats_ready = false;
  DSR_TRY(launch_bn_bwd_apply(a, s));
  return 0;
}

int conv_backward(dsr_plan* p, ConvLayer& c, cudaStream_t s) {
  if (p->side != nullptr && p->use_side && p->profile != 2) {
    cudaError_t e = cudaEventRecord(p->ev_fork, s);              // dR of this layer is complete on the main stream
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaStreamWaitEvent(p->side, p->ev_fork, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
    DSR_TRY(run_wgrad(p, c, p->side));
  } else {
    DSR_TRY(run_wgrad(p, c, s));
  }
  for (int i = 0; i < c.ndgrad; ++i) {
    if (i > 0 && c.has_merged && !p->debug_conv && !getenv("DSR_NO_MERGE")) break;   // one launch covered all classes
    DSR_TRY(run_dgrad(p, c, i, s));
  }
  return 0;
}

int backward_level(dsr_plan* p, int i, const float* params, float* grads, cudaStream_t s) {
  Level& L = p->lv[i];
  const bool last = (i + 1 == p->num_scales);
  acc_t* acc = reinterpret_cast<acc_t*>(p->base);
  int rc;
  // ---- decoder ----
  if (!(i == 0 && p->fuse_top) &&        // level 0: done by the fused top kernels (dsr_net_backward)
      (rc = bn_backward(p, L.u2, params, grads, L.g_u2a.ptr, kNC, 0, nullptr, s)))
    return rc;
  if ((rc = conv_backward(p, L.u2, s))) return rc;
  if ((rc = bn_backward(p, L.u1, params, grads, L.u2.gin.ptr, kNC, 0, nullptr, s))) return rc;
  if ((rc = conv_backward(p, L.u1, s))) return rc;
  UpcatBwdArgs ub{};
  ub.f = upcat_args(p, i, params);
  ub.gcat = L.u1.gin.ptr;
  ub.cbstats = acc + L.cbstats_off;
  ub.dup_pad = L.dup.ptr;
  ub.dsy = static_cast<float*>(L.dsy.ptr);
  ub.sbstats = acc + L.sbstats_off;
  ub.dcat_gamma = grads + L.cat_g;
  ub.dcat_beta = grads + L.cat_be;
  ub.gs = static_cast<const float*>(p->gscale.ptr);
  void* ddeep = last ? L.g_d2a.ptr : p->lv[i + 1].g_u2a.ptr;
  // the consumer's BN-backward sums ride on pass C where the launch saved outweighs the extra register pressure of the
  // element-wise pass (measured: a gain up to 128 x 128 low-resolution pixels, a loss at 256 x 256)
  if (p->lowres_upcat && p->fuse_ubstats && !p->debug_conv && L.h * L.w <= 128 * 128) {
    ConvLayer& cons = last ? L.d2 : p->lv[i + 1].u2;
    ub.cons_raw = cons.raw.ptr;
    ub.cons_bn = conv_bn(p, cons, params);
    ub.cons_bstats = acc + cons.bstats_off;
    cons.bstats_ready = true;
  }
  if (p->lowres_upcat) {
    DSR_TRY(launch_upcat_bwd_gather(ub, s));
    DSR_TRY(launch_upcat_bwd_apply_lowres(ub, ddeep, s));
  } else {
    DSR_TRY(launch_upcat_bwd_stats(ub, s));
    DSR_TRY(launch_upcat_bwd_apply(ub, s));
  }
  if (i == 0)       // levels >= 1: the skip branch's backward rides on the BN backward of the activation it reads
  DSR_TRY(launch_skip_bwd(static_cast<const float*>(L.dsy.ptr), static_cast<const float*>(L.sraw.ptr),
                          skip_bn(p, L, params), acc + L.sbstats_off, L.x_pad->ptr, L.Cin,
                          static_cast<float*>(L.dsraw.ptr), acc + L.skip_dw_acc, grads + L.skip_g, grads + L.skip_be,
                          static_cast<const float*>(p->gscale.ptr), L.H, L.W, s));
  if (!p->lowres_upcat) DSR_TRY(launch_upsample_bwd(L.dup.ptr, L.H, L.W, ddeep, L.h, L.w, s));
  if (!last && (rc = backward_level(p, i + 1, params, grads, s))) return rc;
  // ---- encoder ----
  if (last) {
    rc = bn_backward(p, L.d2, params, grads, L.g_d2a.ptr, kNC, 0, nullptr, s);
  } else {
    Level& N = p->lv[i + 1];
    rc = bn_backward(p, L.d2, params, grads, N.d1.gin.ptr, kNC, 1, &N, s);
  }
  if (rc) return rc;
  if ((rc = conv_backward(p, L.d2, s))) return rc;
  if ((rc = bn_backward(p, L.d1, params, grads, L.d2.gin.ptr, kNC, 1, nullptr, s))) return rc;
  if ((rc = conv_backward(p, L.d1, s))) return rc;
  return 0;
}

void reg_tensor(dsr_plan* p, const std::string& name, const Buf* b, int kind, int padded, int H, int W, int C) {
  p->tensors.push_back(TensorInfo{name, b, 0, kind, padded, H, W, C});
}
// acc_off: FLOAT index from the workspace base (fixed-point accumulators: 2 x their acc_t index)
void reg_acc(dsr_plan* p, const std::string& name, size_t acc_off, int H, int W, int C, int kind = 2) {
  p->tensors.push_back(TensorInfo{name, nullptr, acc_off, kind, 0, H, W, C});
}

}  // namespace
}  // namespace dsr

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int dsr_abi_version(void) { return 1; }

const char* dsr_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  switch (code) {
    case -1: return "bad argument";
    case -2: return "unsupported channel count";
    case -3: return "misaligned buffer";
    case -4: return "downsampler tile does not fit shared memory";
    case -5: return "unsupported network configuration";
    case -6: return "plan is not bound to a workspace";
    case -7: return "backward called without a preceding forward";
    case -8: return "workspace too small";
    case -20: return "too many K blocks";
    case -21: return "unsupported channel split";
    case -100: return "cuTensorMapEncodeTiled unavailable";
    default: return (code <= -200) ? "cuTensorMapEncodeTiled failed" : "unknown error";
  }
}

// ---- Lanczos -----------------------------------------------------------------------------------
int dsr_lanczos_kernel(int factor, int support, double* host_kernel, int capacity) {
  if (factor < 1 || (support != 2 && support != 3) || host_kernel == nullptr) return -1;
  std::vector<double> t;
  lanczos_taps(factor, support, t);
  const int k = static_cast<int>(t.size());
  if (capacity < k * k) return -1;
  // get_kernel normalises the 2-D table by its sum (utils/downsampler.py:133); the taps are
  // already normalised to sum 1 so the outer product sums to 1 up to rounding; renormalise the
  // product exactly as the reference does.
  double sum = 0.0;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      host_kernel[i * k + j] = t[i] * t[j];
      sum += host_kernel[i * k + j];
    }
  for (int i = 0; i < k * k; ++i) host_kernel[i] /= sum;
  return k;
}

static int ds_geometry(int factor, int support, int H, int W, int& k, int& pad, int& oh, int& ow) {
  if (factor < 1 || (support != 2 && support != 3) || H < 1 || W < 1) return -1;
  k = 2 * support * factor;
  pad = (k - factor) / 2;                       // utils/downsampler.py:56-59 (even kernel size)
  oh = (H + 2 * pad - k) / factor + 1;
  ow = (W + 2 * pad - k) / factor + 1;
  if (oh < 1 || ow < 1) return -1;
  return 0;
}

size_t dsr_downsampler_table_bytes(int factor, int support, int H, int W) {
  int k, pad, oh, ow;
  if (ds_geometry(factor, support, H, W, k, pad, oh, ow)) return 0;
  // generous bound: taps + per-row / per-column (index + up to k/f + 2 weights)
  const size_t nwmax = static_cast<size_t>(k) + 2;
  return 1024 + 64 + 4 * (static_cast<size_t>(k) + (H + W) * (1 + nwmax));
}

int dsr_downsampler_create(dsr_downsampler_t** out, int factor, int support, int H, int W, void* table_ws,
                           size_t table_bytes, void* stream) {
  if (out == nullptr || table_ws == nullptr) return -1;
  int k, pad, oh, ow;
  if (ds_geometry(factor, support, H, W, k, pad, oh, ow)) return -1;
  dsr_downsampler* d = new dsr_downsampler();
  {
    static unsigned long long next_serial = 0;
    d->serial = __atomic_add_fetch(&next_serial, 1ull, __ATOMIC_RELAXED);
  }
  d->factor = factor; d->support = support; d->H = H; d->W = W; d->oh = oh; d->ow = ow; d->k = k; d->pad = pad;
  std::vector<double> t64;
  lanczos_taps(factor, support, t64);
  // The reference casts the normalised 2-D float64 table to fp32 (utils/downsampler.py:48-50); the separable
  // evaluation uses fp32 1-D taps, whose products differ from the cast 2-D entries by <= 1 ulp each.
  d->taps.resize(k);
  for (int i = 0; i < k; ++i) d->taps[i] = static_cast<float>(t64[i]);
  ds_bwd_table(H, oh, factor, k, pad, d->taps, d->nw_y, d->by0, d->bwy);
  ds_bwd_table(W, ow, factor, k, pad, d->taps, d->nw_x, d->bx0, d->bwx);
  // the kernel uses one nw for both axes: widen the narrower table
  const int nw = d->nw_y > d->nw_x ? d->nw_y : d->nw_x;
  auto widen = [nw](std::vector<float>& w, int n, int old) {
    if (old == nw) return;
    std::vector<float> r(static_cast<size_t>(n) * nw, 0.f);
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < old; ++a) r[static_cast<size_t>(i) * nw + a] = w[static_cast<size_t>(i) * old + a];
    w.swap(r);
  };
  widen(d->bwy, H, d->nw_y);
  widen(d->bwx, W, d->nw_x);
  const size_t need = 4 * (static_cast<size_t>(k) + H + W + static_cast<size_t>(H + W) * nw) + 256 + 64;
  if (need > table_bytes) { delete d; return -8; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(table_ws);
  if (reinterpret_cast<uintptr_t>(base) & 15) { delete d; return -3; }
  cudaMemsetAsync(base, 0, 64, s);                 // loss accumulator + ticket of the fused-MSE kernel
  d->t.loss_acc = reinterpret_cast<acc_t*>(base);
  d->t.loss_ticket = reinterpret_cast<unsigned int*>(base + 16);
  size_t off = 64;
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* dst = base + off;
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    off = align_up(off + bytes, 16);
    return dst;
  };
  d->t.taps = static_cast<const float*>(put(d->taps.data(), 4 * static_cast<size_t>(k)));
  d->t.by0 = static_cast<const int*>(put(d->by0.data(), 4 * static_cast<size_t>(H)));
  d->t.bwy = static_cast<const float*>(put(d->bwy.data(), 4 * static_cast<size_t>(H) * nw));
  d->t.bx0 = static_cast<const int*>(put(d->bx0.data(), 4 * static_cast<size_t>(W)));
  d->t.bwx = static_cast<const float*>(put(d->bwx.data(), 4 * static_cast<size_t>(W) * nw));
  d->t.k = k; d->t.factor = factor; d->t.pad = pad; d->t.nw = nw;
  cudaError_t e = cudaStreamSynchronize(s);     // host vectors are pageable: make the upload complete here
  if (e != cudaSuccess) { delete d; return static_cast<int>(e); }
  *out = d;
  return 0;
}

void dsr_downsampler_destroy(dsr_downsampler_t* d) { delete d; }

int dsr_downsample_fwd(const dsr_downsampler_t* d, const float* x, float* y, int C, void* stream) {
  if (!d || !x || !y || C < 1) return -1;
  return launch_downsample_fwd(x, y, C, d->H, d->W, d->oh, d->ow, d->t, static_cast<cudaStream_t>(stream));
}
int dsr_downsample_bwd(const dsr_downsampler_t* d, const float* gy, float* gx, int C, void* stream) {
  if (!d || !gy || !gx || C < 1) return -1;
  return launch_downsample_bwd(gy, gx, C, d->H, d->W, d->oh, d->ow, d->t, static_cast<cudaStream_t>(stream));
}
int dsr_downsample_mse(const dsr_downsampler_t* d, const float* x, const float* target, float* y, float* gy,
                       float* loss, int C, void* stream) {
  if (!d || !x || !target || !y || !gy || !loss || C < 1) return -1;
  return launch_downsample_mse(x, target, y, gy, loss, C, d->H, d->W, d->oh, d->ow, d->t,
                               static_cast<cudaStream_t>(stream));
}

// ---- plan ---------------------------------------------------------------------------------------
int dsr_plan_create(dsr_plan_t** out, int H, int W, int input_depth, int num_scales, int n_out) {
  return dsr_plan_create_ex(out, H, W, input_depth, num_scales, n_out, 0);
}

int dsr_plan_create_ex(dsr_plan_t** out, int H, int W, int input_depth, int num_scales, int n_out, int flags) {
  if (out == nullptr) return -1;
  if (flags & ~(DSR_PLAN_PAD_ZERO | DSR_PLAN_UP_NEAREST)) return -5;
  if (num_scales < 1 || num_scales > 6 || n_out != 3) return -5;
  if (input_depth != 32 && input_depth != 128) return -5;
  {  // any H, W whose level sizes stay >= 2 (ReflectionPad2d(1) needs two pixels): stride-2 convs give ceil(n / 2),
     // the x2-upsampled deeper branch is centre-cropped to the skip branch (Concat, models/DIP/utils.py:26-36):
     // the size difference is 0 or 1, so the crop offset is 0 and only the last row / column is dropped
    int h = H, w = W;
    for (int i = 0; i < num_scales; ++i) {
      if (h < 2 || w < 2) return -5;
      h = (h + 1) / 2;
      w = (w + 1) / 2;
    }
    if (h < 2 || w < 2) return -5;
  }
  dsr_plan* p = new dsr_plan();
  p->H = H; p->W = W; p->input_depth = input_depth; p->num_scales = num_scales; p->n_out = n_out;
  p->pad_zero = (flags & DSR_PLAN_PAD_ZERO) ? 1 : 0;
  p->nearest = (flags & DSR_PLAN_UP_NEAREST) ? 1 : 0;
  p->lv.resize(num_scales);
  for (int i = 0, h = H, w = W; i < num_scales; ++i) {
    Level& L = p->lv[i];
    L.H = h; L.W = w; L.h = (h + 1) / 2; L.w = (w + 1) / 2;
    L.Cin = (i == 0) ? input_depth : kNC;
    h = L.h;
    w = L.w;
  }
  name_level(p, 0);
  add_conv_params(p, p->pad_zero ? "9.0" : "9.1", p->fin_w, p->fin_b, n_out, kNC, 1);

  Bump ws, accf, accb;
  size_t warena_elems = 0, garena_floats = 0;
  // Accumulator blocks are laid out first (their offsets are float indices from the workspace base).
  // Sizes are only known after the layer walk, so walk with separate bump allocators and place them after.
  for (int i = 0; i < num_scales; ++i) {
    Level& L = p->lv[i];
    const bool last = (i + 1 == num_scales);
    if (i == 0) {
      ws.take(L.xin, (static_cast<size_t>(L.H + 3) * (L.W + 2) + 1) * L.Cin * 2);
      L.x_pad = &L.xin;
    } else {
      L.x_pad = &p->lv[i - 1].d2.act;
    }
    ws.take(L.sraw, static_cast<size_t>(L.H) * L.W * kNS * 4);
    ws.take(L.cat, static_cast<size_t>(L.H + 2) * (L.W + 2) * kCat * 2);
    ws.take(L.g_u2a, static_cast<size_t>(L.H + 2) * (L.W + 2) * kNC * 2);
    ws.take(L.dup, static_cast<size_t>(L.H + 2) * (L.W + 2) * kNC * 2);
    ws.take(L.dsy, static_cast<size_t>(L.H) * L.W * kNS * 4);
    ws.take(L.dsraw, static_cast<size_t>(L.H) * L.W * kNS * 4);
    if (last) ws.take(L.g_d2a, static_cast<size_t>(L.h + 2) * (L.w + 2) * kNC * 2);
    L.skip_stats_off = accf.take_acc(2 * kNS);
    L.cat_stats_off = accf.take_acc(2 * kCat);
    L.sbstats_off = accb.take_acc(2 * kNS);
    L.cbstats_off = accb.take_acc(2 * kCat);
    L.skip_dw_acc = accb.take_acc(static_cast<size_t>(kNS) * L.Cin);
    setup_conv(p, ws, accf, accb, L.d1, "d1", L.Cin, L.Cin, 3, 2, L.H, L.W, L.x_pad, i > 0, 1, 0, warena_elems,
               garena_floats);
    setup_conv(p, ws, accf, accb, L.d2, "d2", kNC, kNC, 3, 1, L.h, L.w, &L.d1.act, true, 1, 0, warena_elems,
               garena_floats);
    setup_conv(p, ws, accf, accb, L.u1, "u1", kNS + kNC, kCat, 3, 1, L.H, L.W, &L.cat, true, 0, 1, warena_elems,
               garena_floats);
    setup_conv(p, ws, accf, accb, L.u2, "u2", kNC, kNC, 1, 1, L.H, L.W, &L.u1.act, true, 0, 0, warena_elems,
               garena_floats);
  }
  ws.take(p->warena, warena_elems * 2);
  p->det = (getenv("DSR_DETERMINISTIC") && atoi(getenv("DSR_DETERMINISTIC"))) ? 1 : 0;
  if (p->det) {
    // room for the per-split partials of the largest layer: up to 160 / groups splits of [taps][128][cin_pad] floats
    size_t need = 0;
    for (Level& L : p->lv)
      for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) {
        const size_t npb = static_cast<size_t>((c->outW + 7) / 8) * ((c->outH + 7) / 8);
        size_t splits = 160 / (c->k == 1 ? 1 : 3);
        if (splits > npb) splits = npb;
        const size_t b = splits * c->k * c->k * kNC * c->cin_pad * 4;
        if (b > need) need = b;
      }
    ws.take(p->wg_part, need);
  }
  ws.take(p->pack_table, sizeof(PackDesc) * 4 * num_scales);
  ws.take(p->bnrun_table, sizeof(BnRunDesc) * 6 * num_scales);
  ws.take(p->small_table, sizeof(SmallGradDesc) * (num_scales + 2));
  ws.take(p->prof_slots, sizeof(unsigned long long) * 2 * kProfSlots);
  ws.take(p->errword, 256);
  ws.take(p->gscale, 256);
  ws.take(p->stepstate, 256);
  // place the accumulator blocks
  ws.off = align_up(ws.off, 1024);
  p->fin_dw_acc = accb.take_acc(3 * kNC + 8);
  p->acc_fwd_off = ws.off;
  p->acc_fwd_bytes = align_up(accf.off, 1024);
  ws.off += p->acc_fwd_bytes;
  p->acc_bwd_off = ws.off;
  const size_t accb_bytes = align_up(accb.off, 1024);
  p->garena_off = (p->acc_bwd_off + accb_bytes) / 4;
  p->acc_bwd_bytes = accb_bytes + garena_floats * 4;
  ws.off += p->acc_bwd_bytes;
  p->ws_bytes = align_up(ws.off, 1024);
  // rebase accumulator indices to the workspace base
  const size_t fbase = p->acc_fwd_off / sizeof(acc_t), bbase = p->acc_bwd_off / sizeof(acc_t);
  p->fin_dw_acc += bbase;
  for (Level& L : p->lv) {
    L.skip_stats_off += fbase; L.cat_stats_off += fbase;
    L.sbstats_off += bbase; L.cbstats_off += bbase; L.skip_dw_acc += bbase;
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) { c->stats_off += fbase; c->bstats_off += bbase; }
  }
  // tables
  for (Level& L : p->lv)
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) {
      c->pack.w_off = c->w_off;
      p->pack_host.push_back(c->pack);
    }
  for (Level& L : p->lv) {
    auto add_run = [&](size_t stats_off, int cstride, int C, float n, long long bn_off, long long bias_off, int perm) {
      BnRunDesc d{};
      d.stats_off = static_cast<long long>(stats_off);
      d.cstride = cstride; d.C = C; d.n = n;
      d.rm_off = bn_off; d.rv_off = bn_off + C;
      d.bias_off = bias_off; d.perm = perm;
      p->bnrun_host.push_back(d);
    };
    add_run(L.skip_stats_off, kNS, kNS, static_cast<float>(L.H) * L.W, L.skip_bn, L.skip_b, 0);
    add_run(L.cat_stats_off, kCat, kNS + kNC, static_cast<float>(L.H) * L.W, L.cat_bn, -1, 1);
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2})
      add_run(c->stats_off, kNC, kNC, static_cast<float>(c->outH) * c->outW, c->bn_off, c->b_off, 0);
  }
  // small-layer gradients accumulated in fixed point by the element-wise kernels
  for (Level& L : p->lv)
    p->small_host.push_back(SmallGradDesc{static_cast<long long>(L.skip_dw_acc), L.skip_w, kNS * L.Cin, 0});
  p->small_host.push_back(SmallGradDesc{static_cast<long long>(p->fin_dw_acc), p->fin_w, n_out * kNC, 0});
  p->small_host.push_back(SmallGradDesc{static_cast<long long>(p->fin_dw_acc) + 3 * kNC, p->fin_b, n_out, 0});
  // introspection table
  for (int i = 0; i < num_scales; ++i) {
    Level& L = p->lv[i];
    const std::string P = "L" + std::to_string(i) + ".";
    if (i == 0) reg_acc(p, "gscale", p->gscale.off / 4, 1, 1, 8);
    reg_tensor(p, P + "x", L.x_pad, 0, 1, L.H, L.W, L.Cin);
    reg_tensor(p, P + "sraw", &L.sraw, 2, 0, L.H, L.W, kNS);
    reg_tensor(p, P + "cat", &L.cat, 0, 1, L.H, L.W, kCat);
    reg_tensor(p, P + "g_u2a", &L.g_u2a, 0, 1, L.H, L.W, kNC);
    reg_tensor(p, P + "dup", &L.dup, 0, 1, L.H, L.W, kNC);
    reg_tensor(p, P + "dsy", &L.dsy, 2, 0, L.H, L.W, kNS);
    reg_tensor(p, P + "dsraw", &L.dsraw, 2, 0, L.H, L.W, kNS);
    if (i + 1 == num_scales) reg_tensor(p, P + "g_d2a", &L.g_d2a, 0, 1, L.h, L.w, kNC);
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) {
      const std::string T = P + c->tag;
      reg_tensor(p, T + "_raw", &c->raw, 0, 0, c->outH, c->outW, kNC);
      reg_tensor(p, T + "_act", &c->act, 0, 1, c->outH, c->outW, kNC);
      reg_tensor(p, T + "_dr", &c->dr, 0, 1, c->outH, c->outW, kNC);
      if (c->need_dgrad) reg_tensor(p, T + "_gin", &c->gin, 0, 1, c->inH, c->inW, c->n_rows);
      reg_acc(p, T + "_stats", c->stats_off * 2, 1, 2, kNC, 3);      // kind 3: 64-bit fixed point, forward scale
      reg_acc(p, T + "_dw", p->garena_off + c->pack.g_off, c->k * c->k, kNC, c->cin_pad);
    }
  }
  *out = p;
  return 0;
}

void dsr_plan_destroy(dsr_plan_t* p) {
  if (p) dsr_plan_set_profile(p, 0);
  if (p && p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  if (p) for (auto& g : p->pass_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (p && p->side) {
    cudaStreamSynchronize(p->side);
    cudaEventDestroy(p->ev_fork);
    cudaEventDestroy(p->ev_join);
    cudaStreamDestroy(p->side);
  }
  delete p;
}

int dsr_plan_num_params(const dsr_plan_t* p) { return p ? static_cast<int>(p->params.size()) : -1; }
long long dsr_plan_param_numel(const dsr_plan_t* p) { return p ? p->nparam : -1; }
int dsr_plan_param_info(const dsr_plan_t* p, int idx, char* name, int name_cap, long long* offset, int* ndim,
                        int* shape4) {
  if (!p || idx < 0 || idx >= static_cast<int>(p->params.size())) return -1;
  const ParamInfo& pi = p->params[idx];
  if (name && name_cap > 0) snprintf(name, static_cast<size_t>(name_cap), "%s", pi.name.c_str());
  if (offset) *offset = pi.off;
  if (ndim) *ndim = pi.ndim;
  if (shape4) for (int i = 0; i < 4; ++i) shape4[i] = pi.shape[i];
  return 0;
}
int dsr_plan_num_bn(const dsr_plan_t* p) { return p ? static_cast<int>(p->bns.size()) : -1; }
long long dsr_plan_bn_numel(const dsr_plan_t* p) { return p ? p->nbn : -1; }
int dsr_plan_bn_info(const dsr_plan_t* p, int idx, char* name, int name_cap, long long* offset, int* channels) {
  if (!p || idx < 0 || idx >= static_cast<int>(p->bns.size())) return -1;
  const BnInfo& bi = p->bns[idx];
  if (name && name_cap > 0) snprintf(name, static_cast<size_t>(name_cap), "%s", bi.name.c_str());
  if (offset) *offset = bi.off;
  if (channels) *channels = bi.C;
  return 0;
}
size_t dsr_plan_workspace_bytes(const dsr_plan_t* p) { return p ? p->ws_bytes : 0; }

static void kstamp_arm();
int dsr_plan_bind(dsr_plan_t* p, void* workspace, size_t bytes, void* stream) {
  if (!p || !workspace) return -1;
  if (bytes < p->ws_bytes) return -8;
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return -3;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }     // captured pointers are stale
  for (auto& g : p->pass_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  p->pass_graphs.clear();
  p->base = static_cast<char*>(workspace);
  cudaError_t e = cudaMemsetAsync(workspace, 0, p->ws_bytes, s);   // zero halos of gradient tensors, pad channels
  if (e != cudaSuccess) return static_cast<int>(e);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (p->num_sms < 1) p->num_sms = 148;
  auto resolve = [p](Buf& b) { b.ptr = b.bytes ? p->base + b.off : nullptr; };
  for (Level& L : p->lv) {
    resolve(L.xin); resolve(L.sraw); resolve(L.cat); resolve(L.g_u2a); resolve(L.dup); resolve(L.dsy);
    resolve(L.dsraw); resolve(L.g_d2a);
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) { resolve(c->raw); resolve(c->act); resolve(c->dr); resolve(c->gin); }
  }
  resolve(p->warena); resolve(p->pack_table); resolve(p->bnrun_table); resolve(p->small_table); resolve(p->prof_slots); resolve(p->wg_part); resolve(p->errword); resolve(p->gscale); resolve(p->stepstate);
  e = cudaMemcpyAsync(p->pack_table.ptr, p->pack_host.data(), sizeof(PackDesc) * p->pack_host.size(),
                      cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemcpyAsync(p->bnrun_table.ptr, p->bnrun_host.data(), sizeof(BnRunDesc) * p->bnrun_host.size(),
                      cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemcpyAsync(p->small_table.ptr, p->small_host.data(), sizeof(SmallGradDesc) * p->small_host.size(),
                      cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  const float gs0[8] = {65536.f, 1.f / 65536.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // initial gradient scale 2^16
  e = cudaMemcpyAsync(p->gscale.ptr, gs0, sizeof(gs0), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return static_cast<int>(e);
  for (Level& L : p->lv)
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) {
      int rc = build_conv(p, *c);
      if (rc) return rc;
    }
  if (p->side == nullptr) {
    if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess) p->side = nullptr;
    if (p->side != nullptr) {
      cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
    }
  }
  kstamp_arm();
  p->use_side = getenv("DSR_NO_SIDE_STREAM") ? 0 : 1;
  p->use_graph = getenv("DSR_NO_GRAPH") ? 0 : 1;
  p->fuse_top = getenv("DSR_NO_FUSE_TOP") ? 0 : 1;
  p->lowres_upcat = (getenv("DSR_NO_LOWRES_UPCAT") && !p->nearest) ? 0 : 1;    // the A/B path is bilinear only
  p->fuse_skip = getenv("DSR_NO_FUSE_SKIP") ? 0 : 1;
  p->fuse_ubstats = getenv("DSR_NO_FUSE_UBSTATS") ? 0 : 1;
  p->bound = true;
  p->have_forward = false;
  return 0;
}

static int net_forward_impl(dsr_plan_t* p, const float* params, const float* z, float* out, float* bn_buffers,
                            void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  cudaError_t e = cudaMemsetAsync(p->base + p->acc_fwd_off, 0, p->acc_fwd_bytes, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  // The fp16 weight packing (25 us) depends only on the parameters and the input packing (29 us) only on z and the
  // fp32 skip weights: they run side by side (pack on the side stream, joined before the first convolution).
  const bool fork_head = p->side != nullptr && p->use_side && p->profile != 2 && !getenv("DSR_NO_HEAD_FORK");
  if (fork_head) {
    e = cudaEventRecord(p->ev_fork, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaStreamWaitEvent(p->side, p->ev_fork, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  DSR_TRY(launch_pack_weights(params, p->warena.ptr, static_cast<const PackDesc*>(p->pack_table.ptr),
                              static_cast<int>(p->pack_host.size()), fork_head ? p->side : s));
  if (fork_head) {
    e = cudaEventRecord(p->ev_join, p->side);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  Level& L0 = p->lv[0];
  {
    // level 0's skip conv and (whole-step path) the input perturbation ride on the packing pass when it has the
    // fast layout; otherwise they are separate launches
    const bool fast = input_pack_fast(L0.Cin, L0.W) != 0;
    const bool fskip = fast && p->fuse_skip;
    float* zw = const_cast<float*>(z);
    const long long nz = static_cast<long long>(L0.Cin) * L0.H * L0.W;
    if (p->pz_saved != nullptr && !fast)
      DSR_TRY(launch_perturb(p->pz_saved, zw, nz, p->psigma, p->pseed, 0, s, static_cast<float*>(p->stepstate.ptr)));
    DSR_TRY(launch_input_pack(zw, L0.xin.ptr, L0.Cin, L0.H, L0.W, s, fskip ? params + L0.skip_w : nullptr,
                              fskip ? static_cast<float*>(L0.sraw.ptr) : nullptr,
                              fskip ? reinterpret_cast<acc_t*>(p->base) + L0.skip_stats_off : nullptr,
                              fast ? p->pz_saved : nullptr, p->psigma, p->pseed,
                              static_cast<const float*>(p->stepstate.ptr), p->pad_zero));
  }
  if (fork_head) {
    e = cudaStreamWaitEvent(s, p->ev_join, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int rc = forward_level(p, 0, params, s);
  if (rc) return rc;
  if (p->fuse_top)
    DSR_TRY(launch_bn_act_final(L0.u2.raw.ptr, conv_bn(p, L0.u2, params), params + p->fin_w, params + p->fin_b, out, L0.H,
                                L0.W, s));
  else
    DSR_TRY(launch_final_conv(L0.u2.act.ptr, params + p->fin_w, params + p->fin_b, out, L0.H, L0.W, s));
  if (bn_buffers != nullptr)
    DSR_TRY(launch_bn_running(static_cast<const BnRunDesc*>(p->bnrun_table.ptr), static_cast<int>(p->bnrun_host.size()),
                              reinterpret_cast<const acc_t*>(p->base), params, bn_buffers, kMomentum, s));
  p->have_forward = true;
  return 0;
}

static int net_backward_impl(dsr_plan_t* p, const float* params, const float* out, const float* grad_out, float* grads,
                             void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  p->launches = 0;
  cudaError_t e = cudaMemsetAsync(p->base + p->acc_bwd_off, 0, p->acc_bwd_bytes, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemsetAsync(grads, 0, static_cast<size_t>(p->nparam) * 4, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  Level& L0 = p->lv[0];
  if (p->fuse_top) {
    TopBwdArgs ta{};
    ta.gout = grad_out;
    ta.out = out;
    ta.raw = L0.u2.raw.ptr;
    ta.bn = conv_bn(p, L0.u2, params);
    ta.w = params + p->fin_w;
    ta.bstats = reinterpret_cast<acc_t*>(p->base) + L0.u2.bstats_off;
    ta.dr_pad = L0.u2.dr.ptr;
    ta.dgamma = grads + L0.u2.g_off;
    ta.dbeta = grads + L0.u2.be_off;
    ta.dw = reinterpret_cast<acc_t*>(p->base) + p->fin_dw_acc;
    ta.db = ta.dw + 3 * kNC;
    ta.gs = static_cast<float*>(p->gscale.ptr);
    ta.H = L0.H;
    ta.W = L0.W;
    DSR_TRY(launch_bn_bwd_top_stats(ta, s));
    DSR_TRY(launch_bn_bwd_top_apply(ta, s));
  } else {
    DSR_TRY(launch_final_bwd(grad_out, out, L0.u2.act.ptr, params + p->fin_w, L0.g_u2a.ptr,
                             reinterpret_cast<acc_t*>(p->base) + p->fin_dw_acc,
                             reinterpret_cast<acc_t*>(p->base) + p->fin_dw_acc + 3 * kNC,
                             static_cast<const float*>(p->gscale.ptr), L0.H, L0.W, s));
  }
  if (p->side != nullptr && p->use_side && p->profile != 2) {     // the side stream must see the zeroed accumulators
    e = cudaEventRecord(p->ev_fork, s);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaStreamWaitEvent(p->side, p->ev_fork, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int rc = backward_level(p, 0, params, grads, s);
  if (rc) return rc;
  if (p->side != nullptr && p->use_side && p->profile != 2) {
    e = cudaEventRecord(p->ev_join, p->side);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaStreamWaitEvent(s, p->ev_join, 0);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  DSR_TRY(launch_unpack_wgrad(reinterpret_cast<const float*>(p->base) + p->garena_off, grads,
                              static_cast<const PackDesc*>(p->pack_table.ptr), static_cast<int>(p->pack_host.size()),
                              static_cast<const float*>(p->gscale.ptr), s));
  DSR_TRY(launch_small_grads_finish(static_cast<const SmallGradDesc*>(p->small_table.ptr),
                                    static_cast<int>(p->small_host.size()), reinterpret_cast<const acc_t*>(p->base), grads,
                                    static_cast<const float*>(p->gscale.ptr), s));
  DSR_TRY(launch_grad_scale_finish(grads, p->nparam, static_cast<float*>(p->gscale.ptr), s));
  ++p->launches;
  return 0;
}

// Runs `fn` eagerly, or -- on a capturable stream, from the second call with the same pointer arguments on --
// replays a CUDA graph of it.
extern "C++" {
template <class Fn>
static int run_pass_cached(dsr_plan_t* p, int kind, const void* a0, const void* a1, const void* a2, const void* a3,
                           cudaStream_t s, Fn fn) {
  const bool capturable = (s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread);
  if (!p->use_graph || p->profile || p->debug_conv || !capturable) return fn();
  dsr_plan::PassGraph* hit = nullptr;
  for (auto& g : p->pass_graphs)
    if (g.kind == kind && g.a0 == a0 && g.a1 == a1 && g.a2 == a2 && g.a3 == a3) { hit = &g; break; }
  if (hit == nullptr) {
    if (p->pass_graphs.size() >= 8) {                    // bounded: drop the oldest entry
      if (p->pass_graphs.front().exec) cudaGraphExecDestroy(p->pass_graphs.front().exec);
      p->pass_graphs.erase(p->pass_graphs.begin());
    }
    p->pass_graphs.push_back(dsr_plan::PassGraph{kind, a0, a1, a2, a3, 1, nullptr});
    return fn();
  }
  if (hit->exec == nullptr) {
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int rc = fn();
    e = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaGraphInstantiate(&hit->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { hit->exec = nullptr; return static_cast<int>(e); }
  }
  ++hit->uses;
  return static_cast<int>(cudaGraphLaunch(hit->exec, s));
}
}  // extern "C++"

int dsr_net_forward(dsr_plan_t* p, const float* params, const float* z, float* out, float* bn_buffers, void* stream) {
  if (!p || !params || !z || !out) return -1;
  if (!p->bound) return -6;
  if (p->in_step) return net_forward_impl(p, params, z, out, bn_buffers, stream);
  const int rc = run_pass_cached(p, 0, params, z, out, bn_buffers, static_cast<cudaStream_t>(stream),
                                 [&]() { return net_forward_impl(p, params, z, out, bn_buffers, stream); });
  if (rc == 0) p->have_forward = true;
  return rc;
}

int dsr_net_backward(dsr_plan_t* p, const float* params, const float* out, const float* grad_out, float* grads,
                     void* stream) {
  if (!p || !params || !out || !grad_out || !grads) return -1;
  if (!p->bound) return -6;
  if (!p->have_forward) return -7;
  if (p->in_step) return net_backward_impl(p, params, out, grad_out, grads, stream);
  return run_pass_cached(p, 1, params, out, grad_out, grads, static_cast<cudaStream_t>(stream),
                         [&]() { return net_backward_impl(p, params, out, grad_out, grads, stream); });
}

int dsr_adam_step(float* pp, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int t, void* stream) {
  if (!pp || !g || !m || !v || n < 0 || t < 1) return -1;
  return launch_adam(pp, g, m, v, n, lr, beta1, beta2, eps, t, static_cast<cudaStream_t>(stream));
}

int dsr_perturb(const float* z_saved, float* z, long long n, float sigma, unsigned long long seed,
                unsigned long long offset, void* stream) {
  if (!z_saved || !z || n < 0) return -1;
  return launch_perturb(z_saved, z, n, sigma, seed, offset, static_cast<cudaStream_t>(stream));
}

// Enqueues one iteration.  t_set > 0: explicit iteration index; t_set == 0: the device counter is incremented
// (graph replay).  `losses` is the base of the loss array: this iteration's loss lands in losses[t - 1].
static int enqueue_step(dsr_plan_t* p, const dsr_downsampler_t* d, const dsr_step_buffers_t* b, float* losses, float lr,
                        float sigma, unsigned long long seed, int t_set, cudaStream_t s) {
  void* stream = static_cast<void*>(s);
  float* st = static_cast<float*>(p->stepstate.ptr);
  const long long nz = static_cast<long long>(p->input_depth) * p->H * p->W;
  int total = 0;
  if (timeline_enabled()) g_timeline.n = 0;          // every (eager or captured) iteration stamps slots 0 .. n - 1
  // profiled pass: the launches are enqueued eagerly, each bracketed by events; a busy-wait in front lets the host
  // get ahead so that the intervals measure the GPU, not the host's launch rate
  if (p->profile == 1 && launch_spin(4000000ull, s)) return -1;
  int rc = launch_step_begin(st, losses, t_set, lr, 0.9f, 0.999f, s, static_cast<const float*>(p->gscale.ptr));
  if (rc) return rc;
  total += 1;
  p->in_step = 1;
  p->pz_saved = b->z_saved;          // z = z_saved + sigma N(0,1) is drawn inside the forward pass (input packing)
  p->psigma = sigma;
  p->pseed = seed;
  rc = dsr_net_forward(p, b->params, b->z, b->out_hr, b->bn_buffers, stream);
  p->pz_saved = nullptr;
  p->in_step = 0;
  if (rc) return rc;
  p->have_forward = true;
  total += p->launches;
  if ((rc = launch_downsample_mse(b->out_hr, b->lr_image, b->out_lr, b->g_out_lr, losses, p->n_out, d->H, d->W, d->oh,
                                  d->ow, d->t, s, st)))
    return rc;
  if ((rc = launch_downsample_bwd(b->g_out_lr, b->g_out_hr, p->n_out, d->H, d->W, d->oh, d->ow, d->t, s))) return rc;
  total += 2;
  p->in_step = 1;
  rc = dsr_net_backward(p, b->params, b->out_hr, b->g_out_hr, b->grads, stream);
  p->in_step = 0;
  if (rc) return rc;
  total += p->launches;
  if ((rc = launch_adam(b->params, b->grads, b->adam_m, b->adam_v, p->nparam, lr, 0.9f, 0.999f, 1e-8f, 1, s, st,
                        static_cast<const float*>(p->gscale.ptr))))
    return rc;
  p->launches = total + 1;
  return 0;
}

int dsr_dip_step(dsr_plan_t* p, const dsr_downsampler_t* d, const dsr_step_buffers_t* b, float lr, float sigma,
                 unsigned long long seed, int t, void* stream) {
  if (!p || !d || !b || t < 1) return -1;
  if (!p->bound) return -6;
  if (d->H != p->H || d->W != p->W) return -1;
  // loss_out receives this iteration's loss: losses[t - 1] == *loss_out
  return enqueue_step(p, d, b, b->loss_out - (t - 1), lr, sigma, seed, t, static_cast<cudaStream_t>(stream));
}

int dsr_dip_run(dsr_plan_t* p, const dsr_downsampler_t* d, const dsr_step_buffers_t* b, float lr, float sigma,
                unsigned long long seed, int t_first, int n_iters, void* stream) {
  if (!p || !d || !b || t_first < 1 || n_iters < 0) return -1;
  if (!p->bound) return -6;
  if (d->H != p->H || d->W != p->W) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* losses = b->loss_out;           // base of the loss array: iteration t writes losses[t - 1]
  int done = 0;
  const bool same = p->graph_exec != nullptr && memcmp(&p->graph_bufs, b, sizeof(*b)) == 0 && p->graph_ds == d &&
                    p->graph_ds_serial == d->serial &&
                    p->graph_lr == lr && p->graph_sigma == sigma && p->graph_seed == seed;
  // the legacy default stream cannot be captured: callers that want graph replay pass a non-default stream
  const bool capturable = (s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread);
  if (!p->use_graph || p->profile || p->debug_conv || !capturable) {
    for (; done < n_iters; ++done) {
      int rc = enqueue_step(p, d, b, losses, lr, sigma, seed, t_first + done, s);
      if (rc) return rc;
    }
    return 0;
  }
  if (!same) {
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    if (n_iters == 0) return 0;
    // first iteration eagerly (sets function attributes, warms the module), then capture one iteration
    int rc = enqueue_step(p, d, b, losses, lr, sigma, seed, t_first, s);
    if (rc) return rc;
    done = 1;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return static_cast<int>(e);
    rc = enqueue_step(p, d, b, losses, lr, sigma, seed, 0, s);
    e = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { p->graph_exec = nullptr; return static_cast<int>(e); }
    p->graph_bufs = *b;
    p->graph_ds = d;
    p->graph_ds_serial = d->serial;
    p->graph_lr = lr;
    p->graph_sigma = sigma;
    p->graph_seed = seed;
  } else if (n_iters > 0) {
    // position the device counter: the next replay must run iteration t_first
    int rc = launch_step_begin(static_cast<float*>(p->stepstate.ptr), nullptr, t_first - 1 > 0 ? t_first - 1 : 0, lr, 0.9f,
                               0.999f, s);
    if (rc) return rc;
    if (t_first == 1) {      // t_set == 0 means "increment": write t = 0 explicitly
      const float zero = 0.f;
      cudaError_t e = cudaMemcpyAsync(p->stepstate.ptr, &zero, 4, cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) return static_cast<int>(e);
    }
  }
  for (; done < n_iters; ++done) {
    cudaError_t e = cudaGraphLaunch(p->graph_exec, s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return 0;
}

// ---- in-graph timeline (DSR_TIMELINE=1; see dsr_launch.cuh) -------------------------------------------
extern "C++" {
namespace dsr {
void kstamp_set_conv(unsigned long long*);
void kstamp_set_elem(unsigned long long*);
void kstamp_set_ds(unsigned long long*);
}
}
static unsigned long long* g_ks_dev = nullptr;
static void kstamp_arm() {                       // DSR_TIMELINE=2 (-DDSR_KSTAMP builds): in-kernel stamps
  if (timeline_mode() != 2 || g_ks_dev != nullptr) return;
  if (cudaMalloc(&g_ks_dev, (1 + 8 * 4000) * sizeof(unsigned long long)) != cudaSuccess) { g_ks_dev = nullptr; return; }
  cudaMemset(g_ks_dev, 0, (1 + 8 * 4000) * sizeof(unsigned long long));
  kstamp_set_conv(g_ks_dev);
  kstamp_set_elem(g_ks_dev);
  kstamp_set_ds(g_ks_dev);
}
// one line per launch of the LAST iteration: index, entry -> wait-done (us), wait-done -> next launch's wait-done (us),
// wait-done -> end of block 0's body (us, 0 if the kernel has no end stamp), grid, name
static int kstamp_dump(char* buf, size_t cap) {
  Timeline& t = g_timeline;
  if (g_ks_dev == nullptr || t.n == 0) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::vector<unsigned long long> h(1 + 8 * 4000);
  if (cudaMemcpy(h.data(), g_ks_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  const long long total = static_cast<long long>(h[0]);
  size_t off = 0;
  int w = snprintf(buf + off, cap - off, "# records %lld, launches per iteration (host list) %d\n", total, t.n);
  off += w;
  if (total < t.n) return static_cast<int>(off);
  for (int i = 0; i < t.n; ++i) {
    const long long r = (total - t.n + i) % 4000, rn = (total - t.n + i + 1) % 4000;
    const unsigned long long t0 = h[1 + 8 * r], t1 = h[2 + 8 * r], te = h[3 + 8 * r], gb = h[4 + 8 * r];
    const unsigned long long t1n = (i + 1 < t.n) ? h[2 + 8 * rn] : t1;
    const char* name = "?";
    cudaFuncGetName(&name, t.fn[i]);
    char marks[96] = "";
    {
      int mo = 0;
      for (int m = 0; m < 4; ++m) {
        const unsigned long long tm = h[5 + 8 * r + m];
        if (tm > t1) mo += snprintf(marks + mo, sizeof(marks) - mo, " m%d=%.2f", m, (double)(tm - t1) * 1e-3);
      }
    }
    w = snprintf(buf + off, cap - off, "%d\t%.2f\t%.2f\t%.2f\t%u/%u\t%u\t%s%s\n", i, (double)(t1 - t0) * 1e-3,
                 (double)(t1n - t1) * 1e-3, te > t1 ? (double)(te - t1) * 1e-3 : 0.0, (unsigned)(gb >> 32),
                 (unsigned)(gb & 0xffffffffu), t.grid[i], name, marks);
    if (w < 0 || static_cast<size_t>(w) >= cap - off) break;
    off += w;
  }
  return static_cast<int>(off);
}

int dsr_timeline_dump(char* buf, size_t cap) {
  if (buf == nullptr || cap == 0) return -1;
  buf[0] = 0;
  if (timeline_mode() == 2) return kstamp_dump(buf, cap);
  Timeline& t = g_timeline;
  if (!timeline_enabled() || t.buf == nullptr || t.n == 0) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::vector<unsigned long long> h(t.n);
  if (cudaMemcpy(h.data(), t.buf, t.n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  size_t off = 0;
  for (int i = 1; i < t.n; ++i) {
    const char* name = "?";
    cudaFuncGetName(&name, t.fn[i]);
    // duration = time since the previous stamp ON THE SAME STREAM (the side stream runs beside the main one); the
    // first launch of a stream after a fork is measured from the launch that preceded the fork
    int prev = i - 1;
    while (prev > 0 && t.stream[prev] != t.stream[i]) --prev;
    double us = (static_cast<double>(h[i]) - static_cast<double>(h[prev])) * 1e-3;
    if (t.stream[i] != t.stream[0]) us = -us;               // side-stream launches are reported with a minus sign
    const int w = snprintf(buf + off, cap - off, "%d\t%.2f\t%u\t%s\n", i, us, t.grid[i], name);
    if (w < 0 || static_cast<size_t>(w) >= cap - off) break;
    off += w;
  }
  return static_cast<int>(off);
}

// ---- introspection ------------------------------------------------------------------------------
int dsr_plan_tensor(const dsr_plan_t* p, const char* name, void** ptr, int* kind, int* padded, int* H, int* W,
                    int* C) {
  if (!p || !name) return -1;
  for (const TensorInfo& t : p->tensors)
    if (t.name == name) {
      if (ptr) *ptr = t.buf ? t.buf->ptr : static_cast<void*>(reinterpret_cast<float*>(p->base) + t.acc_off);
      if (kind) *kind = t.kind;
      if (padded) *padded = t.padded;
      if (H) *H = t.H;
      if (W) *W = t.W;
      if (C) *C = t.C;
      return 0;
    }
  return -1;
}
int dsr_plan_last_launches(const dsr_plan_t* p) { return p ? p->launches : -1; }
int dsr_plan_deterministic(const dsr_plan_t* p) { return p ? p->det : -1; }
int dsr_plan_set_debug_conv(dsr_plan_t* p, int use_checker_kernels) {
  if (!p) return -1;
  // 0: product kernels; 1: CUDA-core checker kernels; 2: product path but with the generic implicit-GEMM kernel
  // instead of the halo-tile kernel on the 3x3 stride-1 layers (A/B comparison)
  p->debug_conv = (use_checker_kernels == 1) ? 1 : 0;
  p->use_halo = (use_checker_kernels == 2) ? 0 : 1;
  return 0;
}
int dsr_plan_debug_replay(dsr_plan_t* p, const char* layer, int what, int use_checker, void* stream) {
  if (!p || !layer || !p->bound) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (size_t i = 0; i < p->lv.size(); ++i) {
    Level& L = p->lv[i];
    for (ConvLayer* c : {&L.d1, &L.d2, &L.u1, &L.u2}) {
      if (("L" + std::to_string(i) + "." + c->tag) != layer) continue;
      const int saved = p->debug_conv;
      p->debug_conv = use_checker ? 1 : 0;
      int rc = 0;
      float* acc = reinterpret_cast<float*>(p->base);
      if (what == 0) {
        cudaMemsetAsync(reinterpret_cast<acc_t*>(p->base) + c->stats_off, 0, 2 * kNC * sizeof(acc_t), s);
        rc = run_fprop(p, *c, s);
      } else if (what == 1) {
        if (!c->need_dgrad) rc = -1;
        for (int k = 0; k < c->ndgrad && rc == 0; ++k) rc = run_dgrad(p, *c, k, s);
      } else if (what == 2) {
        cudaMemsetAsync(acc + p->garena_off + c->pack.g_off, 0,
                        static_cast<size_t>(c->k) * c->k * kNC * c->cin_pad * 4, s);
        rc = run_wgrad(p, *c, s);
      } else {
        rc = -1;
      }
      p->debug_conv = saved;
      return rc;
    }
  }
  return -1;
}
// in-kernel stamps of the recorded launches (device -> host after a full synchronisation)
static int prof_stamps(dsr_plan_t* p, std::vector<unsigned long long>& out) {
  out.assign(2 * kProfSlots, 0ull);
  if (!p->bound || p->prof_slots.ptr == nullptr) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  return cudaMemcpy(out.data(), p->prof_slots.ptr, out.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
// milliseconds of one recorded launch: the kernel's own stamps (dependencies satisfied -> last CTA done) when it
// wrote them, else the event interval around the launch
static double prof_ms(const std::vector<unsigned long long>& st, int slot, float event_ms) {
  if (slot >= 0 && st[2 * slot + 1] > st[2 * slot] && st[2 * slot] != ~0ull)
    return static_cast<double>(st[2 * slot + 1] - st[2 * slot]) * 1e-6;
  return event_ms;
}

int dsr_plan_set_profile(dsr_plan_t* p, int on) {
  if (!p) return -1;
  for (auto& r : p->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  p->prof.clear();
  for (auto& r : p->allprof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  p->allprof.clear();
  p->profile = (on == 2) ? 2 : (on ? 1 : 0);
  if (p->profile == 1 && p->bound) {               // re-arm the in-kernel stamp slots: {min start = max, max end = 0}
    std::vector<unsigned long long> init(2 * kProfSlots);
    for (int i = 0; i < kProfSlots; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
    cudaError_t e = cudaMemcpy(p->prof_slots.ptr, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return 0;
}
// profile mode 2: writes "<microseconds>\t<call site>" lines for every launch recorded since set_profile(2)
int dsr_plan_profile_dump(dsr_plan_t* p, char* buf, size_t cap) {
  if (!p || !buf || cap == 0) return -1;
  size_t off = 0;
  buf[0] = 0;
  for (auto& r : p->allprof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return -1;
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    char name[48];
    size_t n = 0;
    for (const char* c = r.what; *c && *c != '(' && n < sizeof(name) - 1; ++c) name[n++] = *c;
    name[n] = 0;
    const int w = snprintf(buf + off, cap - off, "%.2f\t%s\n", t * 1e3f, name);
    if (w < 0 || static_cast<size_t>(w) >= cap - off) break;
    off += static_cast<size_t>(w);
  }
  return static_cast<int>(off);
}
int dsr_plan_profile_read(dsr_plan_t* p, int cls, double* ms_total, double* flops_total, int* launches) {
  if (!p) return -1;
  double ms = 0.0, fl = 0.0;
  int n = 0;
  std::vector<unsigned long long> stamps;
  if (prof_stamps(p, stamps)) return -1;
  for (auto& r : p->prof) {
    if (r.cls != cls) continue;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return static_cast<int>(e);
    float t = 0.f;
    e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) return static_cast<int>(e);
    ms += prof_ms(stamps, r.slot, t);
    fl += r.flops;
    ++n;
  }
  if (ms_total) *ms_total = ms;
  if (flops_total) *flops_total = fl;
  if (launches) *launches = n;
  return 0;
}
// the largest launch (by algorithmic FLOPs) of class `cls` among the recorded ones: mean milliseconds over its
// occurrences and its FLOPs
int dsr_plan_profile_top(dsr_plan_t* p, int cls, double* ms_mean, double* flops) {
  if (!p) return -1;
  double best = 0.0, ms = 0.0;
  int n = 0;
  std::vector<unsigned long long> stamps;
  if (prof_stamps(p, stamps)) return -1;
  for (auto& r : p->prof)
    if (r.cls == cls && r.flops > best) best = r.flops;
  for (auto& r : p->prof) {
    if (r.cls != cls || r.flops != best) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return -1;
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms += prof_ms(stamps, r.slot, t);
    ++n;
  }
  if (ms_mean) *ms_mean = n ? ms / n : 0.0;
  if (flops) *flops = best;
  return 0;
}
int dsr_plan_device_error(dsr_plan_t* p, int* host_code) {
  if (!p || !p->bound || !host_code) return -1;
  cudaError_t e = cudaMemcpy(host_code, p->errword.ptr, sizeof(int), cudaMemcpyDeviceToHost);
  return static_cast<int>(e);
}
int dsr_debug_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return -1;
  return static_cast<int>(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
}

}  // extern "C"
