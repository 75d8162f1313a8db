/* dsr_b200.h -- C ABI of libdsr_b200.so: the B200-native (sm_100a) Deep-Image-Prior
 * super-resolution step.
 *
 * The reference (LewisClifton/Deep-Super-Resolution) is pure Python on torch.nn: it has no FFI
 * for this path, so there is nothing to "bind" -- the entry points below are what the Python
 * mirror of the reference call surface (deep-super-resolution_b200/{models/DIP,utils}) calls
 * through ctypes.  Each entry point cites the reference code whose arithmetic it replaces
 * (paths relative to the upstream repo root).
 *
 * Conventions
 *   - the reference's own calls (torch.nn modules) take NCHW fp32 tensors of batch 1; so do these entry points;
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; the library keeps no global mutable state besides opaque plans;
 *   - return value: 0 = ok, negative = bad argument / unsupported configuration / protocol
 *     error, positive = cudaError_t.  Nothing throws.
 *   - memory: the caller owns every buffer, including the plan workspace
 *     (dsr_plan_workspace_bytes / dsr_plan_bind); a plan owns only small host tables.
 */
#ifndef DSR_B200_H_
#define DSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsr_plan dsr_plan_t;

/* Library / build identification: returns the ABI version (currently 1). */
int dsr_abi_version(void);
/* Human-readable description of error codes returned by this library (static string). */
const char* dsr_error_string(int code);

/* ---- Lanczos downsampler --------------------------------------------------------------
 * Replaces utils/downsampler.py:5-71 (Downsampler.__init__/forward, kernel_type 'lanczos2' or
 * 'lanczos3', phase 0.5, preserve_size=True) and utils/downsampler.py:73-134 (get_kernel). */

/* Writes the normalised (4f x 4f or 6f x 6f) 2-D kernel in float64 to host_kernel (k*k doubles)
 * exactly as get_kernel(factor,'lanczos',0.5,kernel_width,support) does; returns k, or <0. */
int dsr_lanczos_kernel(int factor, int support, double* host_kernel, int capacity);

typedef struct dsr_downsampler dsr_downsampler_t;
/* Device tables for an n_planes x H x W input; table_ws must hold dsr_downsampler_table_bytes(). */
size_t dsr_downsampler_table_bytes(int factor, int support, int H, int W);
int dsr_downsampler_create(dsr_downsampler_t** out, int factor, int support, int H, int W, void* table_ws,
                           size_t table_bytes, void* stream);
void dsr_downsampler_destroy(dsr_downsampler_t* d);
/* y[C][H/f][W/f] = D(x[C][H][W]), fp32 NCHW (batch 1). */
int dsr_downsample_fwd(const dsr_downsampler_t* d, const float* x, float* y, int C, void* stream);
/* gx = D^T gy (autograd backward of the above w.r.t. its input). */
int dsr_downsample_bwd(const dsr_downsampler_t* d, const float* gy, float* gx, int C, void* stream);
/* Fused closure tail (DIP.py:62-65): y = D(x); *loss += mean((y-target)^2); gy = dloss/dy. */
int dsr_downsample_mse(const dsr_downsampler_t* d, const float* x, const float* target, float* y, float* gy,
                       float* loss, int C, void* stream);

/* ---- skip network ---------------------------------------------------------------------
 * Replaces models/DIP/__init__.py:8-18 (get_net) + models/DIP/skip.py:3-96 (skip) as
 * instantiated at DIP.py:169-174: NET_TYPE 'skip', pad 'reflection', upsample 'bilinear',
 * LeakyReLU, need_sigmoid, need_bias, n33d = n33u = 128, n11 = 4, stride downsampling;
 * input_depth in {8k}, num_scales in [1, 6], any H, W that keep every level >= 2 pixels. */
int dsr_plan_create(dsr_plan_t** out, int H, int W, int input_depth, int num_scales, int n_out);
/* The other two get_net options reachable through its signature (models/DIP/__init__.py:8): pad='zero' -- Conv2d's own
 * zero padding instead of ReflectionPad2d (models/DIP/utils.py:96-102; state_dict keys of the convolutions then end in
 * ".0.weight" instead of ".1.weight") -- and upsample_mode='nearest' (models/DIP/skip.py:77).  flags = 0 is
 * dsr_plan_create. */
#define DSR_PLAN_PAD_ZERO 1
#define DSR_PLAN_UP_NEAREST 2
int dsr_plan_create_ex(dsr_plan_t** out, int H, int W, int input_depth, int num_scales, int n_out, int flags);
void dsr_plan_destroy(dsr_plan_t* p);
/* Flat parameter buffer: concatenation of the reference's net.named_parameters() in order. */
int dsr_plan_num_params(const dsr_plan_t* p);                 /* number of tensors (112)    */
long long dsr_plan_param_numel(const dsr_plan_t* p);          /* total floats (2 217 831)   */
/* idx-th tensor: name (reference state_dict key), offset into the flat buffer, shape (up to 4). */
int dsr_plan_param_info(const dsr_plan_t* p, int idx, char* name, int name_cap, long long* offset, int* ndim,
                        int* shape4);
/* BatchNorm buffers: flat [running_mean(C) | running_var(C)] per BatchNorm in state_dict order. */
int dsr_plan_num_bn(const dsr_plan_t* p);
long long dsr_plan_bn_numel(const dsr_plan_t* p);
int dsr_plan_bn_info(const dsr_plan_t* p, int idx, char* name, int name_cap, long long* offset, int* channels);
size_t dsr_plan_workspace_bytes(const dsr_plan_t* p);
/* Binds (and zero-fills) the workspace and builds the TMA descriptors; must precede forward. */
int dsr_plan_bind(dsr_plan_t* p, void* workspace, size_t bytes, void* stream);

/* out[n_out][H][W] = net(z[input_depth][H][W])   (skip.py forward; BatchNorm in train mode).
 * bn_buffers (may be NULL): running statistics are updated as torch does (momentum 0.1). */
int dsr_net_forward(dsr_plan_t* p, const float* params, const float* z, float* out, float* bn_buffers, void* stream);
/* grads (flat, same layout as params) = d<grad_out, out>/dparams for the LAST forward of this
 * plan; `out` is that forward's output.  Conv biases feeding a BatchNorm get exact zeros. */
int dsr_net_backward(dsr_plan_t* p, const float* params, const float* out, const float* grad_out, float* grads,
                     void* stream);

/* ---- optimiser ------------------------------------------------------------------------
 * Replaces torch.optim.Adam(params, lr) as used by utils/DIP.py:33-38 (defaults: betas
 * (0.9, 0.999), eps 1e-8, no weight decay): one fused pass over flat p, g, m, v; t = 1-based
 * step count. */
int dsr_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int t, void* stream);

/* z = z_saved + sigma * N(0,1) (DIP.py:52) with a counter-based Philox generator on the device. */
int dsr_perturb(const float* z_saved, float* z, long long n, float sigma, unsigned long long seed,
                unsigned long long offset, void* stream);

/* ---- whole DIP iteration (closure of DIP.py:47-95 + utils/DIP.py:35-38), one host call ----
 * perturb -> forward -> downsample -> MSE -> backward -> Adam.  loss_out: device float that
 * receives this iteration's loss.  t = 1-based iteration index. */
typedef struct dsr_step_buffers {
  float* params; float* grads; float* adam_m; float* adam_v;   /* flat, dsr_plan_param_numel floats */
  float* bn_buffers;                                           /* flat, dsr_plan_bn_numel floats, or NULL */
  const float* z_saved; float* z;                              /* [input_depth][H][W] */
  const float* lr_image;                                       /* [n_out][H/f][W/f] */
  float* out_hr; float* out_lr; float* g_out_lr; float* g_out_hr;
  float* loss_out;
} dsr_step_buffers_t;
int dsr_dip_step(dsr_plan_t* p, const dsr_downsampler_t* d, const dsr_step_buffers_t* b, float lr, float sigma,
                 unsigned long long seed, int t, void* stream);
/* n_iters consecutive iterations t_first .. t_first + n_iters - 1 (the loop of utils/DIP.py:35-38).  Here
 * b->loss_out is the BASE of a loss array: iteration t writes loss_out[t - 1].  The iteration counter and Adam's
 * bias corrections live on the device, so one iteration is captured once as a CUDA graph (per set of buffers /
 * hyper-parameters) and replayed: ~180 kernel launches per iteration become one cudaGraphLaunch. */
int dsr_dip_run(dsr_plan_t* p, const dsr_downsampler_t* d, const dsr_step_buffers_t* b, float lr, float sigma,
                unsigned long long seed, int t_first, int n_iters, void* stream);

/* ---- introspection for tests / profiling --------------------------------------------------
 * Named intermediate of the last forward/backward (e.g. "L0.d1_raw"): device pointer, element
 * kind (0 fp16, 1 bf16, 2 fp32, 3 = 64-bit fixed-point accumulator in 2^-20 units), padded-grid flag, H, W, C.
 * Returns 0 or -1 if unknown. */
int dsr_plan_tensor(const dsr_plan_t* p, const char* name, void** ptr, int* kind, int* padded, int* H, int* W,
                    int* C);
/* Kernel launches enqueued by the last forward / backward / step call on this plan. */
int dsr_plan_last_launches(const dsr_plan_t* p);
/* Debug switch (tests only): 1 = run convolutions with the naive CUDA-core checker kernels; 2 = product kernels
 * but the generic implicit-GEMM kernel instead of the halo-tile kernel on 3x3 stride-1 layers; 0 = product. */
int dsr_plan_set_debug_conv(dsr_plan_t* p, int use_checker_kernels);
/* Re-runs ONE tensor-core kernel of one conv layer ("L0.d1", "L2.u1", ...) on the current workspace
 * contents: what = 0 fprop (zeroes the layer's BN sums first), 1 dgrad, 2 wgrad (zeroes the layer's packed
 * gradient first); use_checker = 1 runs the CUDA-core checker kernel instead.  Tests compare the two. */
int dsr_plan_debug_replay(dsr_plan_t* p, const char* layer, int what, int use_checker, void* stream);
/* Per-launch CUDA-event timing of the tensor-core kernels (used by bench.py for the roofline figure; off by
 * default).  set_profile(on) clears earlier records.  profile_read sums, for kernel class `cls` (0 = halo-tile
 * conv kernel, 3x3 stride-1 fprop + dgrad launches (tensor-bound), 1 = wgrad kernels, 2 = generic implicit-GEMM conv
 * kernel: stride-2 layers, 3 = halo-tile conv kernel, 1x1 launches (64 FLOP per byte: HBM-bound)), the event-measured milliseconds, the algorithmic FLOPs
 * (2*M*N*K per pass, true channel counts) and the number of launches recorded since; it synchronises on them. */
/* set_profile(p, 1): the tensor-core kernels stamp %globaltimer themselves -- the span from the moment their
 * dependencies are satisfied to their last CTA's exit, with the pipeline's programmatic overlap intact -- and
 * profile_read / profile_top report those spans (event intervals only for launches without stamps). */
int dsr_plan_set_profile(dsr_plan_t* p, int on);
int dsr_plan_profile_read(dsr_plan_t* p, int cls, double* ms_total, double* flops_total, int* launches);
/* set_profile(p, 2): every launch of forward / backward is bracketed by events on the main stream (side stream and
 * graph replay off); profile_dump writes one "<microseconds>\t<call site>" line per launch into buf, returns bytes. */
int dsr_plan_profile_dump(dsr_plan_t* p, char* buf, size_t cap);
/* The largest recorded launch (by algorithmic FLOPs) of class cls: mean milliseconds over its occurrences, its FLOPs. */
int dsr_plan_profile_top(dsr_plan_t* p, int cls, double* ms_mean, double* flops);
/* Diagnostics (process started with DSR_TIMELINE=1): every kernel launch of an iteration is followed by a stamp
 * kernel that records %globaltimer; dump writes "<index>\t<microseconds since the previous stamp>\t<grid>\t<kernel>"
 * lines for the most recent (replayed) iteration into buf and returns the number of bytes. */
int dsr_timeline_dump(char* buf, size_t cap);
/* Device error word written by a kernel whose mbarrier wait timed out (0 = none). */
int dsr_plan_device_error(dsr_plan_t* p, int* host_code);
/* Device-to-device copy on `stream` (lets ctypes callers read an introspected tensor into their own buffer). */
int dsr_debug_copy(void* dst, const void* src, size_t bytes, void* stream);

/* ---- SRResNet generator, inference ------------------------------------------------------------
 * Replaces models/GAN/generator.py:4-81 (Generator / ResidualBlock / PixelShuffleBlock) in EVAL mode as
 * eval_GAN.py:87-94 runs it: y[B][3][f h][f w] = tanh(conv3(shuffle blocks(x0 + bn1(conv2(residual blocks(x0))))))
 * with x0 = prelu1(conv1(x)), BatchNorm using its running statistics.  factor 8 (3 PixelShuffle blocks) or 16 (4) --
 * the only two the reference constructs (generator.py:55-58).  One plan per (factor, blocks, batch, h, w). */
typedef struct dsr_gen_plan dsr_gen_plan_t;
int dsr_gen_plan_create(dsr_gen_plan_t** out, int factor, int residual_blocks, int batch, int h, int w);
void dsr_gen_plan_destroy(dsr_gen_plan_t* p);
/* Flat fp32 state: the float tensors of the reference module's state_dict() in order (num_batches_tracked
 * omitted).  tensor_info gives the idx-th key, its offset and element count. */
long long dsr_gen_state_numel(const dsr_gen_plan_t* p);
int dsr_gen_num_tensors(const dsr_gen_plan_t* p);
int dsr_gen_tensor_info(const dsr_gen_plan_t* p, int idx, char* name, int name_cap, long long* offset, long long* numel);
size_t dsr_gen_workspace_bytes(const dsr_gen_plan_t* p);
/* Binds (and zero-fills) the caller-owned workspace (1024-byte aligned) and builds the TMA descriptors. */
int dsr_gen_bind(dsr_gen_plan_t* p, void* workspace, size_t bytes, void* stream);
/* Folds eval-mode BatchNorm into the convolutions and packs fp16 GEMM weights from the flat state (device pointer);
 * call again after the weights change (load_state_dict).  Synchronises `stream` once. */
int dsr_gen_load_weights(dsr_gen_plan_t* p, const float* state, void* stream);
/* y = Generator(x).  x: [B][3][h][w] fp32, y: [B][3][f h][f w] fp32, both NCHW device buffers. */
int dsr_gen_forward(dsr_gen_plan_t* p, const float* x, float* y, void* stream);
int dsr_gen_last_launches(const dsr_gen_plan_t* p);
/* Tests: intermediate activation ("x0", "t1", "xa", "xb", "s0".."s3"): fp16 [rows][W][64], image b at rows
 * [b * img_rows, b * img_rows + H). */
int dsr_gen_debug_tensor(const dsr_gen_plan_t* p, const char* name, void** ptr, int* rows, int* img_rows, int* H, int* W);
int dsr_gen_device_error(dsr_gen_plan_t* p, int* host_code);

/* ---- image-quality metrics of the logging branch (DIP.py:71-87,183-185) ---------------------------------
 * Replace the torchmetrics objects DIP.py:157-158 constructs -- PeakSignalNoiseRatio() and
 * StructuralSimilarityIndexMeasure(data_range=1.) (third-party torchmetrics, not part of the reference checkout;
 * algorithm restated in csrc/dsr_metrics.cu) -- by one kernel launch each that leaves a
 * single float on the device.  workspace: dsr_metric_workspace_bytes() bytes, zeroed once by the caller (the kernels
 * re-arm it), 16-byte aligned, used by one launch at a time.  pred / target: fp32, same shape. */
size_t dsr_metric_workspace_bytes(void);
/* out[0] = 10 log10(range^2 / mean((pred - target)^2)); data_range <= 0: range = max(target) - min(min(target), 0)
 * (torchmetrics' data_range=None on a single update). */
int dsr_psnr(const float* pred, const float* target, long long n, float data_range, void* workspace, float* out,
             void* stream);
/* out[0] = mean SSIM over `planes` (batch x channels) H x W planes: 11 x 11 Gaussian window (sigma 1.5), k1 0.01,
 * k2 0.03, averaged over the pixels whose window lies inside the image. */
int dsr_ssim(const float* pred, const float* target, int planes, int H, int W, float data_range, void* workspace,
             float* out, void* stream);
/* ---- image writer path of eval_GAN.py:50-53 (GAN_ISR_Batch_eval) and utils/common.py:76-88 (np_to_pil) ----------
 * out_hwc[b][y][x][c] = uint8 of chw[b][c][y][x] * 255 on the device, so that a quarter of the bytes crosses PCIe and
 * the host only encodes the PNG.  clip = 0: numpy's (x * 255).astype(np.uint8) of eval_GAN.py:52 (truncate towards
 * zero, keep the low byte: the generator ends in tanh and negative values wrap; NaN -> 0); clip = 1:
 * np.clip(x * 255, 0, 255).astype(np.uint8) of utils/common.py:81.  chw 16-byte aligned, out_hwc 4-byte aligned. */
int dsr_image_to_u8_hwc(const float* chw, int B, int C, int H, int W, int clip, unsigned char* out_hwc, void* stream);
/* 1 when the plan was created with DSR_DETERMINISTIC=1 (two-stage split-K weight gradients: bit-identical runs). */
int dsr_plan_deterministic(const dsr_plan_t* p);

/* ---- SRGAN training step (train_GAN.py:38-71, do_epoch) ---------------------------------------------------
 * Replaces, for one batch of B <= 8 patches: Generator (models/GAN/generator.py:44-81) in TRAIN mode with its backward
 * pass, Discriminator (models/GAN/discriminator.py:21-74) forward / backward, nn.BCELoss against constant targets
 * (utils/GAN.py:96-107) and the VGG19 content loss behind the IMAGENET1K_V1 transform (utils/GAN.py:62-88) with its
 * gradient w.r.t. the generated image.  bf16 tensor-core convolutions (gconv_kernel / gwgrad_kernel), fp32 master
 * parameters.  net: 0 generator, 1 discriminator, 2 VGG19 features[:36].  Parameters / gradients of a net are ONE flat
 * fp32 array in named_parameters() order (param_info gives name, offset, numel); BatchNorm running statistics are a
 * second flat array (buffer_info; num_batches_tracked is the caller's).  Gradients are ACCUMULATED (+=) into `grads`.
 * The optimiser is dsr_adam_step on the flat arrays; data-parallel training all-reduces the flat gradient arrays. */
typedef struct dsr_gant dsr_gant_t;
int dsr_gant_create(dsr_gant_t** out, int batch, int lr_h, int lr_w, int factor, int residual_blocks, int with_vgg);
void dsr_gant_destroy(dsr_gant_t* p);
long long dsr_gant_param_numel(const dsr_gant_t* p, int net);
long long dsr_gant_buffer_numel(const dsr_gant_t* p, int net);
int dsr_gant_num_params(const dsr_gant_t* p, int net);
int dsr_gant_num_buffers(const dsr_gant_t* p, int net);
int dsr_gant_param_info(const dsr_gant_t* p, int net, int idx, char* name, int name_cap, long long* offset, long long* numel);
int dsr_gant_buffer_info(const dsr_gant_t* p, int net, int idx, char* name, int name_cap, long long* offset, long long* numel);
size_t dsr_gant_workspace_bytes(const dsr_gant_t* p);
/* Binds (and zero-fills) the caller-owned workspace (1024-byte aligned). */
int dsr_gant_bind(dsr_gant_t* p, void* workspace, size_t bytes, void* stream);
/* bf16 GEMM layouts of one net's convolution weights from its flat parameters; call after every parameter update. */
int dsr_gant_pack(dsr_gant_t* p, int net, const float* params, void* stream);
/* out = Generator(lr) in train mode.  lr [B][3][h][w], out [B][3][f h][f w] fp32 NCHW.  buffers (may be NULL): running
 * statistics, updated bn_updates times with this batch (do_epoch runs the generator twice on identical inputs). */
int dsr_gant_g_forward(dsr_gant_t* p, const float* params, float* buffers, const float* lr_nchw, float* out_nchw,
                       int bn_updates, void* stream);
int dsr_gant_g_backward(dsr_gant_t* p, const float* params, const float* dout_nchw, float* grads, void* stream);
/* prob[B] = Discriminator(img); slot 0 / 1 = which activation set keeps the pass for dsr_gant_d_backward. */
int dsr_gant_d_forward(dsr_gant_t* p, int slot, const float* params, float* buffers, const float* img_nchw, float* prob,
                       void* stream);
/* dprob [B] = d(loss)/d(prob), or NULL: BCE(prob, target) with mean reduction fused in. */
int dsr_gant_d_backward(dsr_gant_t* p, int slot, const float* params, const float* dprob, float target, float* grads,
                        void* stream);
/* Both kept passes in one call (loss_D.backward() of do_epoch): slot 0 against target0, slot 1 against target1, BCE fused;
 * the dense head's 73 728 x 1024 matrix and its gradient are swept once for the two passes. */
int dsr_gant_d_backward_pair(dsr_gant_t* p, const float* params, float target0, float target1, float* grads, void* stream);
/* on != 0: dsr_gant_d_backward_pair WRITES the gradient of dense1.weight (73 728 x 1024, 302 MB) instead of adding to it,
 * so the caller need not clear that part of `grads` before the call (loss_D.backward() after zero_grad(),
 * train_GAN.py:50-52): saves one 302 MB fill and one 302 MB read per step.  Every other gradient is still added. */
int dsr_gant_dense_grad_overwrite(dsr_gant_t* p, int on);
/* loss[0] (= | +=) nn.BCELoss()(prob[0..n), target) */
int dsr_gant_bce(dsr_gant_t* p, const float* prob, float target, int n, float* loss, int accumulate, void* stream);
/* loss[0] (= | +=) MSE(VGG(T(fake)), VGG(T(real))); dfake (may be NULL) = its gradient w.r.t. fake.  real_nchw may be
 * NULL when dsr_gant_vgg_real computed the real batch's features before (utils/GAN.py:72-93: the two feature passes are
 * independent; only the second depends on the generator). */
int dsr_gant_vgg_loss(dsr_gant_t* p, const float* fake_nchw, const float* real_nchw, float* loss, int accumulate,
                      float* dfake_nchw, void* stream);
/* VGG(T(real)) alone, kept for the next dsr_gant_vgg_loss(..., NULL, ...); it must be complete before that call starts. */
int dsr_gant_vgg_real(dsr_gant_t* p, const float* real_nchw, void* stream);
int dsr_gant_device_error(dsr_gant_t* p, int* host_code);
int dsr_gant_last_launches(const dsr_gant_t* p);
/* Tests: a named activation of the last pass ([B * P][W][C] tall grid, image b at rows [b P, b P + H)); *f32 receives
 * the element type: 0 bf16, 1 fp32, 2 fp16. */
int dsr_gant_tensor(const dsr_gant_t* p, const char* name, void** ptr, int* C, int* W, int* H, int* P, int* B, int* f32);

#ifdef __cplusplus
}
#endif
#endif /* DSR_B200_H_ */
