/*
 * cgb200.h — C ABI of libcgb200.so: the sm_100a (B200) kernels behind the
 * ClimateGAN conv-GAN hot path (SURVEY.md §8b "Downward" row).
 *
 * The reference (cc-ai/climategan) has no FFI: its "kernels" are torch.nn library
 * calls.  Every entry point below therefore cites the reference *Python* call it
 * replaces (file:line under the reference tree).  A maintainer binds these with
 * ctypes from the reference's own modules — see INTEGRATION.md.
 *
 * Conventions
 *  - All tensors are device pointers owned by the caller (PyTorch's allocator).
 *    The library never allocates, frees or synchronises; all work is enqueued on
 *    `stream` (a cudaStream_t passed as void*).
 *  - Activations are NHWC, contiguous, with the channel count rounded up to a
 *    multiple of 8 ("storage channels"); pad channels hold zeros.
 *  - dtype: CGB_F32 or CGB_BF16 for activations / packed weights.  Accumulation
 *    is always fp32 (fp64 for normalisation statistics).
 *  - Packed conv weights: [co][kh*kw][ci] (K-major, ci fastest), zero padded to the
 *    storage channel counts.  Weight gradients are fp32 in the same layout.
 *  - Return value: CGB_OK (0) or a negative cgb_status; cgb_last_error() returns a
 *    thread-local message.  There is no CPU fallback: without an sm_100 device
 *    every compute entry point returns CGB_UNSUPPORTED_ARCH.
 */
#ifndef CGB200_H
#define CGB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CGB_OK = 0,
  CGB_BAD_ARG = -1,
  CGB_UNSUPPORTED = -2,
  CGB_LAUNCH_FAILURE = -3,
  CGB_UNSUPPORTED_ARCH = -4
} cgb_status;

typedef enum { CGB_F32 = 0, CGB_BF16 = 1, CGB_F16 = 2 } cgb_dtype;

typedef enum {
  CGB_ACT_NONE = 0,
  CGB_ACT_RELU = 1,
  CGB_ACT_LRELU = 2, /* slope given separately (0.2 everywhere in the reference) */
  CGB_ACT_TANH = 3,
  CGB_ACT_SIGMOID = 4,
  CGB_ACT_SELU = 5     /* nn.SELU (Conv2dBlock activation="selu", blocks.py:97-98); elementwise kernels only */
} cgb_act;

typedef enum { CGB_PAD_ZERO = 0, CGB_PAD_REFLECT = 1 } cgb_pad_mode;

typedef enum {
  CGB_ENGINE_AUTO = 0,   /* tcgen05 when the shape qualifies, else SIMT */
  CGB_ENGINE_SIMT = 1,   /* CUDA-core implicit GEMM (any shape, fp32 or bf16) */
  CGB_ENGINE_TCGEN05 = 2 /* TMA + tcgen05/TMEM implicit GEMM (bf16); error if the shape does not qualify */
} cgb_engine;

/* One 2-D convolution, forward geometry.  Used unchanged by fwd / dgrad / wgrad. */
typedef struct cgb_conv_desc {
  int32_t n, hi, wi, ci; /* input  [n,hi,wi,ci]  (ci = storage channels, %8==0) */
  int32_t ho, wo, co;    /* output [n,ho,wo,co]  (co = storage channels, %8==0) */
  int32_t kh, kw;
  int32_t stride, dil, pad; /* symmetric */
  int32_t pad_mode;         /* cgb_pad_mode */
  int32_t dtype;            /* cgb_dtype of x, w, y, residual, mask_src */
  int32_t act;              /* epilogue activation (fwd) */
  float slope;              /* leaky-relu slope */
  int32_t engine;           /* cgb_engine */
  int32_t res_before_act;   /* 1: y = act(conv + bias + residual) (ResNet bottleneck, resnetmulti_v2.py:50-52); 0: act first */
} cgb_conv_desc;

/* ---- library --------------------------------------------------------------------------- */
const char* cgb_version(void);
const char* cgb_last_error(void);
/* 1 when the current device is sm_100 (B200), else 0. */
int cgb_device_ok(void);
/* number of kernels this library has launched since load / since the last reset (bench "gpu_launches") */
int64_t cgb_launch_count(void);
void cgb_launch_count_reset(void);
/* Per-launch timing of the conv engines (CUDA events on the launching stream), for bench.py's roofline
 * leg — the analogue of the reference's utils.Timer (climategan/utils.py:919-959).
 * cgb_prof_dump: after the caller synchronised, writes one line per (op, engine, shape):
 *   "which tc n hi wi ci ho wo co kh kw stride dil count total_ms"  and clears the records. */
void cgb_prof_enable(int on);
int cgb_prof_dump(char* buf, int64_t cap);
/* 1 if cgb_conv2d_fwd / dgrad / wgrad with this desc would run on tcgen05 */
int cgb_conv2d_uses_tcgen05(const cgb_conv_desc* d, int which /*0 fwd, 1 dgrad, 2 wgrad*/);

/* ---- convolution -------------------------------------------------------------------------
 * Replaces nn.Conv2d / F.conv2d as used by Conv2dBlock (climategan/blocks.py:138-144),
 * SPADE.mlp_shared/mlp_gamma/mlp_beta (climategan/norms.py:164-171,180-182) and the
 * SPADEResnetBlock convs (climategan/blocks.py:349-353), padding (blocks.py:66-71) included.
 *   y = act(conv(x, w) + bias) (+ residual)
 * residual: optional [n,ho,wo,co], added after the activation (blocks.py:196, :377).
 */
int cgb_conv2d_fwd(const cgb_conv_desc* d, const void* x, const void* w, const float* bias,
                   const void* residual, void* y, void* stream);

/* The same conv, with the per-channel statistics of ITS OWN OUTPUT accumulated by the kernel's epilogue — the conv -> BatchNorm
 * chains of the masker in train mode (climategan/deeplab/resnetmulti_v2.py:40-56, deeplab_v2.py:23,100-104,146, depth.py:57-105;
 * Conv2dBlock with norm="batch", blocks.py:138-144) need batch statistics of the conv output before they can normalise it; here
 * they come out of the conv launch instead of a second pass over the tensor.  tcgen05 engine only (cgb_conv2d_uses_tcgen05).
 *   stats_partial: [cgb_conv2d_stats_rows()][2][co] fp32 (zeroed by the call): row b = CTA b's sum and sum of squares of the
 *   STORED (storage-dtype-rounded) output over the pixels it produced; feed it to cgb_bn_train_fwd_partials. */
int32_t cgb_conv2d_stats_rows(void);
int cgb_conv2d_fwd_stats(const cgb_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                         float* stats_partial, void* stream);

/* Data gradient of the same conv (autograd of F.conv2d w.r.t. its input):
 *   gx = conv_transpose(gy, w)   [n,hi,wi,ci]
 * dact/mask_src (optional): multiply gx by the derivative of the activation that PRODUCED x,
 * evaluated from mask_src = that layer's output ([n,hi,wi,ci]): relu/lrelu use sign(mask_src),
 * tanh uses 1-mask_src^2.  This fuses e.g. the ReLU of SPADE.mlp_shared (norms.py:165) into
 * the dgrad of mlp_gamma/mlp_beta.  gy is the gradient w.r.t. the conv's pre-activation output. */
int cgb_conv2d_dgrad(const cgb_conv_desc* d, const void* gy, const void* w, const void* wt, int32_t dact,
                     const void* mask_src, void* gx, void* stream);

/* Weight packing the tcgen05 engine uses for the data gradient: wt[ci][kh*kw][co] with the taps reversed,
 * wt[c][T-1-t][o] = w[o][t][c] — the stride-1 dgrad is then an ordinary conv of gy with wt (pad' = dil*(k-1)-pad).
 * cgb_conv2d_dgrad takes w (SIMT engine) and/or wt (tcgen05 engine; NULL -> SIMT). */
int cgb_conv2d_pack_dgrad_weight(const cgb_conv_desc* d, const void* w, void* wt, void* stream);
/* Forward packing in one launch: w OIHW fp32 [o][i][kh][kw] (nn.Conv2d.weight, or W / sigma of SpectralNorm, norms.py:111-112)
 * -> wp [cos][taps = kh*kw][cis] in `dtype`, zeros in the channel padding (cos >= o, cis >= i, cis % 8 == 0). */
int cgb_pack_weight(const float* w, void* wp, int32_t dtype, int32_t o, int32_t i, int32_t taps, int32_t cos, int32_t cis,
                    void* stream);
/* The same plus the dgrad packing wt[cis][taps-1-t][cos] (cgb_conv2d_pack_dgrad_weight's layout) in ONE launch — weights change
 * every optimiser step, and a weight whose conv needs a data gradient is packed both ways once per step. */
int cgb_pack_weight_dual(const float* w, void* wp, void* wt, int32_t dtype, int32_t o, int32_t i, int32_t taps, int32_t cos,
                         int32_t cis, void* stream);

/* Weight (+bias) gradient: gw[co][kh*kw][ci] (fp32), gbias[co] (fp32, optional).
 * accumulate=0 zero-fills gw/gbias first. */
int cgb_conv2d_wgrad(const cgb_conv_desc* d, const void* x, const void* gy, float* gw, float* gbias,
                     int32_t accumulate, void* stream);

/* ---- normalisation -----------------------------------------------------------------------
 * nn.InstanceNorm2d(affine=False), eps 1e-5, biased variance (climategan/norms.py:151,176;
 * discriminator.py:71-73).  x [n,hw,c] -> mean,rstd [n,c] fp32.  ws: cgb_instnorm_ws_doubles(n,hw,c) doubles scratch. */
int cgb_instnorm_stats(const void* x, int32_t dtype, int32_t n, int32_t hw, int32_t c, float eps,
                       double* ws, float* mean, float* rstd, void* stream);
/* scratch the call above needs, in doubles (per-chunk fp32 partial sums; no atomics, deterministic) */
int64_t cgb_instnorm_ws_doubles(int32_t n, int32_t hw, int32_t c);

/* SPADE de-normalisation + following activation (norms.py:184 then blocks.py:372-373,394):
 *   out = act( (x-mean)*rstd * (1+gamma) + beta ),  gamma = gb[...,0:c], beta = gb[...,c:2c]
 * gb is the [n,hw,2c] output of the fused mlp_gamma||mlp_beta conv. */
int cgb_spade_modulate_fwd(const void* x, const float* mean, const float* rstd, const void* gb,
                           void* out, int32_t dtype, int32_t n, int32_t hw, int32_t c, int32_t act,
                           float slope, void* stream);

/* Backward of the above, part 1.  Recomputes the pre-activation, then
 *   gs = gout*act'(pre); ggb = [gs*xhat || gs]; gxhat = gs*(1+gamma)
 * and accumulates sums[n,c,0] += sum_hw gxhat, sums[n,c,1] += sum_hw gxhat*xhat (fp64; caller zeroes). */
int cgb_spade_modulate_bwd(const void* x, const float* mean, const float* rstd, const void* gb,
                           const void* gout, void* ggb, void* gxhat, double* sums, int32_t dtype,
                           int32_t n, int32_t hw, int32_t c, int32_t act, float slope, void* stream);

/* The same pass, also returning the bias gradients of SPADE.mlp_gamma / mlp_beta (norms.py:169-172, nn.Conv2d(nhidden, norm_nc, ...), bias on by default):
 *   bsum[0:c] += sum_{n,hw} gs*xhat,  bsum[c:2c] += sum_{n,hw} gs   (fp64, the column sums of ggb; caller zeroes)
 * so the weight-gradient launch of the gamma||beta conv needs no separate column-sum pass over ggb. */
int cgb_spade_modulate_bwd_bias(const void* x, const float* mean, const float* rstd, const void* gb,
                                const void* gout, void* ggb, void* gxhat, double* sums, double* bsum, int32_t dtype,
                                int32_t n, int32_t hw, int32_t c, int32_t act, float slope, void* stream);

/* Backward of instance norm, part 2 (in place on gxhat -> gx):
 *   gx = rstd * (gxhat - sums0/hw - xhat*sums1/hw) */
int cgb_instnorm_bwd(const void* x, const float* mean, const float* rstd, const double* sums,
                     void* gxhat_inout, int32_t dtype, int32_t n, int32_t hw, int32_t c, void* stream);

/* InstanceNorm2d(affine=False) applied + LeakyReLU, the NLayerDiscriminator block (discriminator.py:120-133,146-148):
 *   y = act((x-mean)*rstd).   Backward part 1: gxhat = gy*act'(xhat), sums += (sum gxhat, sum gxhat*xhat); part 2 is
 *   cgb_instnorm_bwd. */
int cgb_instnorm_apply_fwd(const void* x, const float* mean, const float* rstd, void* y, int32_t dtype, int32_t n,
                           int32_t hw, int32_t c, int32_t act, float slope, void* stream);
int cgb_instnorm_apply_bwd(const void* x, const float* mean, const float* rstd, const void* gy, void* gxhat,
                           double* sums, int32_t dtype, int32_t n, int32_t hw, int32_t c, int32_t act, float slope,
                           void* stream);

/* nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) between discriminator scales
 * (discriminator.py:223-225).  ho = (hi-1)/2+1. */
int cgb_avgpool3s2_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t c, void* stream);
int cgb_avgpool3s2_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t c, void* stream);

/* ---- resampling --------------------------------------------------------------------------
 * F.interpolate(mode="nearest") (blocks.py:39-43 InterpolateNearest2d; norms.py:179; painter.py:152):
 * src index = floor(dst * in/out). */
int cgb_resize_nearest_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi,
                           int32_t ho, int32_t wo, int32_t c, void* stream);
/* adjoint of the above for integer up-scaling factors (ho = hi*f): gx = sum over the f*f replicas */
int cgb_upsample_nearest_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi,
                             int32_t f, int32_t c, void* stream);
/* adjoint of cgb_resize_nearest_fwd for any size ratio (the masker's SPADE layers down-size a differentiable conditioning
 * tensor, norms.py:179), gather form: gx[iy,ix] = sum of gy over the output pixels whose nearest source is (iy,ix) */
int cgb_resize_nearest_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo,
                           int32_t c, void* stream);

/* im2col of a few-channel NHWC tensor: y[n,oy,ox, tap*c + ch] = x[n, oy+dy*dil-pad, ox+dx*dil-pad, ch] (zero outside,
 * zero for channels >= k*k*c).  Turns the 3-channel SPADE.mlp_shared 3x3 conv (norms.py:164-166) into a K=32 1x1
 * GEMM for the tensor cores; computed once per resolution and shared by every SPADE layer at that resolution. */
int cgb_im2col(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t cs_in, int32_t c,
               int32_t k, int32_t pad, int32_t dil, int32_t cs_out, void* stream);
/* The same with an output stride: y is [n, ho, wo, cs_out], ho = (h + 2*pad - dil*(k-1) - 1)/stride + 1.  Turns the ResNet stem
 * (7x7 stride-2 conv on the 3-channel image, resnetmulti_v2.py:70) into one K=152 GEMM: as 49 taps of an 8-channel TMA box it
 * ran at 100 GB/s. */
int cgb_im2col_strided(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t cs_in, int32_t c,
                       int32_t k, int32_t pad, int32_t dil, int32_t stride, int32_t cs_out, void* stream);
/* Adjoint of cgb_im2col_strided (gather form, deterministic): gx[n,y,x,ch] = sum of g[n,oy,ox, tap*c + ch] over the taps and output
 * pixels that read input pixel (y,x); gx is [n,h,w,cs_in] (channels >= c zero), g is [n,ho,wo,cs_col], c <= 8.  With it a
 * first-layer conv on an image that needs a data gradient — NLayerDiscriminator's model0 under the generator loss, D(fake) -> G
 * (discriminator.py:122, trainer.py:1421-1440) — runs as im2col + ONE K = round8(k*k*c) GEMM like the ResNet stem. */
int cgb_col2im_strided(const void* g, void* gx, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t cs_in, int32_t c,
                       int32_t k, int32_t pad, int32_t dil, int32_t stride, int32_t cs_col, void* stream);

/* ---- masker (inference) helpers, NHWC storage ---------------------------------------------
 * nn.MaxPool2d(3, stride=2, padding=0, ceil_mode=True) (deeplab/resnetmulti_v2.py:76-78); caller passes ho/wo.
 * F.interpolate bilinear (deeplab_v2.py:116,198; generator.py:227) and bicubic (depth.py:144-149, align_corners=False).
 * torch.mean(z, dim=1, keepdim=True) (depth.py:142); z * z_depth (deeplab_v2.py:193, blocks.py:306).
 * OmniGenerator.make_m_cond (generator.py:196-230): cat[normalize(d), softmax(s), x resized]; mm = n*2 floats scratch. */
int cgb_maxpool3s2_ceil_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo,
                            int32_t c, void* stream);
int cgb_resize_bilinear_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo,
                            int32_t c, int32_t align_corners, void* stream);
int cgb_resize_bicubic_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo,
                           int32_t c, void* stream);
int cgb_channel_mean(const void* x, void* y, int32_t dtype, int64_t pixels, int32_t cs, int32_t c_logical, void* stream);
int cgb_mul(const void* a, const void* b, void* y, int32_t dtype, int64_t count, void* stream);
int cgb_make_m_cond(const void* d, const void* s, const void* xr, float* mm, void* out, int32_t dtype, int32_t n, int32_t hw,
                    int32_t ss, int32_t ns, int32_t cs_out, void* stream);
/* Adjoint of cgb_make_m_cond w.r.t. d and s (gen.m.spade.detach = false, defaults.yaml:182; generator.py:217-219): through
 * the per-sample min-max normalisation (first-occurrence argmin / argmax, as torch.min / max with dim) and the softmax.
 * d [n,hw,8], out / gout [n,hw,cs_out] (the forward's output and its gradient), mm = the forward's per-sample {min,max};
 * writes gd [n,hw,8] and gs [n,hw,ss] (pad channels zero). */
int cgb_make_m_cond_bwd(const void* d, const void* out, const float* mm, const void* gout, void* gd, void* gs, int32_t dtype,
                        int32_t n, int32_t hw, int32_t ss, int32_t ns, int32_t cs_out, void* stream);

/* ---- masker training path -----------------------------------------------------------------
 * nn.BatchNorm2d in TRAIN mode (batch statistics; resnetmulti_v2.py:16-18,30-34,72 freeze only weight/bias;
 * deeplab_v2.py:23,100,104,146; depth.py:57-105 via Conv2dBlock(norm="batch") blocks.py:95-96), fused with the ReLU /
 * LeakyReLU that follows and the bottleneck's residual add (resnetmulti_v2.py:50-52).  Statistics come from
 * cgb_instnorm_stats with n=1, hw=N*H*W.  weight/bias are per-channel fp32 (NULL = 1 / 0), padded to c.
 *   fwd:      y = act((x-mean)*rstd*weight + bias (+ residual))
 *   bwd (1):  gpre = gy*act'(y) (also the gradient of the residual) ; sums[c][0] = sum gpre (= gbias),
 *             sums[c][1] = sum gpre*xhat (= gweight)   (fp64; buffer of cgb_bn_bwd_ws_doubles(npix, c) doubles)
 *   bwd (2):  gx = weight*rstd*(gpre - sums0/M - xhat*sums1/M)
 *   running:  running_mean/var <- (1-momentum)*running + momentum*batch (unbiased variance), as F.batch_norm does. */
int cgb_bn_apply_fwd(const void* x, const float* mean, const float* rstd, const float* weight, const float* bias,
                     const void* residual, void* y, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope,
                     void* stream);
int cgb_bn_apply_bwd(const void* x, const float* mean, const float* rstd, const void* y, const void* gy, void* gpre,
                     double* sums, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope, void* stream);
/* doubles `sums` must hold for cgb_bn_apply_bwd / cgb_bn_train_bwd: sums[c][2] followed by per-chunk fp32 partials */
int64_t cgb_bn_bwd_ws_doubles(int64_t npix, int32_t c);
int cgb_bn_bwd_finalize(const void* x, const float* mean, const float* rstd, const float* weight, const double* sums,
                        const void* gpre, void* gx, int32_t dtype, int64_t npix, int32_t c, void* stream);
int cgb_bn_update_running(const float* mean, const float* rstd, float* running_mean, float* running_var, int32_t c,
                          int64_t count, float momentum, float eps, void* stream);
/* The same, one call per direction (three fewer host round trips per BatchNorm): train_fwd = statistics (ws: cgb_instnorm_ws_doubles(1,npix,c)
 * doubles of scratch) -> running update of the first c_logical channels, num_batches_tracked += 1 (both optional) -> apply;
 * train_bwd = part 1 then part 2 (gx NULL: part 1 only). */
int cgb_bn_train_fwd(const void* x, const float* weight, const float* bias, const void* residual, void* y, float* mean,
                     float* rstd, double* ws, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                     int32_t dtype, int64_t npix, int32_t c, int32_t c_logical, float momentum, float eps, int32_t act,
                     float slope, void* stream);
int cgb_bn_train_bwd(const void* x, const float* mean, const float* rstd, const float* weight, const void* y, const void* gy,
                     void* gpre, void* gx, double* sums, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope,
                     void* stream);
/* The same when the BatchNorm'd output fed TWO consumers (the ResNet bottleneck's output: the next block's conv1 and its identity
 * branch, resnetmulti_v2.py:40-56): gy and gy2 are the two incoming gradients, summed in fp32 inside the first pass instead of by
 * a separate pass over the tensor (autograd's accumulation: 66 bf16 adds of [8,80,80,1024] per train step).  gy2 may be NULL. */
int cgb_bn_train_bwd2(const void* x, const float* mean, const float* rstd, const float* weight, const void* y, const void* gy,
                      const void* gy2, void* gpre, void* gx, double* sums, int32_t dtype, int64_t npix, int32_t c, int32_t act,
                      float slope, void* stream);
/* cgb_bn_train_fwd without its statistics pass: mean / rstd (fp64 fold) and the running update come from the per-CTA partial
 * sums [rows][2][c] a cgb_conv2d_fwd_stats launch left behind, then the same apply pass. */
int cgb_bn_train_fwd_partials(const void* x, const float* partial, int32_t rows, const float* weight, const float* bias,
                              const void* residual, void* y, float* mean, float* rstd, float* running_mean, float* running_var,
                              int64_t* num_batches_tracked, int32_t dtype, int64_t npix, int32_t c, int32_t c_logical,
                              float momentum, float eps, int32_t act, float slope, void* stream);
/* adjoints of cgb_maxpool3s2_ceil_fwd (gradient to the first maximum of each window, as ATen), cgb_resize_bilinear_fwd,
 * cgb_channel_mean; nn.ReflectionPad2d (blocks.py:66-67) as an explicit copy + its fold-back adjoint so reflect-padded
 * convs run as pad-0 convs on the tcgen05 engine; dst[n,hw,c] = src[n,c]*scale (AdaptiveAvgPool2d(1) backward and the
 * 1x1 -> HxW bilinear of the ASPP image-pool branch, deeplab_v2.py:97-102,116); nn.Dropout (deeplab_v2.py:106,148,152)
 * with a counter-based mask from (seed, element index): the same call on the gradient is its backward. */
int cgb_maxpool3s2_ceil_bwd(const void* x, const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi,
                            int32_t ho, int32_t wo, int32_t c, void* stream);
/* nn.MaxPool2d(3, stride=2, padding=pad) for pad in {0,1} (deeplab/resnet101_v3.py:75: the v3 stem pools with padding 1);
 * the caller passes ho / wo (floor or ceil mode); the adjoint routes to the first maximum of each window. */
int cgb_maxpool3s2_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo, int32_t c,
                       int32_t pad, void* stream);
int cgb_maxpool3s2_bwd(const void* x, const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho,
                       int32_t wo, int32_t c, int32_t pad, void* stream);
int cgb_resize_bilinear_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho,
                            int32_t wo, int32_t c, int32_t align_corners, void* stream);
int cgb_reflect_pad_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad,
                        void* stream);
int cgb_reflect_pad_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad,
                        void* stream);
int cgb_channel_mean_bwd(const void* gy, void* gx, int32_t dtype, int64_t pixels, int32_t cs, int32_t c_logical, void* stream);
/* nn.ReplicationPad2d(pad) (Conv2dBlock pad_type="replicate", blocks.py:68-69) as an explicit copy, and its adjoint (every
 * border pixel gathers the gradients of the padding cells that replicate it; deterministic, no atomics). */
int cgb_replicate_pad_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad,
                          void* stream);
int cgb_replicate_pad_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad,
                          void* stream);
/* Per-(sample, channel) affine map + activation, the building block of the Conv2dBlock normalisations that are not on the
 * default path (norm = "instance" / "layer" / "adain", blocks.py:77-90; norms.py:8-49,51-81): the caller turns per-(n, c)
 * moments into scale / shift with a few [n, c]-sized tensor ops and this kernel applies them in one pass.
 *   fwd: y[n,p,ch] = act(x[n,p,ch] * scale[n,ch] + shift[n,ch])                       scale, shift: fp32 [n, c]
 *   bwd: gpre = gy * act'(y) ; gx = gpre * scale ; sums[n][ch][0] += sum_p gpre ; sums[n][ch][1] += sum_p gpre * x
 *        (sums fp64 [n, c, 2], zeroed by the caller: the gradients w.r.t. shift and scale). */
int cgb_affine_nc_fwd(const void* x, const float* scale, const float* shift, void* y, int32_t dtype, int32_t n, int32_t hw,
                      int32_t c, int32_t act, float slope, void* stream);
int cgb_affine_nc_bwd(const void* x, const void* y, const void* gy, const float* scale, void* gx, double* sums, int32_t dtype,
                      int32_t n, int32_t hw, int32_t c, int32_t act, float slope, void* stream);
int cgb_broadcast_hw(const void* src, void* dst, int32_t dtype, int32_t n, int32_t hw, int32_t c, float scale, void* stream);
int cgb_dropout(const void* x, void* y, int32_t dtype, int64_t count, float p, uint64_t seed, void* stream);
/* Same, the seed read from device memory at execution time: a launch captured in a CUDA graph draws a fresh mask on every
 * replay (the host writes this step's seed — torch's CPU generator, as above — into *seed_dev before the replay). */
int cgb_dropout_dev(const void* x, void* y, int32_t dtype, int64_t count, float p, const uint64_t* seed_dev, void* stream);

/* ---- masker losses (NCHW fp32, the layout Trainer.masker_{d,s,m}_loss receive; trainer.py:1389-1616) --------------
 * Each *_loss entry ADDS the mean-reduced loss to the device scalar `loss` (caller zeroes) and writes the gradient of
 * that loss w.r.t. its first argument (optional unless stated).
 *   softmax over dim 1 (trainer.py:1449,1475) fwd / bwd
 *   cross_entropy: nn.CrossEntropyLoss (losses.py:106-112), target int64 [n,h,w]
 *   entropy: prob_2_entropy (losses.py:466-471) optionally times a [n,1,h,w] depth map (ADVENTAdversarialLoss.__call__
 *            losses.py:541-543); backward=1 writes ge*d(entropy)/dp into out
 *   minent: MinentLoss v1 / v2 (losses.py:177-196); acc = 1 double scratch
 *   sigmoid_pair: cat[sigmoid(l), 1-sigmoid(l)] (trainer.py:1532-1534); backward=1 writes the logits gradient into out
 *   tv: TVLoss (losses.py:140-171) ; bce_logits: nn.BCEWithLogitsLoss with a tensor target (losses.py:419) ;
 *   ground_intersection: GroundIntersectionLoss (losses.py:449-455; piecewise constant, no gradient)
 *   sigm: SIGMLoss (losses.py:232-278): median/MAD alignment of prediction and target, 0.5/num_pix*sum|R| + Sobel gradient
 *         matching over `scales` nearest-downsampled scales; gpred required; ws = 16 + 2*n*h*w floats of scratch.
 *         The gradient flows through the median (shared among ties, as torch.median's backward) and the MAD scale. */
int cgb_softmax_nchw_fwd(const float* x, float* y, int32_t n, int32_t c, int32_t hw, void* stream);
int cgb_softmax_nchw_bwd(const float* y, const float* gy, float* gx, int32_t n, int32_t c, int32_t hw, void* stream);
int cgb_cross_entropy_nchw(const float* logits, const int64_t* target, float* loss, float* glogits, int32_t n, int32_t c,
                           int32_t hw, void* stream);
int cgb_entropy_nchw(const float* p, const float* depth, const float* ge, float* out, int32_t n, int32_t c, int32_t hw,
                     int32_t backward, void* stream);
int cgb_minent_loss(const float* p, float* loss, float* gp, double* acc, int32_t n, int32_t c, int32_t hw, int32_t version,
                    float lambda_var, void* stream);
int cgb_sigmoid_pair(const float* logits, const float* gprob, float* out, int32_t n, int32_t hw, int32_t backward, void* stream);
int cgb_tv_loss(const float* x, float* loss, float* gx, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);
int cgb_bce_logits_loss(const float* x, const float* target, float* loss, float* gx, int64_t count, void* stream);
int cgb_ground_intersection_loss(const float* pred, const float* ground, float* loss, int64_t count, void* stream);
/* DADADepthLoss (losses.py:596-620; gen.d.loss = "dada"): reverse Huber on |pred - label| with threshold c = 0.2 * max over the
 * whole batch (a constant of the graph: the reference takes it with .item()); adds the mean to `loss`, writes dloss/dpred. */
int cgb_dada_depth_loss(const float* pred, const float* label, float* loss, float* gpred, int64_t count, void* stream);
int cgb_sigm_loss(const float* pred, const float* target, float* loss, float* gpred, float* ws, int32_t n, int32_t h, int32_t w,
                  float gmweight, int32_t scales, void* stream);

/* ---- differentiable augmentation (DiffTransforms, climategan/transforms.py:493-626; gen.p.diff_aug, trainer.py:1079-1081,
 *      1319-1321) — NCHW fp32 images, 1 <= c <= 8 ---------------------------------------------------------------------------
 * params [n][8] fp32 on the device, one row per sample: {b, cf, sf, tx, ty, ox, oy, unused} = brightness offset (rand - 0.5),
 *   contrast factor (rand + 0.5), saturation factor (2 rand), row / column translation, cutout offsets (integers stored as floats):
 *     v1 = x + b ; v2 = (v1 - mean_all(v1)) cf + mean_all(v1) ; v3 = (v2 - mean_ch(v2)) sf + mean_ch(v2)
 *     y(i, j) = v3(i + tx, j + ty), zero outside the image ; y = 0 in rows [ox - cut_h/2, +cut_h) x columns [oy - cut_w/2, +cut_w)
 *   clamped into the image (cut_h = cut_w = 0: no cutout; b = 0, cf = sf = 1: no jitter; tx = ty = 0: no translation).
 * diff_aug_sum ACCUMULATES per-sample fp64 totals into sums[n] (caller zeroes): mode 0 of x (forward: mean_all), mode 1 of the
 *   output gradient over the output pixels that are outside the cutout and whose source pixel is inside the image (backward).
 * diff_aug_fwd: y from x, params and the mode-0 sums.  diff_aug_bwd: dL/dx from dL/dy, params and the mode-1 sums. */
int cgb_diff_aug_sum(const float* t, const float* params, double* sums, int32_t n, int32_t c, int32_t h, int32_t w,
                     int32_t cut_h, int32_t cut_w, int32_t mode, void* stream);
int cgb_diff_aug_fwd(const float* x, const float* params, const double* sums, float* y, int32_t n, int32_t c, int32_t h,
                     int32_t w, int32_t cut_h, int32_t cut_w, void* stream);
int cgb_diff_aug_bwd(const float* gy, const float* params, const double* gsums, float* gx, int32_t n, int32_t c, int32_t h,
                     int32_t w, int32_t cut_h, int32_t cut_w, void* stream);

/* ---- validation metrics (Trainer.eval_images trainer.py:1706-1799; accuracy / mIOU climategan/eval_metrics.py:68-130) --------
 * argmax_confusion: conf[p][l] += 1 per pixel with p = argmax over the class axis of logits [n][c][hw] fp32 (first maximum, NaN
 *   counts as the maximum, like torch.argmax) and l = label [n][hw] int64, labels outside [0, c) counted in the extra column c.
 *   conf is int64 [c][c+1] and ACCUMULATES (caller zeroes it); label_max is one int64, atomically raised to the largest label seen
 *   (caller initialises it to INT64_MIN) — mIOU's two-class case scores only class label.max() (eval_metrics.py:105-107).
 *   1 <= c <= 64.  Both metrics are integer functions of conf: bit-exact against the reference. */
int cgb_argmax_confusion(const float* logits, const int64_t* label, int64_t* conf, int64_t* label_max, int32_t n, int32_t c,
                         int64_t hw, void* stream);

/* ---- inference events (Trainer.infer_all, trainer.py:218-334) — NCHW fp32 images at the API edge ---------------------
 * minmax_per_sample: per-sample min/max (tutils.normalize :567-576) -> mm[n][2].
 * fire (climategan/fire.py:68-127): fire_tone = normalize(x,0,255), warm (+40,-10,-20), clamp, uint8, adjust_contrast(c),
 *   adjust_brightness(b) with torchvision's uint8 truncation (gray_sum: n doubles scratch); sky_mask = argmax(seg)==sky_idx with
 *   the bottom third cropped (:95-97); plane_resize_nearest = F.interpolate of the mask (:99-102); box_dilate =
 *   increase_sky_mask (:15-47; on a binary mask the shifted sums + clamp are a box dilation, radius int(p*size)-1);
 *   gauss_blur = kornia filter2d with get_gaussian_kernel2d (:104-111; kornia 0.5.10 is not vendored — restated: outer product
 *   of two normalised 1-D Gaussians, reflect border), run as two 1-D passes; fire_paste = paste_tensor with the (255,g,0)
 *   filter at transparency/255, uint8, adjust_brightness, the two "dummy" corner pixels (:113-125).
 * smog (trainer.py:1879-1939): HazeRD transmission exp(-beta*d) on srgb2lrgb(normalize(x)) with d = normalize(1/normalize(d,
 *   .3,1),.1,1) bilinearly upsampled (align_corners), airlight, lrgb2srgb, yellow filter (alpha and colour already /255).
 * to_uint8_nhwc: normalize(t) -> NHWC -> (t*255).astype(uint8) (trainer.py:312-327); mask_to_uint8: (m > bin)*255 (:330-332). */
int cgb_minmax_per_sample(const float* x, float* mm, int32_t n, int64_t count, void* stream);
int cgb_fire_tone(const float* x, const float* mm, float* out, double* gray_sum, int32_t n, int32_t hw, float contrast,
                  float brightness, void* stream);
int cgb_sky_mask(const float* seg, float* out, int32_t n, int32_t c, int32_t hs, int32_t ws, int32_t sky_idx, int32_t crop_bottom,
                 void* stream);
int cgb_plane_resize_nearest(const float* x, float* y, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo, void* stream);
int cgb_box_dilate(const float* x, float* tmp, float* y, int32_t n, int32_t h, int32_t w, int32_t radius_w, int32_t radius_h,
                   void* stream);
int cgb_gauss_blur(const float* x, float* tmp, float* y, int32_t n, int32_t h, int32_t w, int32_t ksize, float sigma, void* stream);
int cgb_fire_paste(const float* img, const float* sky, float* out, int32_t n, int32_t h, int32_t w, float fr, float fg, float fb,
                   float transparency, float brightness, void* stream);
int cgb_smog(const float* x, const float* mmx, const float* d, const float* mmd, float* out, int32_t n, int32_t h, int32_t w,
             int32_t hd, int32_t wd, float airlight, float beta, float alpha, float yr, float yg, float yb, void* stream);
/* paint_cloudy (generator.py:299-328): perlin_noise = rand_perlin_2d (tutils.py:648-686) from (res0+1)x(res1+1) random angles
 * drawn by the caller; cloudy_mix = sky mask from argmax of the bilinearly upsampled seg logits == sky_idx, then mix_noise
 * (tutils.py:689-694) with mm_noise = per-sample (min,max) of the noise plane. */
int cgb_perlin_noise(const float* angles, float* out, int32_t h, int32_t w, int32_t res0, int32_t res1, void* stream);
int cgb_cloudy_mix(const float* x, const float* seg, const float* noise, const float* mm_noise, float* out, int32_t n, int32_t h,
                   int32_t w, int32_t c, int32_t hs, int32_t ws, int32_t sky_idx, float weight, void* stream);
int cgb_to_uint8_nhwc(const float* x, const float* mm, uint8_t* out, int32_t n, int32_t hw, void* stream);
/* Input edge (apply_events.py resize_and_crop :211-241 + to_m1_p1 :179-195; transforms.PrepareInference :292-360, skimage on the
 * CPU in the reference): src uint8 [h, w, 3] -> anti-aliased bilinear resize to (rh, rw) -> crop (th, tw) at (top, left) ->
 * optional truncation to uint8 (:231) -> (v/255 - 0.5)*2 into dst fp32 [3, th, tw] (one image slot of an NCHW batch). */
int cgb_resize_crop_u8(const uint8_t* src, float* dst, int32_t h, int32_t w, int32_t rh, int32_t rw, int32_t top, int32_t left,
                       int32_t th, int32_t tw, int32_t quantize, void* stream);
int cgb_mask_to_uint8(const float* m, uint8_t* out, float bin_value, int64_t count, void* stream);

/* ---- layout / elementwise ----------------------------------------------------------------
 * NCHW fp32 (the reference's tensor layout at the API edge) <-> NHWC storage. */
int cgb_nchw_to_nhwc(const float* x, void* y, int32_t dtype, int32_t n, int32_t c, int32_t hw,
                     int32_t cs, void* stream);
int cgb_nhwc_to_nchw(const void* x, float* y, int32_t dtype, int32_t n, int32_t c, int32_t hw,
                     int32_t cs, void* stream);
/* gx = gy * act'(y)  (y = activation output) */
int cgb_act_bwd(const void* gy, const void* y, void* gx, int32_t dtype, int64_t count, int32_t act,
                float slope, void* stream);
/* The same on a [npix, c] tensor, also returning gbias[c] = sum over pixels of gx (fp32; zeroed here): the bias gradient of a
 * Conv2dBlock's conv (blocks.py:138-144, nn.Conv2d(..., bias=True) -> activation), so the conv's weight-gradient launch needs
 * no column-sum pass over gx. */
int cgb_act_bwd_bias(const void* gy, const void* y, void* gx, float* gbias, int32_t dtype, int64_t npix, int32_t c,
                     int32_t act, float slope, void* stream);
/* y = act(x) elementwise (F.leaky_relu before conv_img, painter.py:166) */
int cgb_act_fwd(const void* x, void* y, int32_t dtype, int64_t count, int32_t act, float slope,
                void* stream);

/* ---- compositing -------------------------------------------------------------------------
 * OmniGenerator.paint (climategan/generator.py:279-297), NCHW fp32 at the API edge:
 *   cond = x*(1-m)                       -> NHWC storage (painter input)
 *   out  = x*(1-m) + fake*m              (paste_original_content)
 * m is [n,1,h,w]; x, fake, out are [n,3,h,w] fp32. */
int cgb_mask_cond(const float* x, const float* m, void* cond, int32_t dtype, int32_t n, int32_t hw,
                  int32_t cs, void* stream);
int cgb_paste_fwd(const float* x, const float* m, const float* fake, float* out, int32_t n, int32_t hw,
                  void* stream);
/* gfake = gout * m */
int cgb_paste_bwd(const float* gout, const float* m, float* gfake, int32_t n, int32_t hw, void* stream);
/* Adjoints of the two above w.r.t. the MASK, for the painter loss of the masker (Trainer.painter_loss_for_masker,
 * trainer.py:1618-1651: the mask is the masker's own prediction): gm = -sum_c x_c gcond_c and gm = sum_c gout_c (fake_c - x_c);
 * gm fp32 [n,1,hw]. */
int cgb_mask_cond_bwd(const float* x, const void* gcond, float* gm, int32_t dtype, int32_t n, int32_t hw, int32_t cs,
                      void* stream);
int cgb_paste_bwd_mask(const float* gout, const float* x, const float* fake, float* gm, int32_t n, int32_t hw, void* stream);

/* ---- losses ------------------------------------------------------------------------------
 * mean |a-b| and its gradient w.r.t. a (nn.L1Loss; climategan/losses.py:290-301, FeatMatchLoss :86-103):
 *   loss[0] += scale * sum|a-b| ;  ga = scale * sign(a-b)   (ga optional)
 * `loss` is a device fp32 scalar the caller zeroes. */
int cgb_l1_loss(const float* a, const float* b, float* loss, float* ga, int64_t count, float scale,
                void* stream);

/* Mean-reduced losses against a constant target on fp32 arrays — GANLoss (BCEWithLogits / MSE, losses.py:13-83) and
 * HingeLoss (losses.py:550-593):  kind 0 BCE-with-logits, 1 MSE, 2 hinge-D-real, 3 hinge-D-fake, 4 -mean(x).
 *   loss[0] += scale*sum(l_i) ; gx_i = scale*dl_i/dx_i  (gx optional; scale = weight/count). */
int cgb_const_target_loss(const float* x, float* loss, float* gx, int64_t count, int32_t kind, float target, float scale,
                          void* stream);
/* Same, the target read from device memory at execution time (GANLoss's soft-label / flip draws, losses.py:52-70, change every
 * step: a launch captured in a CUDA graph must not bake one draw in). */
int cgb_const_target_loss_dev(const float* x, float* loss, float* gx, int64_t count, int32_t kind, const float* target_dev,
                              float scale, void* stream);
/* L1 between two storage tensors (FeatMatchLoss, losses.py:86-103): loss[0] += scale*sum|a-b| ; ga = scale*sign(a-b). */
int cgb_l1_loss_storage(const void* a, const void* b, float* loss, void* ga, int32_t dtype, int64_t count, float scale,
                        void* stream);

/* VGG perceptual-loss edge (climategan/losses.py:304-350, tutils.py:416-427):
 *   vgg_preprocess(img * m): RGB->BGR, [-1,1]->[0,255], minus the caffe channel means; NCHW fp32 (+ optional [n,1,h,w]
 *   mask) -> NHWC storage with 8 channels.  _bwd is its adjoint (gradient w.r.t. img).
 *   nn.MaxPool2d(2,2) of torchvision's vgg19.features, NHWC (bwd routes to the first maximum, like ATen). */
int cgb_vgg_preprocess_fwd(const float* x, const float* m, void* y, int32_t dtype, int32_t n, int32_t hw, void* stream);
int cgb_vgg_preprocess_bwd(const void* gy, const float* m, float* gx, int32_t dtype, int32_t n, int32_t hw, void* stream);
int cgb_maxpool2_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t c, void* stream);
int cgb_maxpool2_bwd(const void* x, const void* y, const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi,
                     int32_t wi, int32_t c, void* stream);

/* ---- optimiser ---------------------------------------------------------------------------
 * ExtraAdam (climategan/optim.py:137-291; Trainer.g_opt_step/d_opt_step, trainer.py:674-694) over flat fp32 arrays:
 *   m,v Adam moments; u = -lr*sqrt(1-b2^step)/(1-b1^step) * m/(sqrt(v)+eps)
 *   mode 0 = extrapolation: (save_copy ? c = p : -) ; p += u        mode 1 = step: p = c + u
 * `step` is the per-parameter Adam step count AFTER this call's increment (the reference increments it in both modes). */
int cgb_extra_adam(float* p, const float* g, float* m, float* v, float* c, int64_t count, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int32_t step, int32_t mode, int32_t save_copy,
                   void* stream);

/* ---- spectral norm -----------------------------------------------------------------------
 * SpectralNorm._update_u_v (climategan/norms.py:100-112), one power iteration:
 *   v <- normalize(W^T u) ; u <- normalize(W v) ; sigma = u.(W v)
 * W is w_bar viewed [rows, cols] fp32 row-major; u[rows], v[cols] updated in place; sigma -> device scalar.
 * Three small multi-CTA kernels (W <= 640x5760). */
int cgb_spectral_power_iter(const float* w, float* u, float* v, float* sigma, int32_t rows,
                            int32_t cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGB200_H */
