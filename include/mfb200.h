/* mfb200.h — C ABI of libmfb200.so: the B200 (sm_100a) kernels of the MirrorFusion denoising hot path.
 *
 * The reference (val-iisc/Reflecting-Reality, a diffusers 0.27 fork) has no FFI: its "backend" is
 * torch.nn.functional.  Each entry point below therefore names the reference call site(s) it replaces
 * (S/ = MirrorFusion/src/diffusers/).  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated; the caller owns all buffers
 *     and the stream (a cudaStream_t passed as void*); the library owns only plans it returns;
 *   - activations are NHWC (channels-last) bf16, i.e. [B, H*W, C] row-major token matrices;
 *   - every function returns MFB_OK (0) or a negative MFB_E* code; mfb_last_error() gives the message;
 *   - there is NO CPU fallback: mfb_init() fails on anything but compute capability 10.x.
 */
#ifndef MFB200_H
#define MFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFB_ABI_VERSION 1
#define MFB_OK 0
#define MFB_EINVAL (-1)       /* bad argument / unsupported shape */
#define MFB_ECUDA (-2)        /* a CUDA runtime/driver call failed */
#define MFB_EUNSUPPORTED (-3) /* device is not sm_100 */

int mfb_abi_version(void);
/* Select + validate the device, resolve cuTensorMapEncodeTiled. Idempotent. */
int mfb_init(int device);
const char* mfb_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear (tcgen05.mma + TMEM accumulators + TMA operand tiles).
 * Replaces F.conv2d / F.linear behind LoRACompatibleConv/Linear (S/models/lora.py:363-377,445-451) for:
 *   ResnetBlock2D conv1/conv2/conv_shortcut (S/models/resnet.py:367,396,398-401), Downsample2D
 *   (S/models/downsampling.py:146-152), Upsample2D.conv (S/models/upsampling.py:179-184), Transformer2DModel
 *   proj_in/proj_out (S/models/transformers/transformer_2d.py:340-344,417-421), Attention to_q/k/v/out
 *   (S/models/attention_processor.py:1246-1274), FeedForward/GEGLU (S/models/attention.py:668-675,
 *   S/models/activations.py:100-103), BrushNet zero-convs (S/models/brushnet.py:832-834,851,891-893).
 *
 * (mfb_plan_flops reports ALGORITHMIC work: 2*M*N*Ktot of the convolution as the reference computes it.)
 *
 *   out[b,oh,ow,:] = ( sum_taps x[b, oh*s+kh-pad, ow*s+kw-pad, :] . w_tap  +  sum_e extra_x[e][b,oh,ow,:] . w_e
 *                      + bias + rowbias[b,:] ) * alpha  + res1[b,oh,ow,:] + res2[b,oh,ow,:]
 *
 * w is the packed weight [Cout, Ktot] bf16, K order = (kh, kw, cin) then the extra segments in order.
 * A linear layer over an [M, K] token matrix is ksize=1, B=1, H=1, W=M, Cin=K.
 * With geglu=1 the packed rows are interleaved per 128: 64 value rows then the 64 matching gate rows, and
 * out is [M, Cout/2] = value * gelu_erf(gate).
 */
typedef struct mfb_conv_desc {
    int B, H, W;          /* input geometry (NHWC) */
    int Cin, Cout;        /* Cin % 64 == 0, Cout % 8 == 0 */
    int ksize;            /* 1 or 3 (padding = ksize/2) */
    int stride;           /* 1 or 2 (2 only with ksize 3) */
    const void* x;        /* [B,H,W,Cin] bf16 */
    int n_extra;          /* 0..3 extra 1x1 K-segments at OUTPUT resolution (shortcut over concat halves, ...) */
    const void* extra_x[3];
    int extra_C[3];
    const void* w;        /* [Cout, Ktot] bf16 */
    const float* bias;    /* [Cout] fp32 or NULL */
    const float* rowbias; /* [B, rowbias_ld] fp32 or NULL: time_emb_proj(silu(emb)) (S/models/resnet.py:369-379) */
    int rowbias_ld;
    const float* alpha;   /* device scalar or NULL (=1): BrushNet conditioning_scale (S/models/brushnet.py:904-906) */
    const void* res1;     /* [B,Ho,Wo,Cout] bf16 or NULL: residual / identity shortcut */
    const void* res2;     /* second residual: the BrushNet tap (S/models/unets/unet_2d_blocks.py:1388-1398 ...) */
    void* out;            /* [B,Ho,Wo,Cout] bf16 ([.., Cout/2] with geglu) */
    int geglu;
    int block_n;          /* 0 = auto, else 64, 80, 128 or 160 */
    int up2x;             /* 1: the conv runs over the nearest-2x upsample of x (Upsample2D, S/models/upsampling.py:167-184).
                             B,H,W are the LOW-resolution input dims, out/res/extras are [B,2H,2W,.]; w holds the four
                             sub-pixel phases [4][Cout][4*Cin + extras] (phase = py*2+px, taps that hit the same source
                             pixel pre-summed); the plan issues 4 launches and never materialises the upsampled tensor */
    int igemm_mode;       /* 0 = auto (env MFB_IGEMM_MODE / MFB_IGEMM_SPLITK or independent CTAs); 1 independent CTAs, 2 CTA pair + weight
                             multicast, 3 CTA pair + cta_group::2 UMMA (256-row tile), 4 split-K CTA pair where the geometry qualifies
                             (few-tile launches: both CTAs of a pair work on one wide tile, each on half of the K blocks, partial
                             accumulator handed over through distributed shared memory), independent CTAs otherwise */
    int pad0;             /* stride 2 only.  0: conv padding 1 (UNet Downsample2D).  1: F.pad(x, (0,1,0,1)) + conv padding 0 —
                             the VAE encoder's Downsample2D(padding=0) (S/models/downsampling.py:141-143): taps read input rows
                             2*oh + kh (not 2*oh + kh - 1), the zero row / column sits at the bottom / right edge */
    int dtype;            /* 0 = bf16 tensors, tcgen05 path (the product).  1 = fp32 PARITY MODE: x / extras / w / res / out
                             are fp32, CUDA-core FFMA accumulation (csrc/fp32mode.cu) — same descriptor semantics */
} mfb_conv_desc;

/* ------------------------------------------------------------------------------------------------------------
 * Recorded launch programs: the model-level entry point.
 * Between mfb_program_begin and mfb_program_end every step-level call made ON THE SAME THREAD (mfb_plan_run, mfb_groupnorm[_prestat],
 * mfb_layernorm, mfb_attention, mfb_conv_in / _out, mfb_cfg_sched_step, mfb_timestep_sinusoid, mfb_linear_small, the layout / dtype
 * conversions, mfb_softmax_rows, mfb_transpose_tokens, mfb_latent_sample, mfb_copy_f32 and their *_f32 parity-mode forms) is
 * executed AND appended to the program with its arguments; mfb_program_run replays the whole sequence on a stream from ONE call.
 * Recorded once over StepEngine's step it is `mfb_step_fused` of SURVEY.md §8b: latents -> BrushNetModel.forward -> UNet2DConditionModel
 * .forward with the 28 taps -> CFG combine -> scheduler step (pipeline_brushnet.py:1250-1315), 512 launches, no host language in the
 * loop; per-step inputs (latents, the 12 scheduler coefficients, the timestep's row biases, the tap scales) live in fixed device
 * buffers the caller rewrites between runs.  Everything a recorded call references (buffers, plans) must outlive the program. */
typedef struct mfb_program mfb_program;
int mfb_program_begin(mfb_program** out);
int mfb_program_end(void);
int mfb_program_size(const mfb_program* prog);            /* recorded calls */
int mfb_program_run(mfb_program* prog, void* stream);
int mfb_program_destroy(mfb_program* prog);
int mfb_copy_f32(float* dst, const float* src, long long n, void* stream);   /* dst[i] = src[i]: the `torch.cat([latents] * 2)` of the loop */

typedef struct mfb_plan mfb_plan;
int mfb_conv_plan_create(const mfb_conv_desc* desc, mfb_plan** out);
int mfb_plan_run(mfb_plan* plan, void* stream);
int mfb_plan_destroy(mfb_plan* plan);
double mfb_plan_flops(const mfb_plan* plan); /* 2*M*N*Ktot */
int mfb_plan_ktotal(const mfb_plan* plan);
int mfb_plan_igemm_mode(const mfb_plan* plan); /* kernel variant the plan chose: 0 independent CTAs, 1 / 2 the pair modes, 3 split-K pair */
int mfb_plan_launches(const mfb_plan* plan); /* kernel launches per mfb_plan_run (4 for up2x plans) */
/* Fused GroupNorm statistics: the epilogue can also emit, per (image, tile, output channel), the sum and sum of squares
 * of the bf16 values it stores — the statistics F.group_norm of the CONSUMER would otherwise re-read the tensor for
 * (S/models/resnet.py:337,381).  mfb_plan_stats_floats = size of the [B][tiles][Cout][2] fp32 buffer to provide
 * (0: this plan cannot — GEGLU, or tiles that straddle images); mfb_plan_set_stats installs it (NULL = off). */
long long mfb_plan_stats_floats(const mfb_plan* plan);
int mfb_plan_stats_tiles(const mfb_plan* plan);
int mfb_plan_set_stats(mfb_plan* plan, float* buf);

/* ------------------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) over NHWC bf16, fp32 statistics.  Replaces F.group_norm + F.silu in ResnetBlock2D
 * (S/models/resnet.py:337-338,381,393), conv_norm_out (S/models/unets/unet_2d_condition.py:1337-1338) and the
 * GroupNorm of Transformer2DModel (transformer_2d.py:338, eps 1e-6, silu=0).  x2 (optional) is a second tensor
 * concatenated after x1 along channels — the skip concat of the up blocks (unet_2d_blocks.py:2586,2728) — so the
 * concatenated tensor is only ever materialised normalised.  stats_ws: MFB_GN_WS_FLOATS(B, groups) floats of scratch,
 * ZERO-INITIALISED once by the caller (it holds self-resetting ticket counters); the reduction is deterministic.
 */
#define MFB_GN_MAX_CHUNKS 64
#define MFB_GN_WS_FLOATS(B, groups) (2 * (B) * (groups) * (1 + MFB_GN_MAX_CHUNKS) + (B))
int mfb_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, int groups, float eps,
                  const float* gamma, const float* beta, int silu, float* stats_ws, void* out, void* stream);

/* Same, with the statistics of each source already available as igemm partials (mfb_plan_set_stats): the full-tensor
 * statistics pass is replaced by a tiny fixed-order reduction. */
int mfb_groupnorm_prestat(const void* x1, int C1, const float* part1, int tiles1, const void* x2, int C2, const float* part2,
                          int tiles2, int B, int HW, int groups, float eps, const float* gamma, const float* beta, int silu,
                          float* stats_ws, void* out, void* stream);

/* LayerNorm over the last dim of [rows, C] bf16 (S/models/attention.py:313,360,386). */
int mfb_layernorm(const void* x, int rows, int C, float eps, const float* gamma, const float* beta, void* out,
                  void* stream);

/* Row softmax over [rows, cols] bf16 (in place allowed): the VAE mid block's single-head d = 512 attention runs as two
 * GEMMs around it (S/models/attention_processor.py:1266-1268). */
int mfb_softmax_rows(const void* x, int rows, int cols, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Scaled-dot-product attention, flash style on tcgen05 (S in TMEM, online softmax, P through shared memory).
 * Replaces F.scaled_dot_product_attention in AttnProcessor2_0.__call__
 * (S/models/attention_processor.py:1266-1268): no mask, non-causal, scale = head_dim^-0.5.
 *   q  : [B, Tq, ldq]  bf16, head h occupies columns [h*d, (h+1)*d)
 *   k  : [B, Tk, ldk]  bf16, same column convention
 *   v  : [B, Tk, ldv] bf16 (head h at columns [h*d, (h+1)*d)) — consumed as stored (MN-major tensor-core operand)
 *   out: [B, Tq, ldo]  bf16
 * head_dim in {40, 80, 160} (SD1.5: 320/640/1280 channels over 8 heads) or any multiple of 16 <= 160.
 */
int mfb_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B,
                  int heads, int head_dim, int Tq, int Tk, void* stream);
/* [B, T, ld] column block [col0, col0+C) -> transposed [B, C, ldt] (ldt >= T); pads [T, ldt) with zeros. */
int mfb_transpose_tokens(const void* x, int ld, int col0, int C, int B, int T, void* out, int ldt, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Boundary / small layers.
 */
/* conv_in (S/models/unets/unet_2d_condition.py:1182) and conv_in_condition over cat([sample, brushnet_cond])
 * (S/models/brushnet.py:810-811): NCHW fp32 inputs (sample [B,Ca,H,W], cond [B,Cb,H,W] or NULL) -> NHWC bf16
 * [B,H,W,Cout].  w: [3,3,Ca+Cb,Cout] fp32.  If tap != NULL also writes out_post = out + tap
 * (the pre-tap tensor stays the first skip: unet_2d_condition.py:1215-1218). */
int mfb_conv_in(const float* sample, int Ca, const float* cond, int Cb, int B, int H, int W, const float* w,
                const float* bias, int Cout, void* out, const void* tap, void* out_post, void* stream);
/* conv_out (unet_2d_condition.py:1339): NHWC bf16 [B,H,W,Cin] -> NCHW fp32 [B,Cout<=4,H,W]; w [Cout,3,3,Cin] fp32. */
int mfb_conv_out(const void* x, int Cin, int B, int H, int W, const float* w, const float* bias, int Cout, float* out,
                 void* stream);
/* nearest x2 upsample of NHWC bf16 (S/models/upsampling.py:167-173). */
int mfb_upsample2x(const void* x, int B, int H, int W, int C, void* out, void* stream);
/* layout conversion at the API boundary */
int mfb_nchw_f32_to_nhwc_bf16(const float* x, int B, int C, int H, int W, void* out, void* stream);
int mfb_nhwc_bf16_to_nchw_f32(const void* x, int B, int C, int H, int W, float* out, void* stream);
int mfb_f32_to_bf16(const float* x, long long n, void* out, void* stream);

/* Timestep path (S/models/embeddings.py:27-67,226-237; S/models/resnet.py:369-376): all in fp32.
 * sinusoid: t [M] (fp32) -> [M, dim] = [cos | sin] (flip_sin_to_cos=True, shift 0). */
int mfb_timestep_sinusoid(const float* t, int M, int dim, float* out, void* stream);
/* y[M,N] = act_out( W[N,K](bf16) . act_in(x[M,K]) + b[N] ), fp32 in/out, M small. act: 0 none, 1 SiLU. */
int mfb_linear_small(const float* x, int M, int K, const void* w, const float* b, int N, int act_in, int act_out,
                     float* y, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * CFG combine + scheduler step fused (S/pipelines/brushnet/pipeline_brushnet.py:1310-1315 with
 * S/schedulers/scheduling_unipc_multistep.py:425,567-572,703-709 / S/schedulers/scheduling_ddim.py:404-450).
 * eps_uncond / eps_cond: the two halves of the [2*Bimg, n] fp32 noise prediction (uncond half first in the
 * reference batch; pass the same pointer twice with g = 0 for an already-guided prediction).  All tensors fp32
 * [Bimg, n].  coef: device array of 12 floats:
 *   g, c_x, c_eps                      : m_t   = c_x*x + c_eps*eps_guided           (x0-prediction)
 *   a_last, a_m0, a_m1, a_mt, use_corr : x_c   = a_last*last + a_m0*m0 + a_m1*m1 + a_mt*m_t   (UniC; skipped if use_corr==0)
 *   b_x, b_mt, b_m0, b_eps             : x_new = b_x*x_c + b_mt*m_t + b_m0*m0 + b_eps*eps_guided  (UniP / DDIM)
 * Writes x_new -> x (in place), x_c -> last, shifts m0 -> m1 and m_t -> m0.
 */
/* SynMirror input preprocessing of the eval sweep, inputs already at the target resolution:
 *   mfb_prep_image_u8   : uint8 HWC RGB [N,H,W,3] -> fp32 NCHW in [-1,1]   (VaeImageProcessor.preprocess, S/image_processor.py:446-530)
 *   mfb_prep_mask_depth : uint8 mask [N,H,W] (+ metric depth fp32 [N,H,W] or NULL) -> latent-resolution mask {0,1} and depth
 *                         in [-1,1] by nearest sampling at (factor*i, factor*j)  (pipeline_brushnet.py:1139,1190-1202;
 *                         E/dataset/dataset.py:131-145 with max_scene_depth = max depth over mask>0 + delta); scratch: N ints
 *   mfb_post_image_u8   : fp32 NCHW in [-1,1] -> uint8 HWC                  (VaeImageProcessor.postprocess) */
int mfb_prep_image_u8(const void* rgb_hwc, int N, int H, int W, float* out_nchw, void* stream);
int mfb_prep_mask_depth(const void* mask_u8, const float* depth, int N, int H, int W, int factor, float delta, float* mask_lat,
                        float* depth_lat, int* scratch_n_ints, void* stream);
int mfb_post_image_u8(const float* img_nchw, int N, int H, int W, void* out_hwc, void* stream);

/* Inputs that are NOT at the target resolution: the dataset's `transforms.Resize(res, BICUBIC)` + `CenterCrop(res)` on float
 * tensors (E/dataset/dataset.py:70-76 RGB, :86-92 mask, :155-165 normalised depth; torchvision resizes tensors with
 * F.interpolate(mode="bicubic", align_corners=False, antialias=True), i.e. ATen's separable Keys a = -0.5 filter whose support
 * grows with the down-scale factor).
 *   mfb_resize_crop_bicubic : fp32 [NC, Hs, Ws] planes -> [NC, res/step, res/step]: shorter side resized to `res` (longer:
 *                             int(res * long / short)), centre crop res x res (offsets int(round(d / 2)), half to even), and
 *                             only the crop's pixels (step*i, step*j) evaluated — step = 1 is the transform itself, step = 8
 *                             fuses the nearest sampling to latent resolution the pipeline applies to depth (:1196-1199).
 *   mfb_depth_normalize     : metric depth fp32 [N,H,W] + uint8 mask [N,H,W] -> 2 * clip(d, 0, dmax) / dmax - 1 with
 *                             dmax = max depth over mask > 0 + delta, at the map's own resolution (:131-145); scratch: N ints */
int mfb_resize_crop_bicubic(const float* in, int NC, int Hs, int Ws, int res, int step, float* out, void* stream);
int mfb_depth_normalize(const float* depth, const void* mask_u8, int N, int H, int W, float delta, float* out,
                        int* scratch_n_ints, void* stream);

/* Latent sample of DiagonalGaussianDistribution (S/models/autoencoders/vae.py:769-791) times a scale:
 * out = scale * (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise); noise == NULL gives the mode (scale * mean).  fp32, n elements. */
int mfb_latent_sample(const float* mean, const float* logvar, const float* noise, float scale, float* out, long long n,
                      void* stream);

int mfb_cfg_sched_step(const float* eps_uncond, const float* eps_cond, float* x, float* last, float* m0, float* m1,
                       const float* coef, int Bimg, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 parity mode (BASELINE.json configs[0]; north_star: per-step noise prediction within rel-L2 1e-4 in fp32 mode).
 * Same argument meaning as the bf16 entry points above with every bf16 tensor replaced by fp32 (conv / linear: set
 * mfb_conv_desc.dtype = 1).  Plain CUDA-core kernels: a correctness instrument for the shared host program, not a
 * performance path.
 */
int mfb_groupnorm_f32(const float* x1, int C1, const float* x2, int C2, int B, int HW, int groups, float eps,
                      const float* gamma, const float* beta, int silu, float* out, void* stream);
int mfb_layernorm_f32(const float* x, int rows, int C, float eps, const float* gamma, const float* beta, float* out,
                      void* stream);
int mfb_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                      int B, int heads, int head_dim, int Tq, int Tk, void* stream);
int mfb_conv_in_f32(const float* sample, int Ca, const float* cond, int Cb, int B, int H, int W, const float* w,
                    const float* bias, int Cout, float* out, const float* tap, float* out_post, void* stream);
int mfb_conv_out_f32(const float* x, int Cin, int B, int H, int W, const float* w, const float* bias, int Cout, float* out,
                     void* stream);
int mfb_linear_small_f32(const float* x, int M, int K, const float* w, const float* b, int N, int act_in, int act_out,
                         float* y, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training-step glue of the BrushNet fine-tune step (BASELINE config 4): the part of
 * E/train_brushnet_mirror.py:1404-1466 around the two nets.  fp32 tensors; every reduction is fixed-order (deterministic).
 */
/* DDPMScheduler.add_noise / get_velocity (S/schedulers/scheduling_ddpm.py:501-546): per sample b with a = alphas_cumprod[t_b]
 *   noisy = sqrt(a) x0 + sqrt(1-a) noise ;  velocity = sqrt(a) noise - sqrt(1-a) x0     (either output may be NULL)
 * x0 / noise / outputs [B, n] fp32, timesteps [B] int64 (device), alphas_cumprod [num_train_timesteps] fp32 (device). */
int mfb_add_noise(const float* x0, const float* noise, const long long* timesteps, const float* alphas_cumprod,
                  int num_train_timesteps, int B, long long n, float* noisy, float* velocity, void* stream);
/* F.mse_loss(pred.float(), target.float(), "mean") and the min-SNR weighted form (train_brushnet_mirror.py:1433-1450):
 *   per_sample[b] = mean_n (pred - target)^2 ;  loss = mean_b( weights[b] * per_sample[b] )   (weights NULL = 1)
 * and, in the same pass, grad = d loss / d pred = 2 (pred - target) weights[b] / (B n)  (grad / per_sample may be NULL).
 * ws: MFB_MSE_WS_FLOATS(B) floats of scratch. */
#define MFB_MSE_MAX_CHUNKS 64
#define MFB_MSE_WS_FLOATS(B) ((B) * MFB_MSE_MAX_CHUNKS)
int mfb_mse_loss(const float* pred, const float* target, const float* weights, int B, long long n, float* per_sample,
                 float* loss, float* grad, float* ws, void* stream);
/* Sum of squares of a flat fp32 gradient buffer -> *out_sq (device scalar; accumulate=1 adds to it): the total_norm^2 of
 * torch.nn.utils.clip_grad_norm_ as accelerator.clip_grad_norm_ calls it (train_brushnet_mirror.py:1460-1463).
 * ws: MFB_SQNORM_WS_FLOATS floats of scratch. */
#define MFB_SQNORM_WS_FLOATS 1184
int mfb_grad_sqnorm(const float* g, long long n, float* ws, float* out_sq, int accumulate, void* stream);
/* Gradient clipping + torch.optim.AdamW (train_brushnet_mirror.py:1190-1200,1464) + re-quantisation of the working
 * weights, one pass over flat fp32 buffers of n elements:
 *   g' = grad * grad_scale * min(1, max_grad_norm / (sqrt(*grad_sqnorm) * |grad_scale| + 1e-6))   (no clipping if grad_sqnorm NULL or max <= 0)
 *   p *= 1 - lr*wd ; m = lerp(m, g', 1-beta1) ; v = beta2 v + (1-beta2) g'^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
 * hyper (device, MFB_ADAMW_HYPER_FLOATS = 12): lr, beta1, beta2, eps, weight_decay, bc1 = 1-beta1^step, sqrt(bc2) = sqrt(1-beta2^step),
 * grad_scale, 1-beta1, 1-beta2, 1-lr*wd, lr/bc1 — the last four evaluated by the caller in float64 like torch does (1 - 0.999f in fp32 is off by 1.3e-5).
 * param_bf16 (optional) receives the bf16 rounding of the updated parameters (what the tcgen05 kernels read). */
#define MFB_ADAMW_HYPER_FLOATS 12
int mfb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16, long long n,
                   const float* hyper, const float* grad_sqnorm, float max_grad_norm, void* stream);
/* Weight (and bias) gradient of a conv3x3 (stride 1 or 2) / conv1x1 / linear layer (autograd of F.conv2d / F.linear behind
 * LoRACompatibleConv/Linear, S/models/lora.py:363-377,445-451):
 *   dw[co][(kh, kw, ci)] (+)= sum_{b,h,w} dy[b,h,w,co] * x[b,h+kh-k/2,w+kw-k/2,ci] ;  dbias[co] (+)= sum dy[.,co]
 * x [B,H,W,Cin], dy [B,H/stride,W/stride,Cout] NHWC, dtype 0 = bf16 / 1 = fp32; stride 1, or 2 with ksize 3 and even H, W (Downsample2D,
 * S/models/downsampling.py:146-152: taps read x[b, 2*oh+kh-1, 2*ow+kw-1]); dw fp32 in the packed [Cout, k*k*Cin] K order of
 * mfb_conv_desc.w.  CUDA-core kernel with fp32 accumulation, one CTA per output tile (deterministic).  The DATA gradient
 * of the same layer is mfb_conv_plan_create on dy with the flipped / transposed weight (ops.pack_conv_dgrad_weight). */
int mfb_conv_wgrad(const void* x, const void* dy, int dtype, int B, int H, int W, int Cin, int Cout, int ksize, int stride, float* dw,
                   float* dbias, int accumulate, void* stream);
/* The same weight gradient for bf16 operands on the tensor cores: split-K warp-MMA GEMM (mma.sync.m16n8k16, fp32 accumulate) that
 * consumes x and dy as stored (the pixel index is the reduction dimension: ldmatrix.trans, no transposed copies), fp32 partial
 * tiles per K slice in ws, summed in slice order (deterministic).  Needs Cin % 8 == 0 and Cout % 8 == 0 (every layer of the two
 * nets except conv_in / conv_in_condition, which take the CUDA-core mfb_conv_wgrad).  ws: mfb_conv_wgrad_tc_ws_floats(...) floats. */
long long mfb_conv_wgrad_tc_ws_floats(int B, int H, int W, int Cin, int Cout, int ksize, int stride);
int mfb_conv_wgrad_tc(const void* x, const void* dy, int B, int H, int W, int Cin, int Cout, int ksize, int stride, float* dw,
                      float* dbias, int accumulate, float* ws, long long ws_floats, void* stream);
/* Backward of GroupNorm (+SiLU) (autograd of F.group_norm + F.silu, S/models/resnet.py:337-338,381,393), NHWC, dtype 0 = bf16 /
 * 1 = fp32 tensors, fp32 / fp64 math, statistics recomputed from x.  x2 / dx2: the second tensor of a channel concat (or NULL, C2 = 0)
 * exactly as in mfb_groupnorm; dy is [B, HW, C1+C2].  dgamma / dbeta [C1+C2] fp32 ((+)= with accumulate; either may be NULL).
 * dres (optional, [B, HW, C1+C2], same dtype): a gradient arriving over the residual / shortcut path, added to dx in the same pass
 * (ResnetBlock2D: d x = GroupNorm-backward(...) + d out, S/models/resnet.py:403).
 * ws: MFB_GN_BWD_WS_FLOATS(B, C) floats of scratch.  Deterministic (fixed-order reductions). */
#define MFB_GN_BWD_WS_FLOATS(B, C) (2 * (B) * (C))
int mfb_groupnorm_bwd(const void* x1, int C1, const void* x2, int C2, const void* dy, int dtype, int B, int HW, int groups, float eps,
                      const float* gamma, const float* beta, int silu, const void* dres, void* dx1, void* dx2, float* dgamma,
                      float* dbeta, float* ws, int accumulate, void* stream);
/* out[b][c] = sum over image b's HW pixels of dy[b][p][c] (fp32 [B, C]): the gradient of the per-image row bias
 * time_emb_proj(silu(emb))[:, :, None, None] (S/models/resnet.py:369-379).  dtype 0 = bf16 / 1 = fp32 dy. */
int mfb_rowsum_per_image(const void* dy, int dtype, int B, int HW, int C, float* out, void* stream);
/* y = silu(x) and / or dx = dy * silu'(x) over n fp32 elements (dy: dtype 0 = bf16 / 1 = fp32; y or dx may be NULL): the elementwise
 * piece of the timestep MLP's backward (S/models/embeddings.py:226-237, S/models/resnet.py:369-376). */
int mfb_silu_bwd(const float* x, const void* dy, int dy_dtype, float* y, float* dx, long long n, void* stream);

/* fp32 parity-mode backward of the three ops the frozen UNet's data-gradient chain adds (BASELINE config 4): CUDA-core correctness
 * instruments written to the algorithms pinned in oracle/train_oracle.py.  Verified on B200 (tests/test_gpu_zz_train_net.py); the bf16 product-path versions are below.
 * mfb_attention_bwd_f32: backward of F.scaled_dot_product_attention (S/models/attention_processor.py:1266-1268), same tensor layout as
 *   mfb_attention_f32 (head h at columns [h*d, (h+1)*d)); two deterministic passes (per query: L, D = dO.O, dq; per key: dk, dv);
 *   stats_ws: 2 * B * heads * Tq floats.  head_dim <= 160.
 * mfb_layernorm_bwd_f32: data gradient of F.layer_norm over the last dim (S/models/attention.py:313,360,386).
 * mfb_geglu_f32: on an un-fused GEGLU projection [rows, 2C] = [h | gate] (S/models/activations.py:100-103): out = h * gelu_erf(gate)
 *   (if out != NULL) and d_proj from d_out (if d_proj != NULL). */
int mfb_attention_bwd_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const float* d_out, int ldo, float* dq,
                          int lddq, float* dk, int lddk, float* dv, int lddv, float* stats_ws, int B, int heads, int head_dim, int Tq, int Tk,
                          void* stream);
int mfb_layernorm_bwd_f32(const float* x, const float* dy, int rows, int C, float eps, const float* gamma, float* dx, void* stream);
int mfb_geglu_f32(const float* proj, long long rows, int C, float* out, const float* d_out, float* d_proj, void* stream);

/* ---- bf16 backward of the bandwidth-bound ops (BASELINE config 4; csrc/train_bf16.cu): what autograd runs under accelerator.backward
 * (E/train_brushnet_mirror.py:1459) for the ops below.  All tensors bf16 channels-last unless noted; fp32 math; deterministic.
 * mfb_groupnorm_stats: stats_ws[0 : 2*B*groups] = per-(image, group) {sum, sum of squares} of the (two-source) input — the forward
 *   statistics pass alone (workspace contract of mfb_groupnorm); the backward below reads them.
 * mfb_groupnorm_bwd2: backward of F.group_norm (+ F.silu if silu) (S/models/resnet.py:337-338,381,393; transformer_2d.py:338; eps per
 *   call site), two-source concat like mfb_groupnorm.  dx1 / dx2 per source; dres / dres2 ([B, HW, C1+C2], may be NULL): gradients
 *   arriving over residual / skip paths, added in the same pass (S/models/resnet.py:403; the UNet's skip fan-out).  dgamma / dbeta
 *   [C] fp32 (NULL for a frozen layer), accumulate != 0 adds to them.  ws: mfb_groupnorm_bwd2_ws_floats(B, C, groups) floats, zero on
 *   first use (ticket counters).  Two vectorised passes (per-channel sums -> group means; dx).
 * mfb_layernorm_bwd: data gradient of F.layer_norm over the last dim (S/models/attention.py:313,360,386) + optional residual-path
 *   gradient dres [rows, C] (the transformer's `+ hidden_states`, :330,372,409).
 * mfb_geglu: un-fused GEGLU on proj [rows, 2C] = [h | gate] (S/models/activations.py:100-103): out [rows, C] = h * gelu_erf(gate) if
 *   out != NULL; d_proj [rows, 2C] from d_out [rows, C] if d_proj != NULL (training keeps proj for the backward).
 * mfb_conv_out_bwd: data gradient of conv_out (3x3, Cin -> Cout, S/models/unets/unet_2d_condition.py:1339): dy fp32 NCHW
 *   [B, Cout, H, W] (d loss / d model_pred, from mfb_mse_loss), w fp32 [Cout, 3, 3, Cin] (the layout mfb_conv_out reads),
 *   dx bf16 [B, H*W, Cin]. */
/* mfb_attention_lse: mfb_attention that also writes lse [B, heads, Tq] fp32 = log2 sum_j 2^(q_i.k_j * d^-1/2 * log2 e), kept for the backward.
 * mfb_attention_bwd: backward of F.scaled_dot_product_attention (S/models/attention_processor.py:1266-1268) on tcgen05, flash style
 *   (csrc/attn_bwd.cu): q / k / v / o / d_o bf16 in the layouts of mfb_attention (leading dimensions in elements, head h at columns
 *   [h*d, (h+1)*d)); lse from the forward; dvec [B, heads, Tq] fp32 scratch (D = dO . O, written here); dq always; dk / dv both or
 *   neither (NULL: cross attention to a frozen context, only the query-side kernel runs).  Deterministic: every output element
 *   has one writer. */
int mfb_attention_lse(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo, int B, int heads,
                      int head_dim, int Tq, int Tk, float* lse, void* stream);
int mfb_attention_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* o, int ldo, const void* d_o,
                      int lddo, const float* lse, float* dvec, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int B,
                      int heads, int head_dim, int Tq, int Tk, void* stream);
long long mfb_groupnorm_bwd2_ws_floats(int B, int C, int groups);
int mfb_groupnorm_stats(const void* x1, int C1, const void* x2, int C2, int B, int HW, int groups, float* stats_ws, void* stream);
int mfb_groupnorm_bwd2(const void* x1, int C1, const void* x2, int C2, const void* dy, int B, int HW, int groups, float eps,
                       const float* gamma, const float* beta, int silu, const float* stats, const void* dres, const void* dres2,
                       void* dx1, void* dx2, float* dgamma, float* dbeta, float* ws, int accumulate, void* stream);
int mfb_layernorm_bwd(const void* x, const void* dy, int rows, int C, float eps, const float* gamma, const void* dres, void* dx,
                      void* stream);
int mfb_geglu(const void* proj, long long rows, int C, void* out, const void* d_out, void* d_proj, void* stream);
int mfb_conv_out_bwd(const float* dy, int B, int H, int W, int Cin, int Cout, const float* w, void* dx, void* stream);
/* Packed weight of a stride-1 conv's / linear's DATA gradient from its packed forward weight (both bf16): wd [Cin, k*k*Cout] with
 * wd[ci, (k*k-1-t)*Cout + co] = w[co, t*Cin + ci] (autograd's dgrad of F.conv2d / F.linear, S/models/lora.py:363-377,445-451, as
 * "the same implicit GEMM on dy with the flipped / transposed weight"); re-derived after every optimizer step for trainable layers. */
int mfb_dgrad_repack(const void* w, int Cout, int Cin, int ksize, void* wd, void* stream);
/* Adjoint of Upsample2D's nearest-x2 replication (S/models/upsampling.py:167-173): dx [B, H*W, C] = 2x2 sums of du [B, 2H*2W, C] (bf16). */
int mfb_sumpool2x2(const void* du, int B, int H, int W, int C, void* dx, void* stream);
/* fp32 PARITY-MODE forms of the two kernels above (all tensors fp32, same semantics), and y += x over n fp32 elements: in parity
 * mode the residual / skip-path gradients that the bf16 kernels fold into their last pass (`dres` of mfb_layernorm_bwd, `dres2` of
 * mfb_groupnorm_bwd2) are added by this launch.  Together with mfb_attention_bwd_f32 / mfb_layernorm_bwd_f32 / mfb_geglu_f32 /
 * mfb_groupnorm_bwd(dtype 1) and the fp32 conv plans they run the frozen UNet's data-gradient chain of the fine-tune step
 * (E/train_brushnet_mirror.py:836-888,1459) in fp32 on the device, for the 1e-3 bar against float64 autograd. */
int mfb_conv_out_bwd_f32(const float* dy, int B, int H, int W, int Cin, int Cout, const float* w, float* dx, void* stream);
int mfb_sumpool2x2_f32(const float* du, int B, int H, int W, int C, float* dx, void* stream);
int mfb_add_f32(float* y, const float* x, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFB200_H */
