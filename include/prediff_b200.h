/* prediff_b200 - C ABI of the B200-native PreDiff sampling path.
 *
 * Drop-in boundary for the one hot path of gaozhihan/PreDiff:
 *   LatentDiffusion.p_sample_loop -> CuboidTransformerUNet.forward -> AutoencoderKL.encode / decode.
 * Every entry point takes plain pointers and sizes (no torch types), returns 0 on success and a negative
 * PD_ERR_* code on failure (never throws; message via pd_last_error()), and enqueues its work on the CUDA
 * stream passed as `stream` (a cudaStream_t; NULL = legacy default stream). Device pointers must be 16-byte
 * aligned. There is no CPU fallback: on a non-sm_100 device every call fails with PD_ERR_ARCH.
 *
 * Tensor layouts are channels-last, fp32 unless stated:
 *   latent   z, cond : [B][T][H][W][C]      (the reference's "NTHWC", latent_diffusion.py:394-402)
 *   VAE pixel frames : [N][H][W]            (== the reference's (N,1,H,W))
 *   VAE latents      : [N][H/8][W/8][C]     (the reference's (N,C,h,w) permuted; the Python shim permutes views)
 */
#ifndef PREDIFF_B200_H
#define PREDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_OK 0
#define PD_ERR_CUDA (-1)
#define PD_ERR_SHAPE (-2)
#define PD_ERR_ARCH (-3)
#define PD_ERR_WEIGHT (-4)
#define PD_ERR_ARG (-5)
#define PD_ERR_STATE (-6)

/* ---- library ------------------------------------------------------------------------------------------- */
/* Checks the current device is sm_100, resolves driver entry points, raises kernel smem limits. Idempotent. */
int pd_init(void);
/* Message of the last failure on the calling thread ("" if none). */
const char* pd_last_error(void);
const char* pd_version(void);

/* ---- CuboidTransformerUNet (reference: src/prediff/models/cuboid_transformer/cuboid_transformer_unet.py) -- */
typedef struct pd_unet pd_unet;
typedef struct pd_unet_config {
    int32_t t_in, t_out;   /* input_shape[0], target_shape[0]                       (cfg.yaml:158-159: 7, 6)   */
    int32_t h, w, c;       /* latent H, W, C                                         (16, 16, 64)               */
    int32_t base_units;    /* cfg.yaml:160 (256); level-1 width is 2*base_units                                 */
    int32_t depth[2];      /* cfg.yaml:171 ([4, 4]); exactly two levels                                          */
    int32_t num_heads;     /* cfg.yaml:163 (4)                                                                   */
    int32_t max_batch;     /* largest batch a forward will be asked for (sizes the activation arena)             */
} pd_unet_config;

int pd_unet_create(const pd_unet_config* cfg, pd_unet** out);
/* block_attn_patterns other than 'axial' (cuboid_transformer_unet.py:201-214 -> StackCuboidSelfAttentionBlock
 * cuboid_transformer.py:1026-1110): per level, the (cuboid_size, strategy, shift_size) of every attention layer of the
 * level's stack block, as the pattern functions of cuboid_transformer_patterns.py return them for the level's
 * (T, H, W). strategy: 0 = 'l', 1 = 'd'. padding_type: 0 = 'zeros', 1 = 'ignore'. pd_unet_create == pd_unet_create_ex
 * with the axial pattern at both levels and 'zeros' padding (the shipped SEVIR-LR config). */
#define PD_MAX_ATTN_LAYERS 8
typedef struct pd_unet_pattern {
    int32_t n_layers[2];
    int32_t cuboid_size[2][PD_MAX_ATTN_LAYERS][3];
    int32_t strategy[2][PD_MAX_ATTN_LAYERS][3];
    int32_t shift_size[2][PD_MAX_ATTN_LAYERS][3];
    int32_t padding_type;
} pd_unet_pattern;
int pd_unet_create_ex(const pd_unet_config* cfg, const pd_unet_pattern* pattern, pd_unet** out);
/* ... with global vectors (cuboid_transformer_unet.py:55-60, 124-126: num_global_vectors > 0, separate_global_qkv=False,
 * global_dim_ratio=1): num_global_vectors in 1..32; use_global_vector_ffn / use_global_self_attn as the reference's
 * constructor arguments. pattern may be NULL (axial, 'zeros'). Adds the state_dict keys init_global_vectors,
 * {down,up}_layer_global_proj.0.*, <block>.attn_l.i.{global_qkv,global_proj,global_vec_norm}.* and <block>.global_ffn_l.i.*.
 * bf16 operand precision only. pd_unet_create_gv == pd_unet_create_gv_ex with separate_global_qkv = 0. */
int pd_unet_create_gv(const pd_unet_config* cfg, const pd_unet_pattern* pattern, int num_global_vectors,
                      int use_global_vector_ffn, int use_global_self_attn, pd_unet** out);
/* ... with separate_global_qkv (cuboid_transformer.py:770-795, the shipped cfg.yaml's value for the UNet): the keys
 * <block>.attn_l.i.{l2g_q_net, l2g_global_kv_net, g2l_global_q_net, g2l_k_net, g2l_v_net[, g2g_global_qkv_net]}.weight replace
 * global_qkv.weight. global_dim_ratio is 1. */
int pd_unet_create_gv_ex(const pd_unet_config* cfg, const pd_unet_pattern* pattern, int num_global_vectors,
                         int use_global_vector_ffn, int use_global_self_attn, int separate_global_qkv, pd_unet** out);
void pd_unet_destroy(pd_unet* m);
/* Number of state_dict entries the model expects; name/shape of entry i (reference key names, e.g.
 * "down_self_blocks.0.1.attn_l.2.qkv.weight"). shape has up to 5 dims; returns ndim. */
int pd_unet_num_weights(const pd_unet* m);
int pd_unet_weight_info(const pd_unet* m, int i, const char** name, int64_t shape[5]);
/* Copies one fp32 tensor (host or device pointer, contiguous, reference layout) into the model. */
int pd_unet_load_weight(pd_unet* m, const char* name, const float* data, const int64_t* shape, int ndim);
/* Operand precision of the model's tensor-core GEMMs / convolutions (accumulation, residual stream, normalisation
 * statistics, softmax and the sampler update are fp32 in both):
 *   PD_PRECISION_BF16 (default): bf16 operands, tcgen05.mma kind::f16 - rel-RMS ~6e-3 per UNet step vs the fp32 reference;
 *   PD_PRECISION_TF32: fp32 storage rounded to tf32, tcgen05.mma kind::tf32 (half the bf16 rate) and an fp32 attention
 *     core - the reference's own GPU arithmetic (torch.set_float32_matmul_precision("high"): scripts/prediff/sevirlr/
 *     cfg.yaml:32, train_sevirlr_prediff.py:1143), rel-RMS <= 2e-3 per step and after the 50-step loop.
 * Call before pd_unet_finalize (a change un-finalizes the model). TF32 is built for the axial pattern (shipped config). */
#define PD_PRECISION_BF16 0
#define PD_PRECISION_TF32 1
int pd_unet_set_precision(pd_unet* m, int precision);
/* CTAs per sample of the stream-K schedule of the long-K convolutions (0 = the default, 36: the same cut for every batch, so
 * a sample's result is bit-identical whatever batch it runs in - the shard-invariance property of the multi-GPU path). A
 * model that serves single samples may ask for more (72: 3.45 -> 3.19 ms per batch-1 denoise step); results then differ
 * from the default cut's in the last bits (another fixed summation order), never between runs. Call before pd_unet_finalize. */
int pd_unet_set_streamk_ctas(pd_unet* m, int ctas_per_sample);
/* Repacks all weights into kernel layouts (bf16 or tf32-rounded fp32, K-major, tap-major convs). Fails if any weight is
 * missing. */
int pd_unet_finalize(pd_unet* m);
/* forward(x, t, cond) -> eps   (cuboid_transformer_unet.py:406-493)
 *   x [B][t_out][h][w][c], t [B] int64, cond [B][t_in][h][w][c]  ->  out [B][t_out][h][w][c]; all device ptrs. */
int pd_unet_forward(pd_unet* m, const float* x, const int64_t* t, const float* cond, float* out, int batch,
                    void* stream);

/* Measurement helpers. profile: one eager forward with every launch bracketed by CUDA events on `stream`;
 * stats = {tensor-core GEMM/conv ms, #GEMM launches, other-kernel ms, #other launches, GEMM FLOPs}.
 * kernels_per_forward: number of kernel launches one forward of this batch size issues. */
int pd_unet_profile_forward(pd_unet* m, const float* x, const int64_t* t, const float* cond, float* out, int batch,
                            void* stream, double stats[5]);
int pd_unet_kernels_per_forward(pd_unet* m, int batch, int* n);
/* Per-launch trace: one eager forward with a %globaltimer stamp kernel after every plan step. ns_dev[i] (device u64,
 * max_slots entries) = stamp after step i-1 (ns_dev[0] = start); labels (host, optional) receives one '\n'-terminated
 * label per step ("L0.stack.ffn1", "L1.res.conv2", ...). Returns the number of steps (>= 0) or a PD_ERR_* code. */
/* Algorithmic FLOPs (2 x MACs) of every plan step, in pd_unet_trace_forward's step order (0 for non-GEMM steps); returns
 * the number of steps or a PD_ERR_* code. Lets bench.py turn the per-launch trace into per-kernel roofline fractions. */
int pd_unet_step_flops(pd_unet* m, int batch, double* flops, int max_slots);
int pd_unet_trace_forward(pd_unet* m, const float* x, const int64_t* t, const float* cond, float* out, int batch,
                          void* stream, unsigned long long* ns_dev, int max_slots, char* labels, int labels_bytes);

/* ---- AutoencoderKL (reference: src/prediff/taming/autoencoder_kl.py:80-113, vae.py:70-86,150-166) --------- */
typedef struct pd_vae pd_vae;
typedef struct pd_vae_config {
    int32_t in_channels;          /* 1 */
    int32_t out_channels;         /* 1 */
    int32_t latent_channels;      /* 64 */
    int32_t block_out_channels[4];/* 128, 256, 512, 512 */
    int32_t layers_per_block;     /* 2 */
    int32_t norm_num_groups;      /* 32 */
    int32_t h, w;                 /* pixel frame size (128, 128) */
    int32_t max_frames;           /* largest N per encode/decode call */
} pd_vae_config;

int pd_vae_create(const pd_vae_config* cfg, pd_vae** out);
void pd_vae_destroy(pd_vae* m);
int pd_vae_num_weights(const pd_vae* m);
int pd_vae_weight_info(const pd_vae* m, int i, const char** name, int64_t shape[5]);
int pd_vae_load_weight(pd_vae* m, const char* name, const float* data, const int64_t* shape, int ndim);
int pd_vae_finalize(pd_vae* m);
/* encode: x [N][H][W] -> moments [N][H/8][W/8][2*latent] (mean | logvar), i.e. quant_conv(Encoder(x)). */
int pd_vae_encode(pd_vae* m, const float* x, float* moments, int n, void* stream);
/* decode: z [N][H/8][W/8][latent] -> [N][H][W], i.e. Decoder(post_quant_conv(z)). */
int pd_vae_decode(pd_vae* m, const float* z, float* out, int n, void* stream);

/* ---- knowledge-alignment network (reference: src/prediff/diffusion/knowledge_alignment/models.py:459-528
 *      NoisyCuboidTransformerEncoder; sevir.py:55-104 SEVIRAvgIntensityAlignment; alignment_pl.py:423-446) -------- */
typedef struct pd_ka pd_ka;
typedef struct pd_ka_config {
    int32_t t;             /* input_shape[0] = number of target frames      (cfg.yaml:105-156: 6)                 */
    int32_t h, w, c;       /* latent H, W, C                                 (16, 16, 64)                          */
    int32_t base_units;    /* 128; level-1 width is 2*base_units                                                   */
    int32_t depth[2];      /* [1, 1]                                                                               */
    int32_t num_heads;     /* 4                                                                                    */
    int32_t max_batch;
} pd_ka_config;

int pd_ka_create(const pd_ka_config* cfg, pd_ka** out);
void pd_ka_destroy(pd_ka* m);
int pd_ka_num_weights(const pd_ka* m);
int pd_ka_weight_info(const pd_ka* m, int i, const char** name, int64_t shape[5]);
int pd_ka_load_weight(pd_ka* m, const char* name, const float* data, const int64_t* shape, int ndim);
int pd_ka_finalize(pd_ka* m);
/* forward(zt, t) -> pred [B][t]  (models.py:459-528 with pool="attention", readout_seq=True; the reference returns
 * (B, t, 1)). zt [B][t][h][w][c] fp32, t [B] int64; device pointers. */
int pd_ka_forward(pd_ka* m, const float* zt, const int64_t* t, float* pred, int batch, void* stream);
/* get_mean_shift (sevir.py:85-104): grad [B][t][h][w][c] = guide_scale * d || mean_T U(zt, t) - avg_x_gt ||_2 / d zt,
 * the L2 norm taken over the whole batch (sevir.py:82). avg_x_gt: device fp32 [B]. Forward + hand-written
 * input-gradient backward on the device (replaces torch.autograd.grad, alignment_pl.py:441-445).
 * loss_out: optional device float receiving the alignment value. */
int pd_ka_mean_shift(pd_ka* m, const float* zt, const int64_t* t, const float* avg_x_gt, float guide_scale, float* grad,
                     float* loss_out, int batch, void* stream);
/* kernel launches of one forward / one backward at this batch size */
int pd_ka_kernels(pd_ka* m, int batch, int* n_forward, int* n_backward);

/* ---- sampler (reference: latent_diffusion.py:228-278,553-684; diffusion/utils.py:17-70) -------------------- */
typedef struct pd_sampler pd_sampler;
#define PD_MODE_DDPM 0  /* the reference's ancestral p_sample_loop, t = timesteps-1 .. 0                        */
#define PD_MODE_DDIM 1  /* DDIM on make_ddim_timesteps('uniform', n, T) with eta        */

/* Builds the beta schedule ("linear": linspace(sqrt(start), sqrt(end), T)^2 in float64) and keeps the per-step
 * coefficient table resident on the device. */
int pd_sampler_create(int num_timesteps, double linear_start, double linear_end, pd_sampler** out);
void pd_sampler_destroy(pd_sampler* s);
/* Copies one of the reference's registered schedule buffers (by its buffer name, e.g. "alphas_cumprod",
 * "posterior_mean_coef1") to host memory `out[num_timesteps]` - for parity tests against register_schedule. */
int pd_sampler_get_buffer(const pd_sampler* s, const char* name, float* out);
/* Runs the whole denoising loop on the device:
 *   z     [B][t_out][h][w][c]  in: z_T, out: z_0 (updated in place)
 *   cond  [B][t_in][h][w][c]
 *   noise NULL (DDIM eta=0) or [n_steps][B*t_out*h*w*c] pre-generated N(0,1) (step k uses slice k)
 *   mode PD_MODE_*; n_steps: DDPM = number of ancestral steps (t = n_steps-1..0), DDIM = number of DDIM steps.
 * The loop state lives in buffers owned by the sampler (z / cond are copied in, z is copied back), so the CUDA graph
 * captured on first use for a given (unet weights generation, batch) is replayed whatever the caller's addresses are;
 * loops of up to 64 steps are ONE graph launch, longer ones replay a one-iteration graph. */
int pd_sample_loop(pd_sampler* s, pd_unet* unet, float* z, const float* cond, const float* noise, int batch, int mode,
                   int n_steps, float eta, void* stream);
/* Same, restricted to the executed steps [k_begin, k_end) of the n_steps-step schedule (k = 0 is the noisiest step);
 * noise[0] belongs to step k_begin. Lets the host interleave callbacks / intermediates / inpainting
 * (latent_diffusion.py:659-680) between device-resident stretches. */
int pd_sample_loop_range(pd_sampler* s, pd_unet* unet, float* z, const float* cond, const float* noise, int batch,
                         int mode, int n_steps, float eta, int k_begin, int k_end, void* stream);
/* pd_sample_loop_range with knowledge-alignment guidance (latent_diffusion.py:592-596 aligned_mean for DDPM;
 * SURVEY.md section 8 S6 for DDIM: eps_hat = eps + sqrt(1 - a_t) * g): every step also evaluates
 * g = pd_ka_mean_shift(z_t, t, avg_x_gt, guide_scale) - concurrently with the UNet on a second stream - and the fused
 * update subtracts coef * g. The whole iteration (UNet + KA forward/backward + update) is one CUDA-graph replay. */
int pd_sample_loop_aligned(pd_sampler* s, pd_unet* unet, pd_ka* ka, float* z, const float* cond, const float* noise,
                           const float* avg_x_gt, float guide_scale, int batch, int mode, int n_steps, float eta,
                           int k_begin, int k_end, void* stream);
/* Number of independent sub-batches (parallel streams) the loop cuts a batch into (env PD_SUB_BATCHES, default 2). */
/* Forward-only diffusion loss = LatentDiffusion.p_losses (latent_diffusion.py:517-551) for the eps-parameterization
 * with a fixed logvar (learn_logvar = False): x_noisy = q_sample(x_start, t, noise) (:489-492), eps = UNet(x_noisy, t,
 * cond), loss_simple_b = mean |noise - eps|^p per sample (p = 2 'l2', loss_l1 = 1 -> p = 1), then
 * out4 = {mean loss_simple, loss_vlb = mean(lvlb_weights[t] loss_simple), loss = l_simple_weight * mean(loss_simple /
 * exp(logvar) + logvar) + original_elbo_weight * loss_vlb, loss_gamma}. This is what validation_step evaluates through
 * self(batch) (train_sevirlr_prediff.py:817-818); no gradients. All pointers on the device: x_start / noise
 * [B][t_out][h][w][c], cond [B][t_in][h][w][c], t int64 [B], per_sample [B], out4 [4]. "lvlb_weights" is also
 * available from pd_sampler_get_buffer. */
int pd_diffusion_losses(pd_sampler* s, pd_unet* unet, const float* x_start, const float* cond, const int64_t* t,
                        const float* noise, int batch, int loss_l1, float logvar, float l_simple_weight,
                        float original_elbo_weight, float* per_sample, float* out4, void* stream);
int pd_sampler_sub_batches(const pd_sampler* s, int batch);
/* clip_denoised of p_mean_variance (latent_diffusion.py:580-581): when on, every DDPM step clamps its z_0 estimate to
 * [-1, 1] before the posterior mean (inside the fused update kernel). Off by default, as in the shipped config. */
int pd_sampler_set_clip_denoised(pd_sampler* s, int on);
/* One reference p_sample step at integer timestep t (all batch rows share t): z <- p_sample(z, cond, t). */
int pd_sample_step_ddpm(pd_sampler* s, pd_unet* unet, float* z, const float* cond, const float* noise, int batch, int t,
                        void* stream);

/* ---- on-device evaluation step (reference: src/prediff/datasets/sevir/evaluation.py:88-285 SEVIRSkillScore.update,
 *      _threshold :12-37; sevir_dataloader.py:652-681 back-transform; torchmetrics MSE / MAE sums beside it in
 *      train_sevirlr_prediff.py:960-962) ------------------------------------------------------------------------ */
/* One pass over forecast and target frames, fp32 [N][T][H][W] in [0,1] (device): values are mapped back to VIL units
 * (x / (1/255)), optionally max-pooled (`pool` x `pool`, 1 = none), compared with each threshold (NaN in either ->
 * neither hit, miss nor false alarm) and ADDED to counts int64 [n_thresholds][T][3] = (hits, misses, false alarms)
 * per threshold and lead time; sums double [2] += (sum (pred - target)^2, sum |pred - target|) over the raw pixels.
 * thresholds_host: host pointer (e.g. 16, 74, 133, 160, 181, 219), at most 8. Integer results are exact. */
int pd_sevir_eval_update(const float* pred, const float* target, int64_t* counts, double* sums, int N, int T, int H,
                         int W, int pool, const float* thresholds_host, int n_thresholds, void* stream);

/* SSIM beside it (train_sevirlr_prediff.py:229-230 `torchmetrics.image.StructuralSimilarityIndexMeasure()`, updated with
 * "(b t) c h w" frames at :964-965; torchmetrics==1.2.0 per setup.py:33, functional/image/ssim.py): 11 x 11 Gaussian window,
 * sigma 1.5, k1 0.01, k2 0.03, borders cropped by 5, per-image mean, reduction elementwise_mean. pred / target fp32
 * [N][H][W] (one channel) on the device; data_range <= 0 means None (max of the two value ranges of this batch).
 * state (device double [2]) += {sum of the per-image SSIM values, N}; SSIM = state[0] / state[1]. */
int pd_ssim_update(const float* pred, const float* target, int N, int H, int W, float data_range, double* state,
                   void* stream);

/* ---- input side (reference: src/prediff/datasets/sevir/sevir_dataloader.py:834-877 SEVIRDataLoader._idx_sample in
 *      'sequent' mode, :610-650 preprocess_data_dict, :71-84 change_layout, constants :25-44) ------------------------- */
/* events_u8: raw VIL events as stored in the SEVIR-LR HDF5 files, uint8 [n_events][H][W][T_raw] ('NHWT') on the device,
 * holding events event_base .. event_base + n_events - 1 of the split. Writes the sequent windows first_seq ..
 * first_seq + batch - 1 (window s = event s / n, frames (s % n) * stride .. + seq_len, n = 1 + (T_raw - seq_len) / stride)
 * as fp32 [batch][seq_len][H][W][1] = scale * (x + offset) - bit-identical to the reference batch in layout 'NTHWC'
 * (scale, offset: 1/255, 0 for rescale '01'; 1/47.54, -33.44 for 'sevir'). PD_ERR_SHAPE if a window needs an event
 * outside the buffer. */
int pd_sevir_windows(const unsigned char* events_u8, int event_base, int n_events, int H, int W, int T_raw,
                     long long first_seq, int batch, int seq_len, int stride, float scale, float offset, float* out,
                     void* stream);

/* ---- kernel-level entry points (used by the parity tests; same kernels the models launch) ------------------ */
/* out = epilogue(conv/linear(A, Wt)): A bf16 [samples][D][H][W][C], Wt bf16 [N][kt*kh*kw*C], zero padding k/2.
 * bias[N], rowvec[samples][N], residual/out_f32 fp32 [M][N], out_bf16 bf16 [M][N]; any of them may be NULL.
 * act: 0 none, 1 GELU(erf), 2 SiLU. block_n: 0 = heuristic, else 32/64/128/256. */
int pd_op_conv_gemm(const void* A_bf16, const void* Wt_bf16, int samples, int D, int H, int W, int C, int kt, int kh,
                    int kw, int N, const float* bias, const float* rowvec, const float* residual, float* out_f32,
                    void* out_bf16, int act, int block_n, void* stream);
/* The same kernels with tf32 operands (PD_PRECISION_TF32): A fp32 [samples][D][H][W][C] (C % 32 == 0), Wt fp32
 * [N][kt*kh*kw*C], fp32 output only; round_out = 1 stores the output rounded to tf32 (it feeds another tf32 GEMM);
 * streamk_ctas_per_sample > 0 runs the stream-K schedule (N % 256 == 0, no activation). The tensor core reads the top 19
 * bits of every operand word; pd_op_pack_tf32 / the producer kernels round to nearest beforehand. */
int pd_op_conv_gemm_tf32(const float* A, const float* Wt, int samples, int D, int H, int W, int C, int kt, int kh, int kw,
                         int N, const float* bias, const float* rowvec, const float* residual, float* out_f32, int act,
                         int round_out, int block_n, int streamk_ctas_per_sample, void* stream);
/* Weight repack to tf32-rounded fp32: taps == 0: Linear fp32 [Co][Ci] -> [Co][Cipad]; taps > 0: conv fp32
 * [Co][Ci][taps] -> [Co][taps][Cipad]. */
int pd_op_pack_tf32(const float* w, float* out, int Co, int Ci, int taps, int Cipad, void* stream);
/* Axial attention core in fp32 (qkv fp32 [B][T][H][W][3C] -> out fp32 [B][T][H][W][C], tf32-rounded). */
int pd_op_axial_attention_f32(const float* qkv, const float* bias_table, float* out, int B, int T, int H, int W, int C,
                              int heads, int axis, void* stream);
/* kind 0: GroupNorm(+SiLU) x fp32 [S][R][C] -> y fp32 (tf32-rounded); kind 1: LayerNorm over C of the S*R rows. */
int pd_op_norm_tf32(int kind, const float* x, const float* gamma, const float* beta, float* y, int S, int R, int C, int G,
                    float eps, int silu, void* stream);
/* x[M][256] += A[M][K] Wt[256][K]^T + bias, and ln_out[M][256] (bf16) = LayerNorm(x row, eps 1e-5) * gamma + beta
 * from the same epilogue (the fusion the UNet uses for proj / ffn_2 / conv2 -> next pre-norm at width 256). */
int pd_op_linear_residual_ln(const void* A_bf16, const void* Wt_bf16, int M, int K, const float* bias, float* x_inout,
                             const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, void* stream);
/* Same for a row width N of 256 or 512. N = 512: the two CTAs that share a row tile form a thread-block cluster and
 * exchange their partial row sums through distributed shared memory (the level-1 width of the UNet). */
int pd_op_linear_residual_ln_n(const void* A_bf16, const void* Wt_bf16, int M, int K, int N, const float* bias,
                               float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                               void* stream);
/* Same launch with clock64() phase stamps of CTA (dbg_block, 0) written to stamps9[0..8] (device u64[16]; [9], [10] = %globaltimer ns at entry / exit):
 * entry, setup done, first operand tile landed, last MMA issued, accumulator ready, first epilogue chunk ready,
 * epilogue done, last bulk store drained, exit. Profiling aid (tools/gemm_phases.py). */
int pd_op_conv_gemm_phases(const void* A_bf16, const void* Wt_bf16, int samples, int D, int H, int W, int C, int kt,
                           int kh, int kw, int N, const float* bias, const float* residual, float* out_f32,
                           void* out_bf16, int act, int block_n, int dbg_block, unsigned long long* stamps9,
                           void* stream);
int pd_op_group_norm(const float* x, const float* gamma, const float* beta, void* y_bf16, int S, int R, int C, int G,
                     float eps, int silu, void* stream);
int pd_op_layer_norm(const float* x, const float* gamma, const float* beta, void* y_bf16, int P, int C, float eps,
                     void* stream);
int pd_op_patch_merge_ln(const float* x, const float* gamma, const float* beta, void* y_bf16, int BT, int H, int W,
                         int C, float eps, void* stream);
int pd_op_axial_attention(const void* qkv_bf16, const float* bias_table, void* out_bf16, int B, int T, int H, int W,
                          int C, int heads, int axis, void* stream);
/* General cuboid self-attention: any cuboid_size / strategy (0 = 'l', 1 = 'd') / shift_size of
 * CuboidSelfAttentionLayer (cuboid_transformer.py:595-966, no global vectors), padding_type 0 = 'zeros', 1 = 'ignore',
 * 2 = 'nearest' (models/utils.py:228-270: the padded grid is a nearest-neighbour resampling of the tokens).
 * pd_cuboid_tables is host-only (no GPU): fills meta = {effective size[3], effective shift[3], pad[3], num_cuboids,
 * volume, rel_off} and, when non-null, tok / lab [num_cuboids*volume] and rel [volume] (see csrc/ops.cuh); returns
 * 1 + axis when the layer is the axial fast path, 0 when it takes the general kernel, < 0 on error.
 * pd_op_cuboid_attention: qkv bf16 [B][T][H][W][3C] -> out bf16 [B][T][H][W][C], bias_table fp32 [n_rel][heads]. */
int pd_cuboid_tables(int T, int H, int W, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                     int padding_type, int32_t meta[12], int32_t* tok, int32_t* lab, int32_t* rel, int64_t capacity);
/* padding_type 2 ('nearest') only: dst [num_cuboids*volume] = the token each slot's result is written to (-1 = nobody) -
 * `tok` of pd_cuboid_tables is then the token the slot's q|k|v rows copy. Returns 1 if the layer has such a table (padding
 * on some axis), 0 if results go to `tok` as for the other padding types (dst untouched), < 0 on error. */
int pd_cuboid_tables_dst(int T, int H, int W, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                         int padding_type, int32_t* dst, int64_t capacity);
int pd_op_cuboid_attention(const void* qkv_bf16, const float* bias_table, void* out_bf16, int B, int T, int H, int W, int C,
                           int heads, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                           int padding_type, void* stream);
/* Same with the kernel chosen explicitly: impl 0 = as the models choose, 1 = warp-level mma.sync kernel (any head dim /
 * volume), 2 = tcgen05 tile kernel (csrc/attention_tc.cu: 128-query tile per (cuboid, head, sample), S and O accumulators in
 * TMEM, K/V chunks of 128 keys through a swizzled shared-memory ring, softmax warps reading their row with tcgen05.ld;
 * head dim 64 or 128). The models use it for cuboid volumes >= 128 (video_swin_PxM, divided_st, full). */
int pd_op_cuboid_attention_impl(const void* qkv_bf16, const float* bias_table, void* out_bf16, int B, int T, int H, int W,
                                int C, int heads, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                                int padding_type, int impl, void* stream);
/* Global vectors (cuboid_transformer.py:864-945, use_global_vector with the shared global_qkv net): K <= 32 vectors per
 * sample that every cuboid's queries see as extra, never-masked keys and whose own queries attend over all num_cuboids *
 * volume slots (+ themselves with use_global_self_attn).
 * pd_cuboid_tables_gmask (host-only): padding_type 1 ('ignore') only - gmask [num_cuboids*volume], 1 = the slot is visible to
 *   the global queries: the validity of the padded, rolled frame flattened in raster order, applied to the cuboid-ordered
 *   slots as the reference does (:915-924). Returns 1 if the layer has such a mask, 0 if every slot is visible, < 0 on error.
 * pd_op_gv_linear: out[r][n] = (res ? res[r][n] : 0) + act(sum_k f(in[r][k]) W[n][k] + bias[n]) on the M = B * K global rows,
 *   fp32 on the reference-layout fp32 weight [N][K]; f = LayerNorm(ln_gamma, ln_beta, 1e-5) if ln_gamma (global_vec_norm /
 *   the global FFN's pre-norm; K <= 512), act 1 = GELU(erf); out_f32 / out_bf16: either or both; res may alias out_f32.
 * pd_op_cuboid_attention_gv: qkv bf16 [B][T][H][W][3C] and the global rows' q|k|v (gqkv_f32 [B][K][3C] and its bf16 copy)
 *   -> out bf16 [B][T][H][W][C] (local + local-to-global attention, :902-913) and gout fp32 [B][K][C] (global-to-local
 *   (+ global-to-global) attention, :928-945), both before their output projections. */
/* pd_op_cuboid_attention_gv2: the same pair of kernels with the operand layouts of both global-vector variants.
 *   tok2_bf16 == NULL: the shared global_qkv net - grow_* = the global rows' [q | k | v], grow_ld = 3C.
 *   tok2_bf16 [B][T][H][W][3C] = the tokens' [l2g_q | g2l_k | g2l_v] rows: separate_global_qkv=True (cuboid_transformer.py:
 *   866-891) - grow_* = [l2g_k | l2g_v | g2l_q] (grow_ld = 3C) or, with self_attn, [... | g2g_q | g2g_k | g2g_v] (6C).
 *   line_kernel = 1: the token grid's part through the axial line kernel (axial layers, n_global <= 16), as the UNet does. */
int pd_op_cuboid_attention_gv2(const void* qkv_bf16, const float* bias_table, const void* tok2_bf16, const float* grow_f32,
                               const void* grow_bf16, int grow_ld, void* out_bf16, float* gout, int B, int T, int H, int W, int C,
                               int heads, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                               int padding_type, int n_global, int self_attn, int line_kernel, void* stream);
/* pd_op_axial_attention_gv: pd_op_axial_attention with the sample's <= 16 global vectors as extra keys of every line (k | v
 *   rows of gqkv_bf16 [B][K][3C]; unmasked, no position bias) - what the UNet runs for axial layers when K <= 16. */
int pd_op_axial_attention_gv(const void* qkv_bf16, const float* bias_table, const void* gqkv_bf16, void* out_bf16, int B, int T,
                             int H, int W, int C, int heads, int axis, int n_global, void* stream);
int pd_cuboid_tables_gmask(int T, int H, int W, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                           int padding_type, int32_t* gmask, int64_t capacity);
int pd_op_gv_linear(const float* in, const float* ln_gamma, const float* ln_beta, const float* W, const float* bias,
                    const float* res, float* out_f32, void* out_bf16, int M, int K, int N, int act, void* stream);
int pd_op_cuboid_attention_gv(const void* qkv_bf16, const float* bias_table, const float* gqkv_f32, const void* gqkv_bf16,
                              void* out_bf16, float* gout, int B, int T, int H, int W, int C, int heads, const int32_t size[3],
                              const int32_t strategy[3], const int32_t shift[3], int padding_type, int n_global, int self_attn,
                              void* stream);
/* q_sample (latent_diffusion.py:489-492): out = sqrt_alphas_cumprod[t_b] x_start + sqrt_one_minus_alphas_cumprod[t_b]
 * noise, bit-exact vs the reference's fp32 tensor expression; tables fp32 [T] and t int64 [B] on the device. */
int pd_op_q_sample(const float* x_start, const float* noise, const int64_t* t, const float* sqrt_alphas_cumprod,
                   const float* sqrt_one_minus_alphas_cumprod, float* out, int B, int64_t n_per_sample, void* stream);
int pd_op_sampler_update(float* z, const float* eps, const float* noise, const float* guide, const float* coef8,
                         int64_t n, void* stream);
int pd_op_timestep_embedding(const int64_t* t, float* out, int B, int dim, void* stream);
int pd_op_small_linear(const float* in, const float* W, const float* bias, float* out, int B, int K, int N, int in_silu,
                       int out_silu, void* stream);
int pd_op_pack_conv(const float* w, void* out_bf16, int Co, int Ci, int taps, int Cipad, void* stream);
int pd_op_pack_linear(const float* w, void* out_bf16, int N, int K, int Kpad, void* stream);
int pd_op_upsample2x_cast(const float* x, void* y_bf16, int F, int H, int W, int C, void* stream);
int pd_op_parity_split_cast(const float* x, void* y_bf16, int F, int H, int W, int C, void* stream);
/* stride-2 3x3 conv with the reference's (0,1,0,1) zero pad (taming/resnet.py:183-188) on the parity-split input */
int pd_op_conv_s2_gemm(const void* planes_bf16, const void* Wt_bf16, int F, int Ho, int Wo, int C, int N,
                       const float* bias, float* out_f32, void* stream);

/* pd_op_conv_gemm (fp32 output, no activation) scheduled stream-K: the (tile, k-block) units of every sample are cut
 * into ctas_per_sample equal contiguous ranges, one CTA each; partial tiles go through a global workspace and are summed
 * in a fixed order by the CTA that owns the head of the tile (csrc/gemm_streamk.cu). N must be a multiple of 256; the
 * optional fused LayerNorm needs N == 256. */
int pd_op_conv_gemm_streamk(const void* A_bf16, const void* Wt_bf16, int samples, int D, int H, int W, int C, int kt, int kh,
                            int kw, int N, const float* bias, const float* rowvec, const float* residual, float* out_f32,
                            const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, int ctas_per_sample,
                            void* stream);
/* Same launch with clock64() phase stamps of schedule CTA dbg_cta written to stamps13 (device u64[16]): [0] entry, [1] setup
 * done, [2] first operand stage landed, [3]/[6] last MMA of segment 0/1 issued, [4]/[7] accumulator 0/1 complete, [5] partial
 * dumped, [8] the tile's partials arrived, [9] epilogue done, [10] exit, [11]/[12] %globaltimer ns at entry/exit.
 * Profiling aid (tools/streamk_phases.py). */
int pd_op_conv_gemm_streamk_phases(const void* A_bf16, const void* Wt_bf16, int samples, int D, int H, int W, int C, int kt,
                                   int kh, int kw, int N, const float* bias, const float* rowvec, const float* residual,
                                   float* out_f32, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                                   int ctas_per_sample, int dbg_cta, unsigned long long* stamps13, void* stream);

/* pd_op_conv_gemm with fp32 output (+ bias, + in-place residual) that also ADDS the GroupNorm statistics of the output to
 * gn_sums[samples][groups][2] (double: sum, sum of squares per (sample, group of N / groups channels)) - the table the
 * GroupNorm that follows would otherwise compute in a pass of its own (time_embed.py:116-117 after :93). N % 256 == 0,
 * N / groups in {8, 16, 32}, D*H*W % 32 == 0. streamk_ctas_per_sample > 0 runs the stream-K schedule. */
int pd_op_conv_gemm_gnstats(const void* A_bf16, const void* Wt_bf16, int samples, int D, int H, int W, int C, int kt, int kh,
                            int kw, int N, const float* bias, const float* residual, float* out_f32, double* gn_sums,
                            int groups, int streamk_ctas_per_sample, void* stream);

/* Fused PositionwiseFFN at width 256 / hidden 1024 (cuboid_transformer.py:182-208 after its pre-norm):
 * x[M][256] += W2 GELU(W1 ln_in + b1) + b2 in one kernel (the hidden activation never leaves the SM), and, if ln_gamma
 * is given, ln_out[M][256] (bf16) = LayerNorm(new x row) * ln_gamma + ln_beta. W1 bf16 [1024][256], W2 bf16 [256][1024]. */
int pd_op_ffn_fused(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                    float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, int M, void* stream);
/* Same launch with clock64() stamps of CTA 0 in stamps32 (device u64[32]): [0] entry, [1] MMA warp ready, [2] A tile
 * landed, [3+c] GEMM-2 of chunk c issued, [12+2c] / [13+2c] GELU epilogue of chunk c begins / ends, [28] accumulator 2
 * complete, [29] final epilogue done, [30] exit. Profiling aid (tools/ffn_phases.py). */
int pd_op_ffn_fused_phases(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16,
                           const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                           int M, unsigned long long* stamps32, void* stream);

/* Fused PositionwiseFFN at width 512 / hidden 2048 (the level-1 width of the shipped UNet): x[M][512] += W2 GELU(W1 ln_in +
 * b1) + b2 in one kernel whose hidden dimension is split over a 4-CTA thread-block cluster (each CTA: 512 hidden columns
 * GELU'd into shared memory, a 128 x 512 partial of FFN-2 in TMEM, reduce-scatter of the four partials through
 * distributed shared memory in rank order); if ln_gamma is given, ln_out[M][512] (bf16) = LayerNorm(new x row); if gn_sums
 * is given, the GroupNorm statistics of the new rows are added to gn_sums[samples][gn_groups][2] (gn_rows rows per
 * sample). W1 bf16 [2048][512], W2 bf16 [512][2048]. */
int pd_op_ffn_cluster(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                      float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, double* gn_sums,
                      int gn_groups, int gn_rows, int M, void* stream);
/* pd_op_ffn_cluster with clock64() phase stamps of CTA 0 written to stamps32[32]: [0] entry, [1] dependency wait passed,
 * [2]/[4] G1(c) accumulator complete, [3]/[5] E1(c) done, [6] partial complete, [8] slices
 * sent, [9] cluster barrier passed, [10] rows reduced and written, [11] LayerNorm barrier passed, [12] done; MMA thread: [16]
 * first operands landed, [17]/[18] G1(c) issued, [19]/[20] before / after the wait for E1(1), [21] G2 issued. Profiling aid
 * (tools/ffn_cluster_phases.py). */
int pd_op_ffn_cluster_phases(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16,
                             const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                             int M, unsigned long long* stamps32, void* stream);
/* pd_op_ffn_cluster with the attention output projection fused in front (width 512; CuboidSelfAttentionLayer proj +
 * StackCuboidSelfAttentionBlock residual, cuboid_transformer.py:952,1151, then PositionwiseFFN :182-208):
 *   x1 = x + att Wp^T + bp;  x <- x1 + W2 GELU(W1 LayerNorm(x1; ln1) + b1) + b2;  ln_out = LayerNorm(x; ln) (optional).
 * Each CTA of the 4-CTA cluster forms 128 columns of x1; LayerNorm(x1) crosses L2 through ln_scratch_bf16 [M][512] (may be
 * the ln_out buffer). Extra stamps: [22] first G0 operands landed, [23] G0 issued, [24] x1 complete, [25] E0 done. */
int pd_op_proj_ffn_cluster(const void* att_bf16, const void* Wp_bf16, const float* bp, const float* ln1_gamma,
                           const float* ln1_beta, void* ln_scratch_bf16, const void* W1_bf16, const float* b1,
                           const void* W2_bf16, const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta,
                           void* ln_out_bf16, double* gn_sums, int gn_groups, int gn_rows, int M, unsigned long long* stamps32,
                           void* stream);
/* Fused QKV projection + axial attention core (csrc/qkv_attn.cu): out[B][T][H][W][C] (bf16) = softmax(q k^T / sqrt(hd) +
 * bias) v along `axis` (0 = T, 1 = H, 2 = W; line length <= 16) with q|k|v = ln Wqkv^T formed inside the kernel (rounded to
 * bf16 as the separate QKV GEMM would). ln bf16 [B][T][H][W][C], Wqkv bf16 [3C][C] (rows: q, k, v; head-major inside each),
 * bias_table fp32 [2L-1][heads]. Replaces CuboidSelfAttentionLayer.forward's qkv Linear + attention for the axial cuboids
 * (cuboid_transformer.py:812-861, 949). stamps32 (may be NULL): clock64() phase stamps of CTA (0,0): [0] entry, [1]
 * dependency wait passed, [2] first head's accumulator complete, [3] staged, [4] first head's lines done, [5] all heads
 * done; MMA thread: [16] first operands landed, [17] first head issued (tools/qkv_attn_phases.py). */
int pd_op_qkv_attn(const void* ln_bf16, const void* Wqkv_bf16, const float* bias_table, void* out_bf16, int B, int T, int H,
                   int W, int C, int heads, int axis, unsigned long long* stamps32, void* stream);
/* The same kernel with the attention output projection fused in front (CuboidSelfAttentionLayer proj +
 * StackCuboidSelfAttentionBlock residual, cuboid_transformer.py:952,1151, then PositionwiseFFN :182-208):
 *   x1 = x + att Wp^T + bp;  x <- x1 + W2 GELU(W1 LayerNorm(x1; ln1) + b1) + b2;  ln_out = LayerNorm(x; ln) (optional).
 * ln_scratch_bf16 [M][256]: scratch the normalised tile round-trips through (may be the ln_out buffer). stamps32 may be
 * NULL. */
int pd_op_proj_ffn_fused(const void* att_bf16, const void* Wp_bf16, const float* bp, const float* ln1_gamma,
                         const float* ln1_beta, void* ln_scratch_bf16, const void* W1_bf16, const float* b1,
                         const void* W2_bf16, const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta,
                         void* ln_out_bf16, int M, unsigned long long* stamps32, void* stream);

/* ---- input-gradient kernels of the knowledge-alignment guidance (csrc/backward.cu); used by the parity tests --- */
/* GroupNorm(+SiLU) backward: x, dy fp32 [S][R][C] -> dx_io fp32 (+= if accumulate) and/or dx_bf16 (either may be NULL) */
int pd_op_group_norm_bwd(const float* x, const float* dy, const float* gamma, const float* beta, float* dx_io,
                         void* dx_bf16, int S, int R, int C, int G, float eps, int silu, int accumulate, void* stream);
int pd_op_layer_norm_bwd(const float* x, const float* gamma, const float* dy, float* dx_io, void* dx_bf16, int P, int C,
                         float eps, int accumulate, void* stream);
int pd_op_patch_merge_ln_bwd(const float* x, const float* gamma, const float* dy, float* dx, void* dx_bf16, int BT, int H,
                             int W, int C, float eps, void* stream);
int pd_op_gelu(const float* pre, void* y_bf16, int64_t n, void* stream);
int pd_op_gelu_bwd(const float* pre, const void* dy_bf16, void* dpre_bf16, int64_t n, void* stream);
int pd_op_axial_attention_bwd(const void* qkv_bf16, const float* bias_table, const void* dout_bf16, void* dqkv_bf16, int B,
                              int T, int H, int W, int C, int heads, int axis, void* stream);
/* dgrad operands: fp32 [N][K] -> bf16 [K][N];  fp32 [Co][Ci][taps] -> bf16 [Ci][taps reversed][Co] */
int pd_op_pack_linear_t(const float* w, void* out_bf16, int N, int K, void* stream);
int pd_op_pack_conv_dgrad(const float* w, void* out_bf16, int Co, int Ci, int taps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PREDIFF_B200_H */
