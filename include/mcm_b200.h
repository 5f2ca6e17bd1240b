/* motioncraft_b200 -- C ABI of the B200-native MotionCraft (configs/mcm) denoising hot path.
 *
 * The reference has NO native boundary: its seam is the mmcv registry + nn.Module call convention
 * (SURVEY.md section 8b).  This header is the boundary a maintainer would bind instead; every entry
 * names the reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C, no torch types: device pointers are raw CUDA pointers (fp32, row-major, contiguous),
 *     `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - every function returns 0 on success, non-zero on error; mcm_last_error() returns the message of
 *     the calling thread's last failure.  Nothing throws, nothing calls exit().
 *   - one context per (device, stream); a context is NOT re-entrant.  The caller keeps OWNERSHIP of all
 *     parameter and activation memory; the context borrows parameter pointers only during
 *     mcm_finalize_params() (it keeps its own packed 16-bit copies; after load_state_dict set every parameter again and
 *     call it again: the copies are re-packed in place, then call mcm_prepare_conditions again).
 *   - requires an sm_100a device.  There is no CPU or non-tcgen05 fallback: creation fails loudly.
 */
#ifndef MCM_B200_H_
#define MCM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcm_ctx mcm_ctx;

/* Architecture of the denoiser, as configs/mcm/mcm_t2m_smplx.py:26-58 spells it. */
typedef struct mcm_config {
  int input_feats;      /* 322                       (input_feats)                               */
  int seq_len;          /* T = max_seq_len; also the SA feature dim (sa_block_cfg.latent_dim)    */
  int latent_dim;       /* 512                                                                   */
  int time_embed_dim;   /* 2048                                                                  */
  int ffn_dim;          /* 1024 (ff_size)                                                        */
  int text_latent_dim;  /* 256                                                                   */
  int num_heads;        /* 4                                                                     */
  int num_layers;       /* 8                                                                     */
  int num_ctrl_blocks;  /* copy_blocks_num of ControlT2MHalf_MCM, 0 = no control branch          */
  int ctrl_cond_feats;  /* in_features of control_cond_input (2048 pre-encoded audio, 35 music)   */
  int max_batch;        /* workspace is sized for this many samples                              */
  int max_text_tokens;  /* 77                                                                    */
  int precise_all;      /* debug: run EVERY GEMM in the 3-pass bf16-split mode (~fp32 accuracy)   */
} mcm_config;

/* Sampler description: float32 copies of the float64 tables of GaussianDiffusion.__init__
 * (mogen/models/utils/gaussian_diffusion.py:354-387), already respaced (SpacedDiffusion.__init__
 * :1416-1431), each of length n_steps, plus timestep_map (:1421-1429).  The float32 cast is the one
 * _extract_into_tensor (:1340) applies. */
typedef struct mcm_sampler {
  int mode;                                   /* 0 = DDIM (ddim_sample :799-852), 1 = DDPM (p_sample :634-696) */
  int n_steps;
  float eta;                                  /* DDIM eta (MotionDiffusion passes 0)                 */
  const int* timestep_map;                    /* host, [n_steps]                                      */
  const float* alphas_cumprod;                /* host, [n_steps]                                      */
  const float* alphas_cumprod_prev;
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  const float* posterior_log_variance_clipped;
  unsigned long long seed;                    /* stochastic samplers without explicit noise: Philox key; step i uses
                                               * stream i, so a run is reproducible from (seed, x_T) alone            */
  int model_mean_type;                        /* 0 = the denoiser predicts eps (configs/mcm/*), 1 = it predicts x_0
                                               * (ModelMeanType.START_X, configs/stmogen/*; gaussian_diffusion.py:555) */
} mcm_sampler;

/* replaces: MCMTransformer.__init__ / DiffusionTransformer.__init__
 * (mogen/models/transformers/diffusion_transformer.py:56-99) -- allocates workspace, no weights yet. */
int mcm_create(const mcm_config* cfg, mcm_ctx** out);
void mcm_destroy(mcm_ctx* ctx);

/* Scheduling options (no effect on results, which are bit-identical in every mode):
 *   "dual"  1 = run the two halves of a batch on two streams (default), 0 = one stream
 *   "graph" 1 = replay a captured CUDA graph per sampler step (default), 0 = eager launches
 *   "chunk" n = pass the batch through the layer stack n samples at a time (default 0 = whole batch)
 *   "fused" 1 = cross-attention + FFN of a layer in one persistent kernel (default; same arithmetic, fp32 reduction
 *               order of the LayerNorm statistics differs), 0 = one kernel per GEMM / row op
 *   "fused_sa" channel attention: 0 = separate kernels, 1 = fused tail kernel, 2 = fused head and tail (default),
 *               3 = also the token softmax + context in one kernel (experimental, slower)
 *   "fused_min_rows" n = use the persistent fused kernels only for launches of at least n rows (B*T, per
 *               stream: half the batch with "dual"); default 2048.  Results of the two schedules agree to ~2e-4, not bit for bit
 *   "fused_sa_min_rows" n = a separate threshold for the fused channel-attention kernels; default -1 = fused_min_rows
 *   "split_sms" 1 = with "dual", every persistent kernel takes half the SMs so the halves run side by side (default 0: slower)
 *   "hoist_mod" 1 (default) = mcm_sample / mcm_sample_host compute the timestep-conditioned AdaLN modulation of ALL steps in one
 *               batched pass before the loop while that table stays below 2 GB (env MCM_HOIST_MOD_MB); 0 = per step always.
 *               Scheduling only: x_0 does not depend on it. */
int mcm_set_option(mcm_ctx* ctx, const char* name, int value);

/* replaces: load_checkpoint / nn.Module.load_state_dict.  `name` is the reference state_dict key
 * (SURVEY.md section 8b), e.g. "temporal_decoder_blocks.3.ca_block.query.weight"; ControlNet keys are
 * "controlnet.{j}.copied_block.*", "controlnet.{j}.before_proj.*", "controlnet.{j}.after_proj.*",
 * "control_cond_input.*" (controlnet_mcm.py:34-53,138-153).  dev_ptr: fp32 device memory. */
int mcm_set_param(mcm_ctx* ctx, const char* name, const float* dev_ptr, long long numel);
/* Pack all parameters into tensor-core operand form.  Fails if a required key is missing. */
int mcm_finalize_params(mcm_ctx* ctx, void* stream);

/* replaces: the step-INVARIANT part of the per-step forward -- the key/value/softmax/K^T V half of every
 * EfficientCrossAttention (efficient_attention.py:74-88), which depends only on xf_out, and
 * ControlT2MHalf_MCM.forward_c (controlnet_mcm.py:155-166), which depends only on c.
 *   xf_out [B, n_tokens, text_latent_dim], xf_proj [B, time_embed_dim]   (get_precompute_condition, mcm.py:58-67)
 *   c      [B, c_len, ctrl_cond_feats] pre-encoded condition or NULL */
int mcm_prepare_conditions(mcm_ctx* ctx, int batch, const float* xf_out, int n_tokens, const float* xf_proj,
                           const float* c, int c_len, void* stream);

/* replaces: DiffusionTransformer.forward + MCMTransformer.forward_test (diffusion_transformer.py:186-238,
 * mcm.py:93-102) or ControlT2MHalf_MCM.forward/forward_test (controlnet_mcm.py:168-233, 306-361) when the
 * context has control blocks and a condition was prepared.
 *   x [B, T, input_feats]; timesteps: device int64 [B] (ORIGINAL 0..999 timesteps) or NULL to use
 *   t_uniform for every sample; eps_out [B, T, input_feats]. */
int mcm_denoise(mcm_ctx* ctx, int batch, const float* x, const long long* timesteps, int t_uniform,
                float* eps_out, void* stream);

/* replaces: one DecoderLayer.forward (mcm.py:25-41) -- the per-block entry ControlT2MHalf_MCM and
 * block-level parity tests use.  kind 0 = temporal_decoder_blocks[index], 1 = controlnet[index].copied_block.
 *   x_inout [B, T, latent_dim] (updated in place), emb [B, time_embed_dim]. */
int mcm_block_forward(mcm_ctx* ctx, int kind, int index, int batch, float* x_inout, const float* emb,
                      void* stream);

/* replaces: MCMTransformer.forward_test(h, src_mask, emb, xf_out) (mcm.py:93-102) and, when control blocks and a
 * condition are prepared, ControlT2MHalf_MCM.forward_test (controlnet_mcm.py:306-361): the decoder layers and `out` on an
 * ALREADY EMBEDDED residual stream.  h [B, T, latent_dim] (read only), emb [B, time_embed_dim] (time embedding + text
 * projection, diffusion_transformer.py:206-213), out [B, T, input_feats].  src_mask is replaced by ones inside the MCM
 * decoder layer (mcm.py:28) and therefore is not an argument; xf_out enters through mcm_prepare_conditions. */
int mcm_layers_forward(mcm_ctx* ctx, int batch, const float* h, const float* emb, float* out, void* stream);

/* replaces: GaussianDiffusion.ddim_sample_loop / p_sample_loop (gaussian_diffusion.py:698-797, 925-1049)
 * with clip_denoised=False, epsilon prediction, fixed_small variance.
 *   x_T [B,T,F] device; x0_out [B,T,F] device.  step_noise: device [n_steps, B, T, F] (index i = retained step i): the
 *   reference's randn_like draws, for reproducing a reference run bit for bit -- or NULL: a stochastic sampler (DDPM, DDIM
 *   with eta != 0) then generates each step's noise on the device (Philox4x32-10 keyed by (s->seed, i) + Box-Muller) into
 *   ONE step-sized buffer, as the reference draws one randn_like per step (:685, :847) instead of n_steps up front. */
int mcm_sample(mcm_ctx* ctx, const mcm_sampler* s, int batch, const float* x_T, const float* step_noise,
               float* x0_out, void* stream);
/* Same through HOST buffers (pinned or pageable): host->device copy of x_T, the loop, device->host copy
 * of x_0, all on `stream`, synchronised before returning.  This is the end-to-end call bench.py times.
 * step_noise_host: host [n_steps, B, T, F] or NULL (generated on the device, as above); a host tensor is copied one
 * step at a time, so device memory stays at one step's worth. */
int mcm_sample_host(mcm_ctx* ctx, const mcm_sampler* s, int batch, const float* x_T_host,
                    const float* step_noise_host, float* x0_out_host, void* stream);

/* RePaint / outpainting description for mcm_sample_repaint (the long-form generation of tools/m2d_test.py:176-195 and
 * tools/s2g_test.py with --repaint: the first frames of a window are pinned to the tail of the previous window). */
typedef struct mcm_repaint {
  int n_times;                      /* length of `times`, or 0 = plain loop over all steps (opt.no_repaint)                */
  const int* times;                 /* host: get_schedule_jump_cjm_ddim (mogen/models/utils/scheduler.py:178-208), ends -1 */
  const float* betas;               /* host [n_steps]: respaced betas, float32 cast (undo, gaussian_diffusion.py:426-435)  */
  const float* gt;                  /* device [B,T,F]: y['gt']                                                              */
  const unsigned char* keep_mask;   /* device [B,T,F]: y['outpainting_mask'], 1 = keep the (noised) ground truth            */
  const float* noise_seq;           /* device [n_draws,B,T,F]: the reference's torch.randn_like draws IN ORDER -- two per
                                     * denoise call (the eta noise of :847, unused at eta = 0, then the blend noise of
                                     * :867), one per undo.  NULL: the draws that are actually read are generated on the
                                     * device (Philox keyed by (sampler seed, draw index)), one buffer, nothing up front   */
  long long n_draws;
  int overlap_len;                  /* opt.overlap_len                                                                      */
  int add_blend;                    /* opt.addBlend                                                                         */
  const float* blend_w;             /* device [overlap_len]: torch.linspace(0, 1, overlap_len) (:873), or NULL              */
} mcm_repaint;

/* replaces: SpacedDiffusion.ddim_sample_loop with y['outpainting_mask'] set (gaussian_diffusion.py:925-997): the
 * harmonising loop ddim_sample_loop_progressive_harmonize (:1050-1118) -- denoise / undo along `times` -- or the plain
 * loop, with the mask blend of ddim_sample (:855-879) after every DDIM update.  eta = 0, same_overlap_noisy = False. */
int mcm_sample_repaint(mcm_ctx* ctx, const mcm_sampler* s, const mcm_repaint* r, int batch, const float* x_T,
                       float* x0_out, void* stream);

/* Result hand-off, on the device before the one device->host copy.
 *
 * replaces: tools/visualize.py:219-249 (motionx branch) -- pred * std + mean, the repack of the 322-dim vector into SMPL-X
 * poses (165) / expressions (100) / translation (3), and scipy.ndimage.gaussian_filter(col, sigma, mode="nearest") per
 * column.  float64 and in scipy's accumulation order: bit-identical to the reference's numpy / scipy result.
 *   pred [B, T, 322] fp32; lengths_dev [B] valid frames (rows beyond stay 0) or NULL; mean / std [322] float64;
 *   denorm_f32 = 1 when the reference's mean / std arrays are float32 (numpy then de-normalises in float32);
 *   w_* : normalised Gaussian taps [2 r + 1] (device), as scipy's _gaussian_kernel1d(sigma, 0, r), r = int(4 sigma + 0.5);
 *   outputs float64 [B, T, 165], [B, T, 100], [B, T, 3]. */
int mcm_handoff_smplx(const float* pred, int B, int T, const int* lengths_dev, const double* mean_dev, const double* std_dev,
                      int denorm_f32, const double* w_pose_dev, int r_pose, const double* w_expr_dev, int r_expr,
                      const double* w_trans_dev, int r_trans, double* pose_out, double* expr_out, double* trans_out,
                      void* stream);
/* replaces: `pred_motion * std + mean` on the host (tools/visualize.py:221, tools/m2d_test.py:203, tools/s2g_test.py:216,
 * MCMTransformer.post_process mcm.py:69-79) with numpy's promotion rules: float32 arithmetic when denorm_f32 = 1 (both
 * arrays float32), float64 otherwise.  out64 [rows, feats] float64 and / or out32 = its float32 rounding (either may be NULL). */
int mcm_handoff_denorm(const float* pred, const double* mean_dev, const double* std_dev, long long rows, int feats,
                       int denorm_f32, double* out64, float* out32, void* stream);
/* replaces: BaseMotionDataset.evaluate's face alignment (mogen/datasets/base_dataset.py:121-125):
 * pred[:, 156:309] = motion[:, 156:309]; pred[:, 312:] = motion[:, 312:]   (rows = B * T, feats = 322). */
int mcm_handoff_align_faces(float* pred, const float* motion, long long rows, int feats, void* stream);

/* Raw tensor-core GEMM, exposed for unit tests of the kernel itself:
 *   C[M,N] (fp32) = A[M,K] * W[N,K]^T + bias[N]   with A, W fp32 device tensors quantised to
 *   fmt 0 = fp16 (1 pass) or 1 = bf16 hi/lo (3 passes). */
int mcm_test_linear(int M, int N, int K, const float* A, const float* W, const float* bias, float* C,
                    int fmt, void* stream);

/* replaces: the classifier-free-guidance combine of STMoGenTransformer.forward_test (stmogen.py:755-759, with scale_func
 * :655-659): out = out_text * text_coef + out_none * none_coef, the Python-float coefficients cast to float32 as torch does. */
int mcm_cfg_combine(const float* out_text, const float* out_none, double text_coef, double none_coef, float* out,
                    long long n, void* stream);

/* replaces: the static human-topology branch of STMA.forward (mogen/models/attentions/st_attention.py:123-128):
 * out[r, h, :] = sum_l softmax(body_weight, dim=1)[h, l] * v[r, l, :];  body_weight [H, H], v / out [rows, H, part_dim]. */
int mcm_part_mix(const float* body_weight, const float* v, float* out, long long rows, int num_parts, int part_dim, void* stream);

/* replaces: SFFN.forward of the STMoGen family (mogen/models/transformers/stmogen.py:581-607) incl. its StylizationBlock
 * (mogen/models/utils/stylization_block.py:29-40):
 *   y[:, :, h] = linear2_h(GELU(linear1_h(x[:, :, h])))  for the H body parts (x viewed as [B, T, H, L]) -- two block-diagonal
 *   tcgen05 GEMMs (one launch each, per-part biases), then  out = x + Linear(SiLU(LN(y) (1 + scale) + shift))  with
 *   (scale | shift) = Linear(SiLU(emb)).
 * All pointers are device fp32: x / out [B, T, H*L]; emb [B, E]; w1 [H, F, L], b1 [H, F]; w2 [H, L, F], b2 [H, L];
 * emb_w [2*H*L, E], emb_b [2*H*L]; ln_w / ln_b [H*L]; out_w [H*L, H*L], out_b [H*L].  H*L <= 1024, L and F multiples of 8.
 * Weights are packed per call (a first Path-B piece: functional and parity-pinned, not yet a resident operator). */
int mcm_sffn_forward(int B, int T, int H, int L, int F, int E, const float* x, const float* emb, const float* w1, const float* b1,
                     const float* w2, const float* b2, const float* emb_w, const float* emb_b, const float* ln_w, const float* ln_b,
                     const float* out_w, const float* out_b, float* out, void* stream);

/* replaces: everything of STMA.forward (mogen/models/attentions/st_attention.py:105-175) AFTER its two mixture-of-experts
 * layers (tutel, un-vendored: not part of this library): the MoE outputs are inputs here.
 *   motion_feat [B, T, H, 4L] = motion_moe(norm(x)) : (body value | key | value | query) per body part
 *   text_feat   [B, Nt, Ht, 2L] = text_moe(text_norm(xf)) : (key | value); Ht = 1 (repeated over the parts, :150, :158) or H
 *   y_s = softmax(body_weight, 1) mixed body values (static_body, :123-128) [+ the 8-head EfficientSelfAttention over the H part
 *         tokens of every frame (dynamic body, :130-135; dyn_* = its LayerNorm and stacked q | k | v Linear, NULL = off)]
 *   y_t = softmax_L(query) (softmax_n([key_text + (1-text_cond) -1e6 ; key_motion + (1-src_mask) -1e6])^T [value_text text_cond ;
 *         value_motion src_mask])   per (sample, part)                                            (:146-171)
 *   out = x + StylizationBlock(y_s + y_t, emb)                                                     (:172)
 * All pointers device fp32; x / out [B, T, H*L]; emb [B, E]; src_mask [B, T]; text_cond [B] (= cond_type % 10 > 0);
 * body_weight [H, H]; dyn_wqkv [3L, L], dyn_bqkv [3L]; StylizationBlock parameters as in mcm_sffn_forward.
 * L in {32, 64, 96, 128}, H*L <= 1024, Nt + T <= 1024.  Weights are packed per call. */
int mcm_stma_mix(int B, int T, int H, int L, int Nt, int Ht, int E, int static_body, const float* x, const float* motion_feat,
                 const float* text_feat, const float* emb, const float* src_mask, const float* text_cond, const float* body_weight,
                 const float* dyn_ln_w, const float* dyn_ln_b, const float* dyn_wqkv, const float* dyn_bqkv, const float* emb_w,
                 const float* emb_b, const float* ln_w, const float* ln_b, const float* out_w, const float* out_b, float* out,
                 void* stream);

/* The on-device noise generator of the stochastic samplers, exposed for unit tests: out[0..n) ~ N(0,1), Philox4x32-10
 * keyed by `seed`, counter (element quad, sub), Box-Muller. */
int mcm_test_randn(float* out_dev, long long n, unsigned long long seed, unsigned long long sub, void* stream);

/* Measurement aid for bench.py's roofline leg: when enabled, every kernel launch of the library is
 * bracketed by CUDA events on its stream; collect() synchronises the device and returns, for class 0
 * (tcgen05 GEMM), class 1 (row kernels) and class 2 (fused cross-attention + FFN kernel), the summed device time
 * [ms], launch count and algorithmic flops, then clears the record.  Arrays have 3 entries. */
void mcm_timing_enable(int on);
int mcm_timing_collect(double* ms, unsigned long long* launches, double* flops);

/* Development aid for the fused cross-attention + FFN kernel: after mcm_set_option(ctx, "fused_stop", k) every tile runs
 * only its first k phases; what = 0 copies the fp16 dump of the shared-memory operand tile [B*T, 512], what = 1 the
 * hidden-activation scratch, into device memory dst_dev. */
int mcm_debug_copy(mcm_ctx* ctx, int what, void* dst_dev, long long bytes);
/* Development aid (MCM_DEBUG_EPI=3): summed clock cycles of the GEMM epilogue warps per phase; out[16]. */
int mcm_debug_read(unsigned long long* out, int reset);
/* Development aid (MCM_FUSED_PROF=1): summed clock cycles of the fused kernel per phase; out[32]. */
int mcm_debug_read32(unsigned long long* out, int reset);

const char* mcm_last_error(void);
/* kernels launched by this library since process start (tcgen05 GEMMs, all kernels) */
unsigned long long mcm_gemm_launches(void);
unsigned long long mcm_kernel_launches(void);
const char* mcm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCM_B200_H_ */
