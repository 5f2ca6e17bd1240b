// motioncraft_b200 -- HBM-bound row kernels between the GEMMs: LayerNorm (+AdaLN modulation +SiLU),
// segment softmax, the transposing LayerNorm of the channel-attention, operand packing, timestep
// embedding and the DDIM / DDPM update.  All statistics in fp32; outputs are 16-bit GEMM operands.
#pragma once
#include "common.cuh"

namespace mcm {

// LayerNorm over the last dim of in[rows, d] (pitch ld_in), eps 1e-5, affine (w, b), then optionally
//   y = y * (1 + scale[batch, :]) + shift[batch, :]    (StylizationBlock, stylization_block.py:38)
//   y = SiLU(y)                                          (out_layers[0], :21)
// batch = row / rows_per_batch; scale/shift have pitch mod_ld.  Output: operand (rows x out.ld), pad 0.
int ln_rows_launch(const float* in, int rows, int d, int ld_in, const float* w, const float* b,
                   const float* scale, const float* shift, int mod_ld, int rows_per_batch, bool act_silu,
                   OpPtr out, int out_fmt, cudaStream_t stream);

// softmax over contiguous segments of length `seg` of every row of in[rows, ncols] (ncols % seg == 0);
// output operand (rows x out.ld), pad columns zero.
int softmax_seg_launch(const float* in, int rows, int ncols, int ld_in, int seg, OpPtr out, int out_fmt,
                       cudaStream_t stream);

// h[B, T, D] fp32 -> LayerNorm over T (per (b, d) column, affine w[T], b[T]) written TRANSPOSED as an
// operand of shape (B*D rows, T cols): row b*D+d.   (EfficientSelfAttention.norm on x^T, mcm.py:28-32)
int ln_transpose_launch(const float* h, int B, int T, int D, const float* w, const float* b, OpPtr out,
                        int out_fmt, cudaStream_t stream);

// plain fp32 [rows, cols] -> operand (optionally SiLU first), pad columns zero
int pack_op_launch(const float* in, int rows, int cols, int ld_in, bool act_silu, OpPtr out, int out_fmt,
                   cudaStream_t stream);

// sinusoidal timestep embedding (position_encoding.py:42-60) -> operand [B, dim]
int timestep_embedding_launch(const long long* t_dev, int t_uniform, int B, int dim, OpPtr out, int out_fmt,
                              cudaStream_t stream);

// start_x: the denoiser predicts x_0 itself (ModelMeanType.START_X, gaussian_diffusion.py:555-556) instead of eps
struct DdimCoefs { float c1, c2, alpha_bar, alpha_bar_prev, eta; int add_noise; int start_x; };
struct DdpmCoefs { float c1, c2, pm1, pm2, log_var; int add_noise; int start_x; };
// x <- DDIM / DDPM update from eps (fp32, n elements); writes x_out (may alias x) and, if xop.hi, the
// operand copy [rows, cols -> xop.ld] the next step's joint_embed GEMM reads.
int ddim_update_launch(const float* x, const float* eps, const float* noise, float* x_out, size_t rows, int cols,
                       DdimCoefs c, OpPtr xop, int op_fmt, cudaStream_t stream);
int ddpm_update_launch(const float* x, const float* eps, const float* noise, float* x_out, size_t rows, int cols,
                       DdpmCoefs c, OpPtr xop, int op_fmt, cudaStream_t stream);

// RePaint / outpainting (gaussian_diffusion.py:855-879, :426-435): mask blend after a DDIM update, and the re-noising
// "undo" step of the harmonising loop.  Both update x in place and rewrite its operand copy.
int repaint_blend_launch(float* x, const float* gt, const unsigned char* keep, const float* noise, size_t rows, int cols,
                         int T, float gt_w, float noise_w, const float* blend_w, int overlap, OpPtr xop, int op_fmt,
                         cudaStream_t stream);
int undo_launch(float* x, const float* noise, size_t rows, int cols, float a, float b, OpPtr xop, int op_fmt,
                cudaStream_t stream);

// t_buf[0..B) = t  (the timestep of the current sampler step, read by the graph-captured timestep embedding)
int fill_timesteps_launch(long long* t_buf, long long t, int B, cudaStream_t stream);
// out[0..n) = src[idx_dev[0] * n + (0..n)) (n % 4 == 0): a per-step table slice selected by a device-resident index
int gather_slice_launch(const float* src, const long long* idx_dev, float* out, size_t n, cudaStream_t stream);

// out[0..n) ~ N(0, 1): Philox4x32-10 keyed by (seed, sub) + Box-Muller; `sub` numbers the draw (sampler step / RePaint draw)
int randn_fill_launch(float* out, size_t n, unsigned long long seed, unsigned long long sub, cudaStream_t stream);

// out = a * wa + b * wb in fp32 without FMA contraction: the classifier-free-guidance combine of STMoGenTransformer.forward_test
// (stmogen.py:755-759), out_text * text_coef + out_none * none_coef
int axpby_launch(const float* a, const float* b, float wa, float wb, float* out, size_t n, cudaStream_t stream);

// out[r, h, :] = sum_l softmax(body_weight (H x H), dim=1)[h, l] * v[r, l, :]   (STMA static branch, st_attention.py:123-128)
// `pitch`: elements between consecutive parts of the INPUT (0 = L, dense); the output is dense
int part_mix_launch(const float* body_weight, const float* v, float* out, size_t rows, int H, int L, cudaStream_t stream,
                    int pitch = 0);

int elementwise_init();
unsigned long long elementwise_launch_count();
void elementwise_count_replayed(unsigned long long n);

}  // namespace mcm
