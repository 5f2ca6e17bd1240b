// motioncraft_b200 -- the fused token kernel: EfficientCrossAttention + FFN of one DecoderLayer
// (efficient_attention.py:64-92, diffusion_transformer.py:25-28, stylization_block.py:29-40; called from
// mcm.py:35-40) for a tile of 256 token rows per CTA pair, in ONE persistent kernel.
//
// Every op of the two sub-blocks is row-local (LayerNorm / per-head softmax / AdaLN / SiLU / GELU over the 512 or
// 1024 features of one token) or a GEMM against weights shared by all tokens, so a tile of rows can run the whole
// chain on-chip: activations live in shared memory (as the next GEMM's A operand) and TMEM (accumulators); only the
// fp32 residual stream h is read and written in HBM (2 reads + 2 reduce-adds per row), weights stream from L2.
// The unfused path round-trips ~12 h-sized tensors per layer through HBM for the same work.
#pragma once
#include "common.cuh"

namespace mcm {

struct FusedBlockArgs {
  float* h;                 // [rows, 512] fp32 residual stream, updated in place
  int rows;                 // B * T
  int T;                    // rows per sample
  int batch;                // B
  // cross attention
  const float *ca_ln_w, *ca_ln_b;      // ca_block.norm
  OpPtr ca_wq;  const float* ca_bq;    // ca_block.query
  OpPtr ca_ctxT;                       // [B*4, 128, 128] fp16 per-(sample, head) context, transposed (first sample of this launch)
  const float *ca_pn_w, *ca_pn_b;      // ca_block.proj_out.norm
  const float *ca_scale, *ca_shift;    // AdaLN modulation of this block, row b at + b * mod_ld
  OpPtr ca_wo;  const float* ca_bo;    // ca_block.proj_out.out_layers.2
  // FFN
  OpPtr f_w1;   const float* f_b1;
  OpPtr f_w2;   const float* f_b2;
  const float *f_pn_w, *f_pn_b;
  const float *f_scale, *f_shift;
  OpPtr f_wo;   const float* f_bo;
  int mod_ld;
  void* hid;                // scratch for the GELU'd hidden activations: fused_block_hid_bytes() bytes, private to one stream
  int stop;                 // debug: run only the first `stop` phases of every tile (0 = all) and dump the operand tile
  void* dbg;                // debug: [rows, 512] fp16 dump of the shared-memory operand tile after phase `stop`
};

// The tail of the channel attention (y = softmax(q) ctx -> AdaLN over T -> SiLU -> Linear(T, T) -> residual into h^T) as
// one persistent kernel; see sa_tail_kernel in fused_block.cu.
struct SaTailArgs {
  float* h;             // [B, T, 512] fp32 residual stream (updated in place, transposed accumulation)
  int T, batch;
  OpPtr qs;             // [B*512, Tp] fp16 softmax(q) operand (pad columns zero)
  OpPtr ctxT;           // [B, T, Tp] fp16 block-diagonal per-sample context, transposed
  OpPtr wo;             // [T, Tp] fp16 sa_block.proj_out.out_layers.2.weight
  const float *pn_w, *pn_b, *scale, *shift, *bo;   // proj_out.norm, AdaLN modulation over T (row b at + b * mod_ld), out bias
  int mod_ld;
};
bool sa_tail_supported(int T, int D);
int sa_tail_launch(const SaTailArgs& a, cudaStream_t stream);

// The head of the channel attention (LN_T(h^T) -> q | k | v -> per-head softmax(q); k, v written back transposed) as one
// persistent kernel; see sa_front_kernel in fused_block.cu.  Same shape support as the tail (sa_tail_supported).
struct SaFrontArgs {
  const float* h;       // [B, T, 512] fp32 residual stream (read only)
  int T, batch, heads;
  const float *ln_w, *ln_b;   // sa_block.norm over T
  OpPtr w;              // [3T, Tp] fp16 query | key | value weights
  const float* bqkv;    // [3T]
  OpPtr qs;             // out: [B*512, Tp] fp16 softmax(q) operand
  float* k32;           // out: [B, T, 512] fp32 key (input of the token softmax)
  OpPtr v16;            // out: [B, T, 512] fp16 value operand
};
int sa_front_launch(const SaFrontArgs& a, cudaStream_t stream);

// The middle of the channel attention: softmax of k over the 512 channel-tokens and the per-head context k^T v, written
// block-diagonal and transposed (the B operand of the tail's q ctx GEMM); see sa_ctx_kernel in fused_block.cu.
struct SaCtxArgs {
  const float* k32;     // [B, T, 512] fp32 key (row = (sample, feature), 512 tokens contiguous)
  OpPtr v16;            // [B, T, 512] fp16 value operand
  OpPtr ctxT;           // out: [B, T, Tp] fp16, ctxT[b][l][d]; pad columns stay zero
  int T, batch, heads;
};
int sa_ctx_launch(const SaCtxArgs& a, cudaStream_t stream);

// true if the fused kernel covers this architecture (latent 512, ffn 1024, 4 heads, T >= 52)
bool fused_block_supported(int T, int D, int F, int H);
size_t fused_block_hid_bytes();
// CTA pairs a launch can keep resident (74 on a 148-SM B200), 0 if the kernel cannot run on this device
int fused_block_max_pairs();
int fused_block_launch(const FusedBlockArgs& a, cudaStream_t stream);
unsigned long long fused_block_launch_count();
void fused_block_count_replayed(unsigned long long n);
// MCM_FUSED_PROF=1: summed clock cycles per phase (development aid); out[32]
int fused_block_prof_read(unsigned long long* out, int reset);

}  // namespace mcm
