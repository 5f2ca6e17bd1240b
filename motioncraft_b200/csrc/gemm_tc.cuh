// motioncraft_b200 -- the tcgen05 GEMM that carries every dense contraction of the denoiser.
//
//   D[128 x BN] (fp32, TMEM) = A_tile[128 x K] * B_tile[BN x K]^T     (both operands K-major, 16-bit)
//
// A "problem" is a set of independent batches (samples, or sample x head), each an (M x N) output;
// N is split into up to 3 SEGMENTS (e.g. q | k | v) that share the A operand but have their own bias,
// destination, layout and activation.  Everything that follows a Linear in the reference as a pointwise
// op (bias, residual / positional addend, GELU / SiLU, block-diagonal head mask, transposition, the
// 16-bit re-quantisation for the next GEMM) is done in the epilogue on the fp32 accumulator.
#pragma once
#include "common.cuh"

namespace mcm {

enum EpiFlags : int {
  EPI_TRANSPOSED = 1,     // element (r, c) of batch `outer` goes to [(outer*trans_rows + c) * ld + r]
  EPI_GELU = 2,           // exact erf GELU (nn.GELU default, diffusion_transformer.py:21)
  EPI_SILU = 4,
  EPI_MASK_BLOCKDIAG = 8, // zero unless r / head_dim == c / head_dim (per-head contexts)
  EPI_ADDEND_BCAST = 16,  // addend row = r (row within batch), e.g. the positional embedding
};

struct EpiSeg {
  const float* bias;    // [n] or nullptr
  const float* addend;  // fp32, addressed like out32 (may alias out32 for a residual add) or nullptr
  float* out32;         // fp32 destination or nullptr
  OpPtr op;             // 16-bit operand destination (hi == nullptr: none)
  int n;                // logical columns of this segment
  int w_row0;           // first row of the B operand (weight) belonging to this segment
  int col0;             // first output column
  int ld32;             // pitch of out32 / addend (elements)
  int flags;
  int op_fmt;           // OpFormat of `op`
  int tile0, n_tiles;   // filled by the launcher
  int n_pad;            // filled by the launcher: columns to cover (n rounded up to 8 if op)
  int vec32;            // filled by the launcher: 4 if fp32 row-major accesses may be float4
};

struct GemmProblem {
  // operands
  OpPtr a;  int a_rows;  int a_k;  int a_batches;   // A tensor: (a_batches, a_rows, a_k[pitch a.ld])
  OpPtr b;  int b_rows;  int b_k;  int b_batches;   // B tensor: (b_batches, b_rows, b_k[pitch b.ld])
  int fmt;               // OpFormat of BOTH operands (OP_F16: 1 pass, OP_BF16X2: 3 passes)
  // iteration space
  int M;                 // rows per batch
  int K;                 // contraction length
  int batches;           // outer * inner
  int inner;             // e.g. heads; batch = outer * inner + inner_idx
  int a_k_inner;         // A k-offset per inner index (per-head slices of a shared A)
  int b_k_inner;         // B k-offset per inner index (per-head diagonal blocks of a block-diagonal B)
  int a_batched;         // 0: A batch index = outer; 1: A batch index = batch (A described per (outer, inner))
  int out_batched;       // 1: output row = batch * out_rows_per_outer + r (columns still + inner * out_col_inner)
  int bias_inner;        // bias offset per inner index (per-group biases of a block-diagonal Linear); 0: one bias per segment
  int b_batched;         // 0: B shared by all batches (weights); 1: B batch index = batch
  int out_col_inner;     // output column offset per inner index
  int out_rows_per_outer;// output row = outer * out_rows_per_outer + r
  int trans_rows;        // see EPI_TRANSPOSED
  int head_dim;          // see EPI_MASK_BLOCKDIAG
  int nseg;
  EpiSeg seg[3];
  double algo_flops;     // ALGORITHMIC flops of this launch for the roofline record; 0 = 2*M*N*K*batches
};

// Launch on `stream`.  Returns 0 or an error status (message via mcm_last_error()).
int gemm_tc_launch(const GemmProblem& prob, cudaStream_t stream);

// device-wide one-time init (driver entry point for cuTensorMapEncodeTiled, smem attribute, SM count)
int gemm_tc_init();

// tensor-map builders shared with the fused block kernel.
//   operand map: 16-bit K-major rows (fmt = OpFormat), dims {k_dim, rows, batches}, pitch ld, box {64, box_rows, 1}, 128B swizzle
//   tile map:    dims {d0, d1, d2} (d0 contiguous), 32 x 32 x 1 boxes; dtype 0 = fp32, 1 = fp16; swizzle 128 / 64 / 0 bytes
int tc_make_operand_map(CUtensorMap* m, const void* ptr, int fmt, int k_dim, int rows, int batches, int ld, int box_rows);
int tc_make_tile_map(CUtensorMap* m, const void* ptr, int dtype, long long d0, long long d1, long long d2,
                     long long stride1_elems, long long stride2_elems, int swizzle_bytes);
//   box map:     like the tile map but with an explicit {box0, box1, 1} box and no swizzle (row-staged stores)
// Persistent kernels launched while share = n use 1 / n of the SMs, so that the kernels of n streams run side by side.
void tc_set_sm_share(int share);
int tc_sm_share();
int tc_make_box_map(CUtensorMap* m, const void* ptr, int dtype, long long d0, long long d1, long long d2,
                    long long stride1_elems, long long stride2_elems, int box0, int box1);
int tc_num_sms();

// number of tcgen05 GEMM launches since process start (bench.py's gpu_launches evidence)
unsigned long long gemm_tc_launch_count();
void gemm_tc_count_replayed(unsigned long long n);   // kernels executed by a CUDA-graph replay
// MCM_DEBUG_EPI=3: summed epilogue-warp cycles per phase (development aid)
int gemm_tc_debug_read(unsigned long long* out, int reset);

}  // namespace mcm
