// motioncraft_b200 -- device side of the Path-B (STMoGen, configs/stmogen/*) pieces that the reference can pin in this image
// (SURVEY.md section 8, row f-1): SFFN and everything of STMA that follows its two mixture-of-experts layers.  The MoE itself
// (tutel.moe.moe_layer, st_attention.py:17-56) is an un-vendored, unpinned dependency and is NOT here: `mcm_stma_mix` takes the
// MoE outputs (motion_feat, text_feat) as inputs.
//
// Both entries are composed from the library's kernels -- the tcgen05 GEMM with its per-(sample, part) batch maps, the row
// kernels -- plus two small kernels of their own (the transposing key / value pack with the reference's masks, the 12-token
// dynamic body attention).  Weights are packed per call: functional and parity-pinned, not yet resident operators.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcm_b200.h"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm_tc.cuh"

namespace mcm {
namespace {

EpiSeg seg_of(int n) {
  EpiSeg s;
  std::memset(&s, 0, sizeof(s));
  s.n = n;
  return s;
}
GemmProblem shared_weight_problem(const OpPtr& a, int rows, const OpPtr& w, int w_rows, int K, int fmt) {
  GemmProblem g;
  std::memset(&g, 0, sizeof(g));
  g.a = a; g.a_rows = rows; g.a_k = a.ld; g.a_batches = 1;
  g.b = w; g.b_rows = w_rows; g.b_k = w.ld; g.b_batches = 1;
  g.fmt = fmt;
  g.M = rows; g.K = K; g.batches = 1; g.inner = 1;
  g.out_rows_per_outer = rows;
  g.nseg = 1;
  return g;
}

// scratch of one call, freed when it goes out of scope (after the stream was synchronised)
struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() { for (void* q : ptrs) cudaFree(q); }
  int bytes(void** out, size_t n) {
    MCM_CUDA(cudaMalloc(out, n < 256 ? 256 : n));
    ptrs.push_back(*out);
    return 0;
  }
  int f32(float** out, size_t n) { return bytes(reinterpret_cast<void**>(out), n * 4); }
  int op(OpPtr* o, size_t elems, int ld, bool lo) {
    o->ld = ld; o->lo = nullptr;
    MCM_TRY(bytes(&o->hi, elems * 2));
    if (lo) MCM_TRY(bytes(&o->lo, elems * 2));
    return 0;
  }
};

struct StyleParams {
  const float *emb_w, *emb_b, *ln_w, *ln_b, *out_w, *out_b;
};
// out = x + Linear(SiLU(LN(y) (1 + scale) + shift)),  (scale | shift) = Linear(SiLU(emb))      stylization_block.py:29-40
// y / x / out [B*T, D]; emb [B, E].  The AdaLN emb GEMM runs in the 3-pass bf16 split, the output GEMM in fp16 operands with
// the residual TMA-loaded into its epilogue -- the precision classes of the configs/mcm path.
int stylization_tail(Scratch& sc, int B, int T, int D, int E, const float* y, const float* x, const float* emb, const StyleParams& sp,
                     float* out, cudaStream_t st) {
  MCM_CHECK(D % 8 == 0 && D <= 1024 && E % 8 == 0, "stylization: need D <= 1024, D and E multiples of 8");
  const int rows = B * T;
  OpPtr embp, ewp, zop, owp;
  float* mod = nullptr;
  MCM_TRY(sc.op(&embp, (size_t)B * E, E, true));
  MCM_TRY(sc.op(&ewp, (size_t)2 * D * E, E, true));
  MCM_TRY(sc.op(&zop, (size_t)rows * D, D, false));
  MCM_TRY(sc.op(&owp, (size_t)D * D, D, false));
  MCM_TRY(sc.f32(&mod, (size_t)B * 2 * D));
  MCM_TRY(pack_op_launch(emb, B, E, E, true, embp, OP_BF16X2, st));      // SiLU(emb)
  MCM_TRY(pack_op_launch(sp.emb_w, 2 * D, E, E, false, ewp, OP_BF16X2, st));
  MCM_TRY(pack_op_launch(sp.out_w, D, D, D, false, owp, OP_F16, st));
  {
    GemmProblem g = shared_weight_problem(embp, B, ewp, 2 * D, E, OP_BF16X2);
    g.seg[0] = seg_of(2 * D);
    g.seg[0].bias = sp.emb_b; g.seg[0].out32 = mod; g.seg[0].ld32 = 2 * D;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  MCM_TRY(ln_rows_launch(y, rows, D, D, sp.ln_w, sp.ln_b, mod, mod + D, 2 * D, T, true, zop, OP_F16, st));
  {
    GemmProblem g = shared_weight_problem(zop, rows, owp, D, D, OP_F16);
    g.seg[0] = seg_of(D);
    g.seg[0].bias = sp.out_b; g.seg[0].addend = x; g.seg[0].out32 = out; g.seg[0].ld32 = D;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// STMA key / value pack (st_attention.py:146-161): the reference concatenates text and motion tokens along n, masks them and
// soft-maxes the keys over n.  Here both tensors are written TRANSPOSED -- row = (part h, feature), column = token n -- so that
// the softmax over tokens is a row softmax and both are K-major operands of the context GEMM:
//   keyT[b, h*L + d, n] = n < Nt ? text_feat[b, n, ht, d]     + (1 - text_cond[b]) * -1e6
//                                : motion_feat[b, n-Nt, h, L + d] + (1 - src_mask[b, n-Nt]) * -1e6           (fp32)
//   valT[b, h*L + l, n] = n < Nt ? text_feat[b, n, ht, L + l] * text_cond[b] : motion_feat[b, n-Nt, h, 2L + l] * src_mask[b, n-Nt]
// (fp16 operand, zero for n >= N).  ht = 0 when the text has one head (`key_text.repeat(1, 1, H, 1)`, :150-151, :158-159).
// grid (ceil(Np / 32), H*L / 32, B), block (32, 8); 32 x 32 tiles through shared memory, both directions coalesced.
__global__ void __launch_bounds__(256)
stma_kv_pack_kernel(const float* __restrict__ motion_feat, const float* __restrict__ text_feat, const float* __restrict__ src_mask,
                    const float* __restrict__ text_cond, int T, int H, int L, int Nt, int Ht, int Np, float* __restrict__ keyT,
                    uint16_t* __restrict__ valT) {
  __shared__ float tk[32][33], tv[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const int h = c0 / L, d0 = c0 - h * L;            // L % 32 == 0: a tile stays inside one part
  const int N = Nt + T;
  const float tc = text_cond[b];
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i;
    float k = 0.f, v = 0.f;
    if (n < Nt) {
      const float* src = text_feat + (((size_t)b * Nt + n) * Ht + (Ht == 1 ? 0 : h)) * (size_t)(2 * L) + d0 + tx;
      k = src[0] + (1.f - tc) * -1000000.f;
      v = src[L] * tc;
    } else if (n < N) {
      const int t = n - Nt;
      const float m = src_mask[(size_t)b * T + t];
      const float* src = motion_feat + (((size_t)b * T + t) * H + h) * (size_t)(4 * L) + d0 + tx;
      k = src[L] + (1.f - m) * -1000000.f;
      v = src[2 * L] * m;
    }
    tk[i][tx] = k;
    tv[i][tx] = v;
  }
  __syncthreads();
  const int n = n0 + tx;
  if (n < Np) {
    for (int i = ty; i < 32; i += 8) {
      const size_t o = ((size_t)b * H * L + c0 + i) * Np + n;
      keyT[o] = tk[tx][i];
      valT[o] = f32_to_f16_bits(n < N ? tv[tx][i] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// STMA dynamic body branch (st_attention.py:130-135): EfficientSelfAttention(latent L, 8 heads, no stylization,
// efficient_attention.py:25-46) over the H (= 12) body-part tokens of ONE frame, mask all ones:
//   q = softmax_hd(q), k = softmax over the H tokens, att[g] = k[:, g]^T v[:, g]  (hd x hd per head g), y = x + q att.
// qkv [rows, H, 3L] fp32 = LN(x) Wqkv^T + b (computed by the GEMM), x = body_value = motion_feat[..., :L] (pitch 4L).
// One block per frame; result ADDED into y_s[row, h, :] (which already holds the static mix).
template <int HEADS>
__global__ void __launch_bounds__(128)
stma_dyn_body_kernel(const float* __restrict__ qkv, const float* __restrict__ motion_feat, int H, int L, float* __restrict__ ys) {
  extern __shared__ float sm[];
  float* q = sm;                        // [H][L]
  float* k = q + H * L;                 // [H][L]
  float* v = k + H * L;                 // [H][L]
  float* att = v + H * L;               // [HEADS][hd][hd]
  const int hd = L / HEADS;
  const size_t row = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < H * L; i += nt) {
    const int tok = i / L, c = i - tok * L;
    const float* src = qkv + (row * H + tok) * (size_t)(3 * L);
    q[i] = src[c]; k[i] = src[L + c]; v[i] = src[2 * L + c];
  }
  __syncthreads();
  // query: softmax over the hd features of each (token, head)
  for (int i = tid; i < H * HEADS; i += nt) {
    float* p = q + (i / HEADS) * L + (i % HEADS) * hd;
    float m = -INFINITY;
    for (int j = 0; j < hd; ++j) m = fmaxf(m, p[j]);
    float s = 0.f;
    for (int j = 0; j < hd; ++j) { p[j] = expf(p[j] - m); s += p[j]; }
    for (int j = 0; j < hd; ++j) p[j] /= s;
  }
  // key: softmax over the H tokens of each feature column
  for (int c = tid; c < L; c += nt) {
    float m = -INFINITY;
    for (int tok = 0; tok < H; ++tok) m = fmaxf(m, k[tok * L + c]);
    float s = 0.f;
    for (int tok = 0; tok < H; ++tok) { const float e = expf(k[tok * L + c] - m); k[tok * L + c] = e; s += e; }
    for (int tok = 0; tok < H; ++tok) k[tok * L + c] /= s;
  }
  __syncthreads();
  for (int i = tid; i < HEADS * hd * hd; i += nt) {
    const int g = i / (hd * hd), r = i - g * hd * hd, d = r / hd, l = r - d * hd;
    float a = 0.f;
    for (int tok = 0; tok < H; ++tok) a = fmaf(k[tok * L + g * hd + d], v[tok * L + g * hd + l], a);
    att[i] = a;
  }
  __syncthreads();
  for (int i = tid; i < H * L; i += nt) {
    const int tok = i / L, c = i - tok * L, g = c / hd, l = c - g * hd;
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(q[tok * L + g * hd + d], att[(g * hd + d) * hd + l], a);
    const float x = motion_feat[(row * H + tok) * (size_t)(4 * L) + c];
    ys[(row * H + tok) * (size_t)L + c] += x + a;
  }
}

}  // namespace
}  // namespace mcm

using namespace mcm;

extern "C" {

// SFFN of the STMoGen family (stmogen.py:581-607) + its StylizationBlock (stylization_block.py:29-40); see the header.
int mcm_sffn_forward(int B, int T, int H, int L, int F, int E, const float* x, const float* emb, const float* w1, const float* b1,
                     const float* w2, const float* b2, const float* emb_w, const float* emb_b, const float* ln_w, const float* ln_b,
                     const float* out_w, const float* out_b, float* out, void* stream) {
  MCM_CHECK(B > 0 && T > 0 && H > 0 && L > 0 && F > 0 && E > 0, "mcm_sffn_forward: bad shape");
  MCM_CHECK(x && emb && w1 && b1 && w2 && b2 && emb_w && emb_b && ln_w && ln_b && out_w && out_b && out, "mcm_sffn_forward: null pointer");
  const int D = H * L, rows = B * T;
  MCM_CHECK(L % 8 == 0 && F % 8 == 0 && E % 8 == 0 && D <= 1024, "mcm_sffn_forward: need L, F, E multiples of 8 and H*L <= 1024");
  MCM_TRY(gemm_tc_init());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Scratch sc;
  auto run = [&]() -> int {
    OpPtr xa, w1p, hid, w2p;
    float* y = nullptr;
    MCM_TRY(sc.op(&xa, (size_t)rows * D, D, false));
    MCM_TRY(sc.op(&w1p, (size_t)H * F * L, L, false));
    MCM_TRY(sc.op(&hid, (size_t)rows * H * F, H * F, false));
    MCM_TRY(sc.op(&w2p, (size_t)H * L * F, F, false));
    MCM_TRY(sc.f32(&y, (size_t)rows * D));
    MCM_TRY(pack_op_launch(x, rows, D, D, false, xa, OP_F16, st));
    MCM_TRY(pack_op_launch(w1, H * F, L, L, false, w1p, OP_F16, st));
    MCM_TRY(pack_op_launch(w2, H * L, F, F, false, w2p, OP_F16, st));
    {  // hid[:, h*F : (h+1)*F] = GELU(x[:, h*L : (h+1)*L] W1_h^T + b1_h): one block-diagonal launch, batch = part
      GemmProblem g;
      std::memset(&g, 0, sizeof(g));
      g.a = xa; g.a_rows = rows; g.a_k = D; g.a_batches = 1; g.a_k_inner = L;
      g.b = w1p; g.b_rows = F; g.b_k = L; g.b_batches = H; g.b_batched = 1;
      g.fmt = OP_F16; g.M = rows; g.K = L; g.batches = H; g.inner = H;
      g.out_rows_per_outer = rows; g.out_col_inner = F; g.bias_inner = F;
      g.nseg = 1;
      g.seg[0] = seg_of(F);
      g.seg[0].bias = b1; g.seg[0].op = hid; g.seg[0].op_fmt = OP_F16; g.seg[0].flags = EPI_GELU;
      MCM_TRY(gemm_tc_launch(g, st));
    }
    {  // y[:, h*L : (h+1)*L] = hid[:, h*F : (h+1)*F] W2_h^T + b2_h
      GemmProblem g;
      std::memset(&g, 0, sizeof(g));
      g.a = hid; g.a_rows = rows; g.a_k = H * F; g.a_batches = 1; g.a_k_inner = F;
      g.b = w2p; g.b_rows = L; g.b_k = F; g.b_batches = H; g.b_batched = 1;
      g.fmt = OP_F16; g.M = rows; g.K = F; g.batches = H; g.inner = H;
      g.out_rows_per_outer = rows; g.out_col_inner = L; g.bias_inner = L;
      g.nseg = 1;
      g.seg[0] = seg_of(L);
      g.seg[0].bias = b2; g.seg[0].out32 = y; g.seg[0].ld32 = D;
      MCM_TRY(gemm_tc_launch(g, st));
    }
    const StyleParams sp{emb_w, emb_b, ln_w, ln_b, out_w, out_b};
    return stylization_tail(sc, B, T, D, E, y, x, emb, sp, out, st);
  };
  int rc = run();
  const cudaError_t e = cudaStreamSynchronize(st);
  if (!rc && e != cudaSuccess) {
    set_error(std::string("mcm_sffn_forward: ") + cudaGetErrorString(e));
    rc = 1;
  }
  return rc;
}

// Everything of STMA.forward (st_attention.py:105-175) after its two mixture-of-experts layers; see the header.
int mcm_stma_mix(int B, int T, int H, int L, int Nt, int Ht, int E, int static_body, const float* x, const float* motion_feat,
                 const float* text_feat, const float* emb, const float* src_mask, const float* text_cond, const float* body_weight,
                 const float* dyn_ln_w, const float* dyn_ln_b, const float* dyn_wqkv, const float* dyn_bqkv, const float* emb_w,
                 const float* emb_b, const float* ln_w, const float* ln_b, const float* out_w, const float* out_b, float* out,
                 void* stream) {
  MCM_CHECK(B > 0 && T > 0 && H > 0 && H <= 32 && L > 0 && Nt > 0 && E > 0 && (Ht == 1 || Ht == H), "mcm_stma_mix: bad shape");
  MCM_CHECK(x && motion_feat && text_feat && emb && src_mask && text_cond && body_weight && emb_w && emb_b && ln_w && ln_b && out_w &&
                out_b && out, "mcm_stma_mix: null pointer");
  const int D = H * L, rows = B * T, N = Nt + T, Np = (N + 7) / 8 * 8;
  MCM_CHECK(L % 32 == 0 && L <= 128 && D <= 1024 && N <= 1024, "mcm_stma_mix: need L in {32, 64, 96, 128}, H*L <= 1024, Nt + T <= 1024");
  const bool dyn = dyn_ln_w != nullptr;
  if (dyn) MCM_CHECK(dyn_ln_b && dyn_wqkv && dyn_bqkv && L % 8 == 0, "mcm_stma_mix: incomplete dynamic-body parameters");
  MCM_TRY(gemm_tc_init());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Scratch sc;
  auto run = [&]() -> int {
    float *keyT = nullptr, *ys = nullptr, *ysum = nullptr;
    OpPtr valT, keyS, attT, qS;
    MCM_TRY(sc.f32(&keyT, (size_t)B * D * Np));
    MCM_TRY(sc.op(&valT, (size_t)B * D * Np, Np, false));
    MCM_TRY(sc.op(&keyS, (size_t)B * D * Np, Np, false));
    MCM_TRY(sc.op(&attT, (size_t)B * H * L * L, L, false));
    MCM_TRY(sc.op(&qS, (size_t)rows * D, L, false));
    MCM_TRY(sc.f32(&ys, (size_t)rows * D));
    MCM_TRY(sc.f32(&ysum, (size_t)rows * D));
    // ---- y_s: static human-topology mix of body_value = motion_feat[..., :L] (:123-128), or body_value itself
    if (static_body) {
      MCM_TRY(part_mix_launch(body_weight, motion_feat, ys, (size_t)rows, H, L, st, 4 * L));
    } else {
      MCM_CUDA(cudaMemcpy2DAsync(ys, (size_t)L * 4, motion_feat, (size_t)4 * L * 4, (size_t)L * 4, (size_t)rows * H,
                                 cudaMemcpyDeviceToDevice, st));
    }
    if (dyn) {
      // ---- + dynamic body attention over the H part tokens of every frame (:130-135)
      OpPtr xn, wq;
      float* qkv = nullptr;
      MCM_TRY(sc.op(&xn, (size_t)rows * H * L, L, false));
      MCM_TRY(sc.op(&wq, (size_t)3 * L * L, L, false));
      MCM_TRY(sc.f32(&qkv, (size_t)rows * H * 3 * L));
      MCM_TRY(ln_rows_launch(motion_feat, rows * H, L, 4 * L, dyn_ln_w, dyn_ln_b, nullptr, nullptr, 4, 1, false, xn, OP_F16, st));
      MCM_TRY(pack_op_launch(dyn_wqkv, 3 * L, L, L, false, wq, OP_F16, st));
      GemmProblem g = shared_weight_problem(xn, rows * H, wq, 3 * L, L, OP_F16);
      g.seg[0] = seg_of(3 * L);
      g.seg[0].bias = dyn_bqkv; g.seg[0].out32 = qkv; g.seg[0].ld32 = 3 * L;
      MCM_TRY(gemm_tc_launch(g, st));
      constexpr int HEADS = 8;                       // st_attention.py:88-93: num_heads = 8, hard-coded
      MCM_CHECK(L % HEADS == 0, "mcm_stma_mix: the dynamic body attention has 8 heads");
      const int hd = L / HEADS;
      const size_t smem = ((size_t)3 * H * L + (size_t)HEADS * hd * hd) * sizeof(float);
      MCM_CHECK(smem <= 48 * 1024, "mcm_stma_mix: dynamic body tile too large");
      stma_dyn_body_kernel<HEADS><<<dim3((unsigned)rows), dim3(128), smem, st>>>(qkv, motion_feat, H, L, ys);
      MCM_CUDA(cudaGetLastError());
    }
    // ---- temporal branch: keys / values of the text and motion tokens, transposed and masked (:146-161)
    stma_kv_pack_kernel<<<dim3((unsigned)((Np + 31) / 32), (unsigned)(D / 32), (unsigned)B), dim3(32, 8), 0, st>>>(
        motion_feat, text_feat, src_mask, text_cond, T, H, L, Nt, Ht, Np, keyT, reinterpret_cast<uint16_t*>(valT.hi));
    MCM_CUDA(cudaGetLastError());
    MCM_TRY(softmax_seg_launch(keyT, B * D, N, Np, N, keyS, OP_F16, st));                 // softmax over the N tokens (:156)
    MCM_TRY(softmax_seg_launch(motion_feat + 3 * L, rows * H, L, 4 * L, L, qS, OP_F16, st));   // query softmax over L (:163-164)
    {  // attT[b, h][l][d] = sum_n valT[b, h*L + l, n] keyS[b, h*L + d, n]        ('bnhd,bnhl->bhdl', :167, as its transpose)
      GemmProblem g;
      std::memset(&g, 0, sizeof(g));
      g.a = valT; g.a_rows = L; g.a_k = Np; g.a_batches = B * H; g.a_batched = 1;
      g.b = keyS; g.b_rows = L; g.b_k = Np; g.b_batches = B * H; g.b_batched = 1;
      g.fmt = OP_F16; g.M = L; g.K = N; g.batches = B * H; g.inner = 1;
      g.out_rows_per_outer = L;
      g.nseg = 1;
      g.seg[0] = seg_of(L);
      g.seg[0].op = attT; g.seg[0].op_fmt = OP_F16;
      MCM_TRY(gemm_tc_launch(g, st));
    }
    {  // ysum[b, t, h, l] = y_s + sum_d q[b, t, h, d] att[b, h][d][l]              ('bnhd,bhdl->bnhl', :169-171)
      GemmProblem g;
      std::memset(&g, 0, sizeof(g));
      g.a = OpPtr{qS.hi, nullptr, D}; g.a_rows = T; g.a_k = D; g.a_batches = B; g.a_k_inner = L;
      g.b = attT; g.b_rows = L; g.b_k = L; g.b_batches = B * H; g.b_batched = 1;
      g.fmt = OP_F16; g.M = T; g.K = L; g.batches = B * H; g.inner = H;
      g.out_rows_per_outer = T; g.out_col_inner = L;
      g.nseg = 1;
      g.seg[0] = seg_of(L);
      g.seg[0].addend = ys; g.seg[0].out32 = ysum; g.seg[0].ld32 = D;
      MCM_TRY(gemm_tc_launch(g, st));
    }
    const StyleParams sp{emb_w, emb_b, ln_w, ln_b, out_w, out_b};
    return stylization_tail(sc, B, T, D, E, ysum, x, emb, sp, out, st);           // y = x + proj_out(y_s + y_t, emb)  (:172)
  };
  int rc = run();
  const cudaError_t e = cudaStreamSynchronize(st);
  if (!rc && e != cudaSuccess) {
    set_error(std::string("mcm_stma_mix: ") + cudaGetErrorString(e));
    rc = 1;
  }
  return rc;
}

}  // extern "C"
