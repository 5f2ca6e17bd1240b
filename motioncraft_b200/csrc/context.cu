// motioncraft_b200 -- the denoiser context: parameter packing, HBM workspace, the per-step schedule of
// kernels for the MCM transformer (+ ControlNet branch) and the DDIM / DDPM sampler loop; plus the
// extern "C" surface declared in include/mcm_b200.h.
//
// Data layout in HBM (B samples, T frames, D = latent 512):
//   x      [B*T, 322]  fp32   sampler state                     xop  [B*T, 328] bf16 hi/lo (joint_embed operand)
//   h      [B*T, D]    fp32   residual stream (never rounded)    hop  [B*T, D]   16-bit copy where a GEMM reads h raw
//   mod    [B, sum 2d] fp32   every block's AdaLN (scale|shift), one GEMM per step
//   16-bit GEMM operands are K-major rows with a pitch that is a multiple of 8 elements, pad columns 0.
// The channel-attention ("SA") works on h^T without ever materialising a transposed fp32 tensor: the
// transposing LayerNorm writes the operand (B*D rows x T), and GEMM epilogues write back transposed.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mcm_b200.h"
#include "elementwise.cuh"
#include "fused_block.cuh"
#include "gemm_tc.cuh"
#include "timing.cuh"

namespace mcm {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
std::string mcm_last_error_string() { return g_last_error; }

inline int rup(int x, int m) { return (x + m - 1) / m * m; }
inline size_t smax(size_t a, size_t b) { return a > b ? a : b; }

struct Block {
  float *sa_ln_w, *sa_ln_b, *sa_bqkv, *sa_pn_w, *sa_pn_b, *sa_bo;
  OpPtr sa_wqkv, sa_wo;
  float *ca_ln_w, *ca_ln_b, *ca_bq, *ca_tn_w, *ca_tn_b, *ca_bkv, *ca_pn_w, *ca_pn_b, *ca_bo;
  OpPtr ca_wq, ca_wkv, ca_wo;
  OpPtr ca_ctxT;  // [B*H, hdD, hdD] step-invariant per-head context, transposed (B operand of q*ctx)
  float *f_b1, *f_b2, *f_pn_w, *f_pn_b, *f_bo;
  OpPtr f_w1, f_w2, f_wo;
  int mod_off;    // [sa scale T | sa shift T | ca scale D | ca shift D | ffn scale D | ffn shift D]
};

struct Ctrl {
  OpPtr before_w, after_w;
  float *before_b, *after_b;
};

// Activations / operands that live only inside one pass through the layer stack.  Two sets exist so that the two
// halves of a batch can run concurrently on two streams (set 1 is sized for half the batch).
struct Scratch {
  float *h32, *f32A, *f32B, *c32;
  OpPtr opA, opB, opC, opD, hop, ctxT_sa, c_op;
  void* hid;   // hidden-activation scratch of the fused cross-attention + FFN kernel
};
}  // namespace mcm

using namespace mcm;

struct mcm_ctx {
  mcm_config cfg;
  int T, Tp, D, E, F, L, H, IN, INp, NTmax, NTp, nL, nC, Cin, Cinp, hdT, hdD, Bmax, mod_total;
  int fused = 1;          // MCM_FUSED=0: run cross-attention + FFN as separate GEMM / row kernels (the round-1 path)
  int fused_sa = 2;       // MCM_FUSED_SA: 0 = channel attention as separate kernels, 1 = fused tail, 2 = + fused head (default),
                          // 3 = + fused token softmax / context (parity-green, but 135 us vs 74 us for the two kernels it
                          //     replaces: its per-sample rounds serialise behind the row softmax; opt-in until reworked)
  int fused_min_rows = 2048;   // the persistent tile kernels need enough 256-row tiles to fill the 74 CTA pairs: below this
                               // many rows (B*T) per launch the kernel-per-op path is faster (B=1: 66 vs 75 ms per 50-step run)
  int fused_sa_min_rows = -1;  // threshold of the fused channel-attention kernels; -1 = fused_min_rows
  int fused_stop = 0;     // debug: truncate the fused kernel after this many phases and dump its operand tile
  void* fused_dbg = nullptr;
  int chunk = 0;          // samples per pass through the layer stack (0 = whole batch); MCM_CHUNK
  bool finalized = false;
  bool cond_ready = false;
  int cond_batch = 0;
  bool have_c = false;
  std::map<std::string, std::pair<const float*, long long>> params;
  std::vector<void*> allocs;

  // packed parameters
  std::vector<Block> blocks;      // nL base blocks followed by nC control copies
  std::vector<Ctrl> ctrls;
  OpPtr w_joint, w_te0, w_te2, w_mod, w_out, w_cci;
  float *b_joint, *b_te0, *b_te2, *b_mod, *b_out, *b_cci, *seq_emb;

  // workspace
  Scratch ws[2];
  float *mod32, *emb32, *xfproj32, *eps32, *x32, *cc32;
  OpPtr xop, te_op, t1_op, emb_op, cc_op;
  cudaStream_t s1 = nullptr;             // second stream: the other half of the batch
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t s0 = nullptr;             // capture / replay stream used when the caller's stream is the legacy default
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  int dual = 1;                          // MCM_DUAL=0: single stream
  int split_sms = 0;                     // MCM_SPLIT_SMS=1: the two halves' persistent kernels each take half the SMs
  // CUDA graph of one denoise step (sampler loop): captured once per (batch, control on/off), replayed every step;
  // the step's timestep is read from `t_buf`, so the graph is identical for all steps.
  int use_graph = 1;                     // MCM_GRAPH=0: eager launches
  long long* t_buf = nullptr;
  struct StepGraph { int B; bool have_c; long long key; cudaGraphExec_t exec; unsigned long long n_gemm, n_row, n_fused, last_use; };
  std::vector<StepGraph> graphs;         // at most MAX_GRAPHS entries, least recently used evicted
  static constexpr size_t MAX_GRAPHS = 8;
  unsigned long long graph_clock = 0;
  float* noise_buf = nullptr;            // [Bmax*T*IN] per-step sampler noise staged / generated on the device (lazy)
  // Hoisted AdaLN modulation (run_sampler): emb = time_embed(t) + xf_proj and mod = emb_layers(SiLU(emb)) depend on the step only
  // through t, so for small batches all steps' modulation vectors are computed by ONE batched pass before the loop
  // (3 GEMMs + 2 row kernels per run instead of per step) and each step's graph just gathers its slice.
  int hoist_mod = 1;                     // MCM_HOIST_MOD=0 / set_option("hoist_mod", 0): per-step modulation always
  float* mod_all = nullptr;              // [steps * B, mod_total]
  float* emb_all = nullptr;              // [steps * B, E]
  long long* t_all = nullptr;            // [steps * B]
  long long* step_buf = nullptr;         // device scalar: index of the current step's slice
  OpPtr te_all, t1_all, embop_all;
  size_t hoist_rows = 0;                 // capacity of the buffers above in rows
  bool hoist_active = false;             // the current sampling run uses mod_all
  // every scheduling option that changes the captured launch sequence is part of the graph key
  long long graph_key() const {
    return (long long)fused + 2ll * fused_sa + 4ll * (hoist_active ? 1 : 0) + 16ll * (dual ? 1 : 0) + 32ll * fused_stop + 256ll * (split_sms ? 1 : 0) + 512ll * (long long)chunk +
           (1ll << 24) * (long long)fused_min_rows + (1ll << 44) * (long long)(fused_sa_min_rows >= 0 ? 1 + fused_sa_min_rows / 64 : 0);
  }
  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }

  int fmt_fast() const { return cfg.precise_all ? OP_BF16X2 : OP_F16; }
  int fmt_prec() const { return OP_BF16X2; }

  ~mcm_ctx() {
    for (void* p : allocs) cudaFree(p);
    drop_graphs();
    if (s1) cudaStreamDestroy(s1);
    if (s0) cudaStreamDestroy(s0);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
};

namespace {

int dev_alloc(mcm_ctx* c, void** out, size_t bytes) {
  void* p = nullptr;
  bytes = smax(bytes, 256);
  MCM_CUDA(cudaMalloc(&p, bytes));
  MCM_CUDA(cudaMemset(p, 0, bytes));
  c->allocs.push_back(p);
  *out = p;
  return 0;
}
int alloc_f32(mcm_ctx* c, float** out, size_t n) { return dev_alloc(c, reinterpret_cast<void**>(out), n * 4); }
int alloc_op(mcm_ctx* c, OpPtr* o, size_t elems, int ld, bool with_lo) {
  o->ld = ld;
  o->lo = nullptr;
  MCM_TRY(dev_alloc(c, &o->hi, elems * 2));
  if (with_lo) MCM_TRY(dev_alloc(c, &o->lo, elems * 2));
  return 0;
}
inline OpPtr view(const OpPtr& o, int ld) { return OpPtr{o.hi, o.lo, ld}; }
inline OpPtr offs(const OpPtr& o, size_t elems) {
  return OpPtr{reinterpret_cast<uint16_t*>(o.hi) + elems, o.lo ? reinterpret_cast<uint16_t*>(o.lo) + elems : nullptr, o.ld};
}

int get_param(mcm_ctx* c, const std::string& name, long long numel, const float** out) {
  auto it = c->params.find(name);
  if (it == c->params.end()) {
    set_error("missing parameter: " + name);
    return 1;
  }
  if (it->second.second != numel) {
    set_error("parameter " + name + " has " + std::to_string(it->second.second) + " elements, expected " + std::to_string(numel));
    return 1;
  }
  *out = it->second.first;
  return 0;
}

// copy an fp32 parameter (or a concatenation of parameters) into context-owned memory
int own_f32(mcm_ctx* c, const std::vector<std::pair<std::string, long long>>& pieces, float** out, cudaStream_t st) {
  long long total = 0;
  for (auto& p : pieces) total += p.second;
  if (*out == nullptr) MCM_TRY(alloc_f32(c, out, (size_t)total));      // a second mcm_finalize_params re-packs in place
  long long off = 0;
  for (auto& p : pieces) {
    const float* src;
    MCM_TRY(get_param(c, p.first, p.second, &src));
    MCM_CUDA(cudaMemcpyAsync(*out + off, src, (size_t)p.second * 4, cudaMemcpyDeviceToDevice, st));
    off += p.second;
  }
  return 0;
}

// pack (a row-concatenation of) Linear weights [rows_i, in] into one K-major operand [sum rows, in_p]
int pack_weight(mcm_ctx* c, const std::vector<std::pair<std::string, int>>& pieces, int in, int fmt, OpPtr* out,
                cudaStream_t st) {
  int rows = 0;
  for (auto& p : pieces) rows += p.second;
  const int in_p = rup(in, 8);
  if (out->hi == nullptr) MCM_TRY(alloc_op(c, out, (size_t)rows * in_p, in_p, fmt == OP_BF16X2));
  int r0 = 0;
  for (auto& p : pieces) {
    const float* src;
    MCM_TRY(get_param(c, p.first, (long long)p.second * in, &src));
    OpPtr dst{reinterpret_cast<uint16_t*>(out->hi) + (size_t)r0 * in_p,
              out->lo ? reinterpret_cast<uint16_t*>(out->lo) + (size_t)r0 * in_p : nullptr, in_p};
    MCM_TRY(pack_op_launch(src, p.second, in, in, false, dst, fmt, st));
    r0 += p.second;
  }
  return 0;
}

EpiSeg seg_default(int n, int w_row0 = 0) {
  EpiSeg s;
  std::memset(&s, 0, sizeof(s));
  s.n = n;
  s.w_row0 = w_row0;
  return s;
}

// shared-weight Linear over `rows` rows: out = act(A W^T + bias + addend)
GemmProblem linear_problem(const OpPtr& a, int rows, const OpPtr& w, int w_rows, int K, int fmt) {
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

int build_block(mcm_ctx* c, const std::string& pfx, Block* b, int mod_off, cudaStream_t st) {
  const int T = c->T, D = c->D, F = c->F, L = c->L;
  const int ff = c->fmt_fast();
  const std::string sa = pfx + ".sa_block.", ca = pfx + ".ca_block.", fn = pfx + ".ffn_temporal.";
  MCM_TRY(own_f32(c, {{sa + "norm.weight", T}}, &b->sa_ln_w, st));
  MCM_TRY(own_f32(c, {{sa + "norm.bias", T}}, &b->sa_ln_b, st));
  MCM_TRY(own_f32(c, {{sa + "query.bias", T}, {sa + "key.bias", T}, {sa + "value.bias", T}}, &b->sa_bqkv, st));
  MCM_TRY(pack_weight(c, {{sa + "query.weight", T}, {sa + "key.weight", T}, {sa + "value.weight", T}}, T, ff, &b->sa_wqkv, st));
  MCM_TRY(own_f32(c, {{sa + "proj_out.norm.weight", T}}, &b->sa_pn_w, st));
  MCM_TRY(own_f32(c, {{sa + "proj_out.norm.bias", T}}, &b->sa_pn_b, st));
  MCM_TRY(pack_weight(c, {{sa + "proj_out.out_layers.2.weight", T}}, T, ff, &b->sa_wo, st));
  MCM_TRY(own_f32(c, {{sa + "proj_out.out_layers.2.bias", T}}, &b->sa_bo, st));

  MCM_TRY(own_f32(c, {{ca + "norm.weight", D}}, &b->ca_ln_w, st));
  MCM_TRY(own_f32(c, {{ca + "norm.bias", D}}, &b->ca_ln_b, st));
  MCM_TRY(own_f32(c, {{ca + "text_norm.weight", L}}, &b->ca_tn_w, st));
  MCM_TRY(own_f32(c, {{ca + "text_norm.bias", L}}, &b->ca_tn_b, st));
  MCM_TRY(pack_weight(c, {{ca + "query.weight", D}}, D, ff, &b->ca_wq, st));
  MCM_TRY(own_f32(c, {{ca + "query.bias", D}}, &b->ca_bq, st));
  MCM_TRY(pack_weight(c, {{ca + "key.weight", D}, {ca + "value.weight", D}}, L, ff, &b->ca_wkv, st));
  MCM_TRY(own_f32(c, {{ca + "key.bias", D}, {ca + "value.bias", D}}, &b->ca_bkv, st));
  MCM_TRY(own_f32(c, {{ca + "proj_out.norm.weight", D}}, &b->ca_pn_w, st));
  MCM_TRY(own_f32(c, {{ca + "proj_out.norm.bias", D}}, &b->ca_pn_b, st));
  MCM_TRY(pack_weight(c, {{ca + "proj_out.out_layers.2.weight", D}}, D, ff, &b->ca_wo, st));
  MCM_TRY(own_f32(c, {{ca + "proj_out.out_layers.2.bias", D}}, &b->ca_bo, st));
  if (b->ca_ctxT.hi == nullptr)
    MCM_TRY(alloc_op(c, &b->ca_ctxT, (size_t)c->Bmax * c->H * c->hdD * c->hdD, c->hdD, ff == OP_BF16X2));

  MCM_TRY(pack_weight(c, {{fn + "linear1.weight", F}}, D, ff, &b->f_w1, st));
  MCM_TRY(own_f32(c, {{fn + "linear1.bias", F}}, &b->f_b1, st));
  MCM_TRY(pack_weight(c, {{fn + "linear2.weight", D}}, F, ff, &b->f_w2, st));
  MCM_TRY(own_f32(c, {{fn + "linear2.bias", D}}, &b->f_b2, st));
  MCM_TRY(own_f32(c, {{fn + "proj_out.norm.weight", D}}, &b->f_pn_w, st));
  MCM_TRY(own_f32(c, {{fn + "proj_out.norm.bias", D}}, &b->f_pn_b, st));
  MCM_TRY(pack_weight(c, {{fn + "proj_out.out_layers.2.weight", D}}, D, ff, &b->f_wo, st));
  MCM_TRY(own_f32(c, {{fn + "proj_out.out_layers.2.bias", D}}, &b->f_bo, st));
  b->mod_off = mod_off;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// one DecoderLayer (mcm.py:25-41) on the fp32 residual stream `h` [B*T, D], in place.
// `mod` points at this block's modulation vectors (pitch mod_ld).  If hop_out_fmt >= 0 the block's
// final GEMM also emits the 16-bit copy of the new h into c->ws[0].hop in that format.
// ---------------------------------------------------------------------------------------------
int run_block(mcm_ctx* c, Scratch& w, const Block& k, int B, float* h, const float* mod, int mod_ld, OpPtr final_op,
              int final_op_fmt, cudaStream_t st, int b0 = 0) {
  const int T = c->T, Tp = c->Tp, D = c->D, F = c->F, H = c->H, hdT = c->hdT, hdD = c->hdD;
  const int ff = c->fmt_fast();
  const OpPtr opA_t = view(w.opA, Tp), opC_t = view(w.opC, Tp);
  const OpPtr opA_d = view(w.opA, D), opB_d = view(w.opB, D), opC_d = view(w.opC, D), opD_d = view(w.opD, D);
  const OpPtr opB_f = view(w.opB, F);
  const OpPtr hop = view(w.hop, D);

  // ---- channel attention (EfficientSelfAttention on x^T, efficient_attention.py:25-46) ----
  const bool big = (long long)B * T >= c->fused_min_rows || c->fused_stop != 0;
  // (the channel-attention kernels may get their own threshold: under ncu they win at every batch size -- B = 1: head 20 us +
  // tail 13 us against 23 + 28 us for the six launches they replace -- but in the replayed graph the gain is within noise at
  // B = 1 and 3 % at B = 8, so by default they follow fused_min_rows and a launch takes one schedule or the other as a whole)
  const bool big_sa = (long long)B * T >= (c->fused_sa_min_rows >= 0 ? c->fused_sa_min_rows : c->fused_min_rows) || c->fused_stop != 0;
  const bool sa_fused = big_sa && c->fused && c->fused_sa && ff == OP_F16 && sa_tail_supported(T, D) && H == 4 && 32 * Tp * 2 <= 16384;
  if (sa_fused && c->fused_sa >= 2) {
    // ---- LN_T(h^T) -> q | k | v -> softmax(q): one persistent kernel (fused_block.cu); k (fp32) and v go back to [B, T', D]
    SaFrontArgs a;
    std::memset(&a, 0, sizeof(a));
    a.h = h; a.T = T; a.batch = B; a.heads = H;
    a.ln_w = k.sa_ln_w; a.ln_b = k.sa_ln_b; a.w = k.sa_wqkv; a.bqkv = k.sa_bqkv;
    a.qs = opC_t; a.k32 = w.f32B; a.v16 = opB_d;
    MCM_TRY(sa_front_launch(a, st));
  } else {
  // xn^T = LayerNorm_T(h^T)                                   -> opA [B*D, Tp]
  MCM_TRY(ln_transpose_launch(h, B, T, D, k.sa_ln_w, k.sa_ln_b, opA_t, ff, st));
  {  // q | k | v = xn^T W^T + b ; q stays row-major fp32, k and v go back to the [B, T', D] layout
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opA_t; g.a_rows = D; g.a_k = Tp; g.a_batches = B;
    g.b = k.sa_wqkv; g.b_rows = 3 * T; g.b_k = Tp; g.b_batches = 1;
    g.fmt = ff; g.M = D; g.K = T; g.batches = B; g.inner = 1;
    g.out_rows_per_outer = D; g.trans_rows = T;
    g.nseg = 3;
    g.seg[0] = seg_default(T, 0);
    g.seg[0].bias = k.sa_bqkv; g.seg[0].out32 = w.f32A; g.seg[0].ld32 = T;
    g.seg[1] = seg_default(T, T);
    g.seg[1].bias = k.sa_bqkv + T; g.seg[1].out32 = w.f32B; g.seg[1].ld32 = D; g.seg[1].flags = EPI_TRANSPOSED;
    g.seg[2] = seg_default(T, 2 * T);
    g.seg[2].bias = k.sa_bqkv + 2 * T; g.seg[2].op = opB_d; g.seg[2].op_fmt = ff; g.seg[2].flags = EPI_TRANSPOSED;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  // q: softmax over each head's T/H features; k: softmax over the D channel-tokens (a row in [B,T',D])
  MCM_TRY(softmax_seg_launch(w.f32A, B * D, T, T, hdT, opC_t, ff, st));
  }   // unfused channel-attention head
  if (sa_fused && c->fused_sa >= 3) {
    // ---- token softmax of k + per-head context k^T v: one persistent kernel (fused_block.cu)
    SaCtxArgs a;
    std::memset(&a, 0, sizeof(a));
    a.k32 = w.f32B; a.v16 = opB_d; a.ctxT = view(w.ctxT_sa, Tp); a.T = T; a.batch = B; a.heads = H;
    MCM_TRY(sa_ctx_launch(a, st));
  } else {
  MCM_TRY(softmax_seg_launch(w.f32B, B * T, D, D, D, opD_d, ff, st));
  // Wide heads (T / H a multiple of 64, >= 128: m2d's T = 1024): only the H diagonal (T/H x T/H) blocks of ctx are non-zero,
  // so both contractions run per (sample, head) -- a quarter of the tensor-core work of the full T x T product.  The
  // off-diagonal part of ctxT_sa is zero from its allocation on (every writer of that buffer stores zeros there).
  static const int perhead_env = [] { const char* e = getenv("MCM_SA_PERHEAD"); return e ? atoi(e) : 1; }();
  const bool per_head = perhead_env && hdT % 64 == 0 && hdT >= 128;
  if (per_head) {
    // ctxT[b][h*hd + l][h*hd + dk] = sum_n v[b, h*hd + l, n] ks[b, h*hd + dk, n]
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opB_d; g.a_rows = hdT; g.a_k = D; g.a_batches = B * H; g.a_batched = 1;
    g.b = opD_d; g.b_rows = hdT; g.b_k = D; g.b_batches = B * H; g.b_batched = 1;
    g.fmt = ff; g.M = hdT; g.K = D; g.batches = B * H; g.inner = H;
    g.out_rows_per_outer = hdT; g.out_batched = 1; g.out_col_inner = hdT;
    g.nseg = 1;
    g.seg[0] = seg_default(hdT, 0);
    g.seg[0].op = view(w.ctxT_sa, Tp); g.seg[0].op_fmt = ff;
    g.algo_flops = 2.0 * T * hdT * D * B;
    MCM_TRY(gemm_tc_launch(g, st));
  } else
  {  // ctx[b] = softmax(k)^T v, kept block-diagonal per head, needed transposed (ctxT[b][l][dk], the B operand of q ctx).
     // Computed AS the transpose -- A = v (rows l), B = softmax(k) (rows dk): D[l][dk] = sum_n v[l, n] ks[dk, n] -- so the
     // epilogue takes the plain TMA-store path instead of the transposing one (46 -> us per launch at B = 256).
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opB_d; g.a_rows = T; g.a_k = D; g.a_batches = B;
    g.b = opD_d; g.b_rows = T; g.b_k = D; g.b_batches = B; g.b_batched = 1;
    g.fmt = ff; g.M = T; g.K = D; g.batches = B; g.inner = 1;
    g.out_rows_per_outer = T; g.head_dim = hdT;
    g.nseg = 1;
    g.seg[0] = seg_default(T, 0);
    g.seg[0].op = view(w.ctxT_sa, Tp); g.seg[0].op_fmt = ff;
    g.seg[0].flags = EPI_MASK_BLOCKDIAG;
    g.algo_flops = 2.0 * T * hdT * D * B;      // only the per-head diagonal blocks are algorithmic work
    MCM_TRY(gemm_tc_launch(g, st));
  }
  }   // unfused token softmax + context
  if (sa_fused) {
    // ---- y = softmax(q) ctx -> AdaLN_T -> SiLU -> Linear(T,T) -> h^T += : one persistent kernel (fused_block.cu)
    SaTailArgs a;
    std::memset(&a, 0, sizeof(a));
    a.h = h; a.T = T; a.batch = B;
    a.qs = opC_t; a.ctxT = view(w.ctxT_sa, Tp); a.wo = k.sa_wo;
    a.pn_w = k.sa_pn_w; a.pn_b = k.sa_pn_b; a.scale = mod + k.mod_off; a.shift = mod + k.mod_off + T; a.bo = k.sa_bo;
    a.mod_ld = mod_ld;
    MCM_TRY(sa_tail_launch(a, st));
  } else {
  static const int perhead_env2 = [] { const char* e = getenv("MCM_SA_PERHEAD"); return e ? atoi(e) : 1; }();
  if (perhead_env2 && hdT % 64 == 0 && hdT >= 128) {
    // y^T[b][:, head] = softmax(q)[b][:, head] ctx[b][head]: K runs over the head's T/H features only
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opC_t; g.a_rows = D; g.a_k = Tp; g.a_batches = B; g.a_k_inner = hdT;
    g.b = view(w.ctxT_sa, Tp); g.b_rows = hdT; g.b_k = Tp; g.b_batches = B * H; g.b_batched = 1; g.b_k_inner = hdT;
    g.fmt = ff; g.M = D; g.K = hdT; g.batches = B * H; g.inner = H;
    g.out_rows_per_outer = D; g.out_col_inner = hdT;
    g.nseg = 1;
    g.seg[0] = seg_default(hdT, 0);
    g.seg[0].out32 = w.f32A; g.seg[0].ld32 = T;
    g.algo_flops = 2.0 * D * T * hdT * B;
    MCM_TRY(gemm_tc_launch(g, st));
  } else
  {  // y^T[b] = softmax(q) ctx                                 -> f32A [B*D, T]
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opC_t; g.a_rows = D; g.a_k = Tp; g.a_batches = B;
    g.b = view(w.ctxT_sa, Tp); g.b_rows = T; g.b_k = Tp; g.b_batches = B; g.b_batched = 1;
    g.fmt = ff; g.M = D; g.K = T; g.batches = B; g.inner = 1;
    g.out_rows_per_outer = D;
    g.nseg = 1;
    g.seg[0] = seg_default(T, 0);
    g.seg[0].out32 = w.f32A; g.seg[0].ld32 = T;
    g.algo_flops = 2.0 * D * T * hdT * B;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  // StylizationBlock over T: SiLU(LN(y) (1 + scale) + shift)    -> opA [B*D, Tp]
  MCM_TRY(ln_rows_launch(w.f32A, B * D, T, T, k.sa_pn_w, k.sa_pn_b, mod + k.mod_off, mod + k.mod_off + T, mod_ld, D,
                         true, opA_t, ff, st));
  {  // h[b, t', d] += (. W_o^T + b_o)[(b,d), t']
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opA_t; g.a_rows = D; g.a_k = Tp; g.a_batches = B;
    g.b = k.sa_wo; g.b_rows = T; g.b_k = Tp; g.b_batches = 1;
    g.fmt = ff; g.M = D; g.K = T; g.batches = B; g.inner = 1;
    g.out_rows_per_outer = D; g.trans_rows = T;
    g.nseg = 1;
    g.seg[0] = seg_default(T, 0);
    g.seg[0].bias = k.sa_bo; g.seg[0].addend = h; g.seg[0].out32 = h; g.seg[0].ld32 = D;
    g.seg[0].flags = EPI_TRANSPOSED;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  }   // unfused channel-attention tail

  if (big && c->fused && ff == OP_F16 && w.hid != nullptr && fused_block_supported(T, D, F, H)) {
    // ---- cross attention + FFN in ONE persistent kernel (fused_block.cu) ----
    FusedBlockArgs a;
    std::memset(&a, 0, sizeof(a));
    a.h = h; a.rows = B * T; a.T = T; a.batch = B;
    a.ca_ln_w = k.ca_ln_w; a.ca_ln_b = k.ca_ln_b; a.ca_wq = k.ca_wq; a.ca_bq = k.ca_bq;
    a.ca_ctxT = offs(k.ca_ctxT, (size_t)b0 * H * hdD * hdD);
    a.ca_pn_w = k.ca_pn_w; a.ca_pn_b = k.ca_pn_b;
    a.ca_scale = mod + k.mod_off + 2 * T; a.ca_shift = mod + k.mod_off + 2 * T + D;
    a.ca_wo = k.ca_wo; a.ca_bo = k.ca_bo;
    a.f_w1 = k.f_w1; a.f_b1 = k.f_b1; a.f_w2 = k.f_w2; a.f_b2 = k.f_b2;
    a.f_pn_w = k.f_pn_w; a.f_pn_b = k.f_pn_b;
    a.f_scale = mod + k.mod_off + 2 * T + 2 * D; a.f_shift = mod + k.mod_off + 2 * T + 3 * D;
    a.f_wo = k.f_wo; a.f_bo = k.f_bo;
    a.mod_ld = mod_ld; a.hid = w.hid; a.stop = c->fused_stop; a.dbg = c->fused_dbg;
    MCM_TRY(fused_block_launch(a, st));
    if (final_op.hi) MCM_TRY(pack_op_launch(h, B * T, D, D, false, final_op, final_op_fmt, st));
    return 0;
  }

  // ---- text cross attention (EfficientCrossAttention, efficient_attention.py:64-92) ----
  MCM_TRY(ln_rows_launch(h, B * T, D, D, k.ca_ln_w, k.ca_ln_b, nullptr, nullptr, 4, T, false, opA_d, ff, st));
  {
    GemmProblem g = linear_problem(opA_d, B * T, k.ca_wq, D, D, ff);
    g.seg[0] = seg_default(D, 0);
    g.seg[0].bias = k.ca_bq; g.seg[0].out32 = w.f32A; g.seg[0].ld32 = D;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  MCM_TRY(softmax_seg_launch(w.f32A, B * T, D, D, hdD, opC_d, ff, st));
  {  // y[b, :, head] = softmax(q)[b, :, head] ctx[b, head]     (context precomputed per run)
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = opC_d; g.a_rows = T; g.a_k = D; g.a_batches = B;
    g.b = offs(k.ca_ctxT, (size_t)b0 * H * hdD * hdD); g.b_rows = hdD; g.b_k = hdD; g.b_batches = B * H; g.b_batched = 1;
    g.fmt = ff; g.M = T; g.K = hdD; g.batches = B * H; g.inner = H; g.a_k_inner = hdD;
    g.out_col_inner = hdD; g.out_rows_per_outer = T;
    g.nseg = 1;
    g.seg[0] = seg_default(hdD, 0);
    g.seg[0].out32 = w.f32A; g.seg[0].ld32 = D;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  MCM_TRY(ln_rows_launch(w.f32A, B * T, D, D, k.ca_pn_w, k.ca_pn_b, mod + k.mod_off + 2 * T,
                         mod + k.mod_off + 2 * T + D, mod_ld, T, true, opA_d, ff, st));
  {  // h += . W_o^T + b_o ; also emit the 16-bit copy of h that linear1 reads
    GemmProblem g = linear_problem(opA_d, B * T, k.ca_wo, D, D, ff);
    g.seg[0] = seg_default(D, 0);
    g.seg[0].bias = k.ca_bo; g.seg[0].addend = h; g.seg[0].out32 = h; g.seg[0].ld32 = D;
    g.seg[0].op = hop; g.seg[0].op_fmt = ff;
    MCM_TRY(gemm_tc_launch(g, st));
  }

  // ---- FFN (diffusion_transformer.py:25-28) ----
  {
    GemmProblem g = linear_problem(hop, B * T, k.f_w1, F, D, ff);
    g.seg[0] = seg_default(F, 0);
    g.seg[0].bias = k.f_b1; g.seg[0].flags = EPI_GELU; g.seg[0].op = opB_f; g.seg[0].op_fmt = ff;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  {
    GemmProblem g = linear_problem(opB_f, B * T, k.f_w2, D, F, ff);
    g.seg[0] = seg_default(D, 0);
    g.seg[0].bias = k.f_b2; g.seg[0].out32 = w.f32A; g.seg[0].ld32 = D;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  MCM_TRY(ln_rows_launch(w.f32A, B * T, D, D, k.f_pn_w, k.f_pn_b, mod + k.mod_off + 2 * T + 2 * D,
                         mod + k.mod_off + 2 * T + 3 * D, mod_ld, T, true, opA_d, ff, st));
  {
    GemmProblem g = linear_problem(opA_d, B * T, k.f_wo, D, D, ff);
    g.seg[0] = seg_default(D, 0);
    g.seg[0].bias = k.f_bo; g.seg[0].addend = h; g.seg[0].out32 = h; g.seg[0].ld32 = D;
    if (final_op.hi) { g.seg[0].op = final_op; g.seg[0].op_fmt = final_op_fmt; }
    MCM_TRY(gemm_tc_launch(g, st));
  }
  return 0;
}

// AdaLN modulation vectors of blocks [blk0, blk0 + nblk) from emb [B, E] (fp32): one GEMM
//   mod[b, :] = SiLU(emb[b]) W_mod^T + b_mod          (StylizationBlock.emb_layers, stylization_block.py:17-20,35)
int run_mod(mcm_ctx* c, int B, const float* emb, int blk0, int nblk, cudaStream_t st) {
  const int per = 2 * c->T + 4 * c->D;
  MCM_TRY(pack_op_launch(emb, B, c->E, c->E, true, c->emb_op, c->fmt_prec(), st));
  GemmProblem g = linear_problem(c->emb_op, B, c->w_mod, c->mod_total, c->E, c->fmt_prec());
  g.seg[0] = seg_default(nblk * per, blk0 * per);
  g.seg[0].bias = c->b_mod + (size_t)blk0 * per;
  g.seg[0].out32 = c->mod32; g.seg[0].ld32 = c->mod_total; g.seg[0].col0 = blk0 * per;
  return gemm_tc_launch(g, st);
}

int check_batch(mcm_ctx* c, int B) {
  MCM_CHECK(c != nullptr, "null context");
  MCM_CHECK(c->finalized, "mcm_finalize_params has not been called");
  MCM_CHECK(B >= 1 && B <= c->Bmax, "batch exceeds max_batch of the context");
  return 0;
}

int run_stack(mcm_ctx* c, Scratch& w, int b0, int B, float* eps_out, cudaStream_t st);

// joint_embed -> decoder layers (+ control branch) -> out, for samples [b0, b0 + B) of the current step
int run_layers(mcm_ctx* c, Scratch& w, int b0, int B, float* eps_out, cudaStream_t st) {
  const int T = c->T, D = c->D, IN = c->IN;
  const int fp = c->fmt_prec();
  const size_t r0 = (size_t)b0 * T;           // first row of this chunk in per-sample tensors
  {  // h = joint_embed(x) + sequence_embedding[:T]               (diffusion_transformer.py:215-218)
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = offs(c->xop, r0 * c->INp); g.a_rows = T; g.a_k = c->INp; g.a_batches = B;
    g.b = c->w_joint; g.b_rows = D; g.b_k = c->INp; g.b_batches = 1;
    g.fmt = fp; g.M = T; g.K = IN; g.batches = B; g.inner = 1;
    g.out_rows_per_outer = T;
    g.nseg = 1;
    g.seg[0] = seg_default(D, 0);
    g.seg[0].bias = c->b_joint; g.seg[0].addend = c->seq_emb; g.seg[0].flags = EPI_ADDEND_BCAST;
    g.seg[0].out32 = w.h32; g.seg[0].ld32 = D;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  return run_stack(c, w, b0, B, eps_out, st);
}

// decoder layers (+ control branch) -> out on the residual stream already in w.h32: MCMTransformer.forward_test
// (mcm.py:93-102) / ControlT2MHalf_MCM.forward_test (controlnet_mcm.py:306-361)
int run_stack(mcm_ctx* c, Scratch& w, int b0, int B, float* eps_out, cudaStream_t st) {
  const int T = c->T, D = c->D, IN = c->IN;
  const int fp = c->fmt_prec();
  const size_t r0 = (size_t)b0 * T;
  const OpPtr none{nullptr, nullptr, 0};
  const OpPtr hop_out = view(w.hop, D);
  const int nL = c->nL, nC = (c->have_c ? c->nC : 0);
  const float* mod = c->mod32 + (size_t)b0 * c->mod_total;
  const int mld = c->mod_total;
  // MCMTransformer.forward_test (mcm.py:93-102) / ControlT2MHalf_MCM.forward_test (controlnet_mcm.py:306-361)
  MCM_TRY(run_block(c, w, c->blocks[0], B, w.h32, mod, mld, (nL == 1) ? hop_out : none, fp, st, b0));
  for (int i = 1; i < nL; ++i) {
    if (i <= nC) {
      const int j = i - 1;
      const Block& cb = c->blocks[nL + j];
      const Ctrl& ct = c->ctrls[j];
      if (j == 0) {  // c = copied_block(x = h + before_proj(c))      (controlnet_mcm.py:65-75)
        GemmProblem g = linear_problem(offs(view(c->cc_op, D), r0 * D), B * T, ct.before_w, D, D, c->fmt_fast());
        g.seg[0] = seg_default(D, 0);
        g.seg[0].bias = ct.before_b; g.seg[0].addend = w.h32; g.seg[0].out32 = w.c32; g.seg[0].ld32 = D;
        MCM_TRY(gemm_tc_launch(g, st));
      }
      MCM_TRY(run_block(c, w, cb, B, w.c32, mod, mld, view(w.c_op, D), c->fmt_fast(), st, b0));
      {  // h = h + after_proj(c)                                       (:75,85 ; :341-349)
        GemmProblem g = linear_problem(view(w.c_op, D), B * T, ct.after_w, D, D, c->fmt_fast());
        g.seg[0] = seg_default(D, 0);
        g.seg[0].bias = ct.after_b; g.seg[0].addend = w.h32; g.seg[0].out32 = w.h32; g.seg[0].ld32 = D;
        MCM_TRY(gemm_tc_launch(g, st));
      }
    }
    MCM_TRY(run_block(c, w, c->blocks[i], B, w.h32, mod, mld, (i == nL - 1) ? hop_out : none, fp, st, b0));
  }
  {  // eps = out(h)                                                (mcm.py:102)
    GemmProblem g = linear_problem(hop_out, B * T, c->w_out, IN, D, fp);
    g.seg[0] = seg_default(IN, 0);
    g.seg[0].bias = c->b_out; g.seg[0].out32 = eps_out + r0 * IN; g.seg[0].ld32 = IN;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  return 0;
}

// the whole denoiser on x32/xop already in place: writes eps32
int run_denoiser(mcm_ctx* c, int B, const long long* t_dev, int t_uniform, float* eps_out, cudaStream_t st) {
  const int D = c->D, E = c->E;
  const int fp = c->fmt_prec();
  MCM_CHECK(c->cond_ready && c->cond_batch >= B, "mcm_prepare_conditions must be called first (for at least this batch)");
  if (c->hoist_active) {
    // this step's modulation vectors were computed before the loop (precompute_mod_all): pick the slice
    MCM_TRY(gather_slice_launch(c->mod_all, c->step_buf, c->mod32, (size_t)B * c->mod_total, st));
  } else {
  // emb = time_embed(sinusoid(t)) + xf_proj                      (diffusion_transformer.py:206-213)
  MCM_TRY(timestep_embedding_launch(t_dev, t_uniform, B, D, c->te_op, fp, st));
  {
    GemmProblem g = linear_problem(c->te_op, B, c->w_te0, E, D, fp);
    g.seg[0] = seg_default(E, 0);
    g.seg[0].bias = c->b_te0; g.seg[0].flags = EPI_SILU; g.seg[0].op = c->t1_op; g.seg[0].op_fmt = fp;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  {
    GemmProblem g = linear_problem(c->t1_op, B, c->w_te2, E, E, fp);
    g.seg[0] = seg_default(E, 0);
    g.seg[0].bias = c->b_te2; g.seg[0].addend = c->xfproj32; g.seg[0].out32 = c->emb32; g.seg[0].ld32 = E;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  MCM_TRY(run_mod(c, B, c->emb32, 0, (int)c->blocks.size(), st));
  }   // per-step modulation
  // Layer stack, in chunks of `chunk` samples: every op is per-sample, so a chunk runs the whole stack on
  // chunk-local activations (h, operands and fp32 scratch reuse the SAME addresses for every chunk) which then stay
  // resident in the 126 MB L2 instead of bouncing through HBM between kernels.
  if (c->dual && c->chunk <= 0 && B >= 2) {
    // Two halves of the batch on two streams.  The tensor-core GEMMs (one persistent CTA per SM, ~25 % of the issue
    // slots) and the issue-bound row kernels stress complementary resources, and the halves are independent, so half
    // A's GEMM co-resides with half B's row kernels on the same SMs.  Fork after the shared prologue, join before the
    // caller's next kernel; everything stays ordered with respect to the caller's stream.
    int B0 = (B + 1) / 2;
    {
      // Wave quantisation of the persistent tile kernels: a launch of `tiles` 256-row tiles on `pairs` CTA pairs takes
      // ceil(tiles / pairs) tile times however empty the last wave is (s2g: 128 x 300 rows = 150 tiles on 74 pairs = three
      // tile times for 2.03 waves of work, in halves 75 = 74 + 1 each).  When the last wave would be nearly empty, the
      // batch is split unevenly instead: the first part fills whole waves, the few peeled samples (fewer than
      // fused_min_rows rows) take the kernel-per-op schedule on the second stream meanwhile.
      const int pairs = fused_block_max_pairs();
      const long long tiles = ((long long)B * c->T + 255) / 256;
      if (pairs > 0 && c->fused && (long long)B * c->T >= c->fused_min_rows && c->ws[0].hid != nullptr && c->fused_stop == 0) {
        const long long full = tiles / pairs, rem = tiles % pairs;
        if (full >= 1 && rem > 0 && rem * 8 <= pairs) {
          const int b_fit = (int)(full * pairs * 256 / c->T);          // samples whose rows fit `full` whole waves
          const int peel = B - b_fit;
          if (b_fit >= 1 && peel >= 1 && peel <= c->Bmax / 2 && (long long)peel * c->T < c->fused_min_rows) B0 = b_fit;
        }
      }
    }
    MCM_CUDA(cudaEventRecord(c->ev_fork, st));
    MCM_CUDA(cudaStreamWaitEvent(c->s1, c->ev_fork, 0));
    // split_sms: every persistent kernel of the two halves is launched on HALF the SMs, so that the halves really run side
    // by side (a 74-pair launch of one half otherwise occupies every SM and the other half queues behind it) and are in
    // different phases -- all CTAs of one launch hit L2 / HBM in lock-step, two launches side by side smooth that demand
    tc_set_sm_share(c->split_sms ? 2 : 1);
    int rc = run_layers(c, c->ws[0], 0, B0, eps_out, st);
    if (rc == 0) rc = run_layers(c, c->ws[1], B0, B - B0, eps_out, c->s1);
    tc_set_sm_share(1);
    MCM_TRY(rc);
    MCM_CUDA(cudaEventRecord(c->ev_join, c->s1));
    MCM_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
    return 0;
  }
  const int chunk = c->chunk > 0 ? std::min(c->chunk, B) : B;
  for (int b0 = 0; b0 < B; b0 += chunk) MCM_TRY(run_layers(c, c->ws[0], b0, std::min(chunk, B - b0), eps_out, st));
  return 0;
}

// One denoise step of the sampler loop (x already packed in c->xop, eps -> c->eps32) through a cached CUDA graph.
int run_denoiser_step(mcm_ctx* c, int B, int t, cudaStream_t st) {
  if (!c->use_graph || timing_enabled()) return run_denoiser(c, B, nullptr, t, c->eps32, st);
  // the legacy default stream cannot be captured: hop onto a private stream, ordered by events on both sides
  const bool hop_stream = (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread);
  cudaStream_t gs = hop_stream ? c->s0 : st;
  if (hop_stream) {
    MCM_CUDA(cudaEventRecord(c->ev_in, st));
    MCM_CUDA(cudaStreamWaitEvent(gs, c->ev_in, 0));
  }
  MCM_TRY(fill_timesteps_launch(c->t_buf, (long long)t, B, gs));
  cudaGraphExec_t exec = nullptr;
  const long long key = c->graph_key();
  for (auto& g : c->graphs)
    if (g.B == B && g.have_c == c->have_c && g.key == key) {
      exec = g.exec;
      g.last_use = ++c->graph_clock;
      gemm_tc_count_replayed(g.n_gemm);          // keep the library's launch counters truthful under graph replay
      elementwise_count_replayed(g.n_row);
      fused_block_count_replayed(g.n_fused);
    }
  if (exec == nullptr) {
    // capture (the launches are recorded, not executed), instantiate, remember
    cudaGraph_t graph = nullptr;
    const unsigned long long g0 = gemm_tc_launch_count(), r0 = elementwise_launch_count(), f0 = fused_block_launch_count();
    MCM_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    const int rc = run_denoiser(c, B, c->t_buf, 0, c->eps32, gs);
    const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
    if (rc != 0) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    MCM_CUDA(ce);
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    MCM_CUDA(ie);
    if (c->graphs.size() >= mcm_ctx::MAX_GRAPHS) {          // evict the least recently used instantiation
      size_t victim = 0;
      for (size_t i = 1; i < c->graphs.size(); ++i)
        if (c->graphs[i].last_use < c->graphs[victim].last_use) victim = i;
      cudaGraphExecDestroy(c->graphs[victim].exec);
      c->graphs.erase(c->graphs.begin() + (long)victim);
    }
    c->graphs.push_back({B, c->have_c, key, exec, gemm_tc_launch_count() - g0, elementwise_launch_count() - r0,
                         fused_block_launch_count() - f0, ++c->graph_clock});
  }
  MCM_CUDA(cudaGraphLaunch(exec, gs));
  if (hop_stream) {
    MCM_CUDA(cudaEventRecord(c->ev_out, gs));
    MCM_CUDA(cudaStreamWaitEvent(st, c->ev_out, 0));
  }
  return 0;
}

// Where the per-step noise of a stochastic sampler comes from: an explicit device tensor [n_steps, B, T, F] (oracle /
// parity tests), an explicit host tensor of the same shape (copied one step at a time into the noise buffer), or -- the
// default of the front end -- generated on the device per step, Philox keyed by (sampler seed, step index).  The last two
// need ONE step's worth of device memory instead of n_steps (DDPM-1000 at B=256, T=196 would be 64.6 GB).
struct NoiseSource {
  const float* dev = nullptr;
  const float* host = nullptr;
  bool generate = false;
  unsigned long long seed = 0;
};

int ensure_noise_buf(mcm_ctx* c) {
  if (c->noise_buf == nullptr) MCM_TRY(alloc_f32(c, &c->noise_buf, (size_t)c->Bmax * c->T * c->IN));
  return 0;
}

int step_noise_ptr(mcm_ctx* c, const NoiseSource& ns, long long index, size_t n, cudaStream_t st, const float** out) {
  *out = nullptr;
  if (ns.dev) {
    *out = ns.dev + (size_t)index * n;
  } else if (ns.host) {
    MCM_CUDA(cudaMemcpyAsync(c->noise_buf, ns.host + (size_t)index * n, n * 4, cudaMemcpyHostToDevice, st));
    *out = c->noise_buf;
  } else if (ns.generate) {
    MCM_TRY(randn_fill_launch(c->noise_buf, n, ns.seed, (unsigned long long)index, st));
    *out = c->noise_buf;
  }
  return 0;
}

// All steps' AdaLN modulation vectors in one batched pass (row r = step * B + sample): the arithmetic per row is that of the
// per-step path (same kernels, same K order), so the sampled result does not depend on whether the hoist is taken.
int precompute_mod_all(mcm_ctx* c, const mcm_sampler* s, int B, cudaStream_t st) {
  const int S = s->n_steps, D = c->D, E = c->E, fp = c->fmt_prec();
  const size_t R = (size_t)S * B;
  MCM_CHECK(c->cond_ready && c->cond_batch >= B, "mcm_prepare_conditions must be called first (for at least this batch)");
  if (R > c->hoist_rows) {
    c->drop_graphs();            // captured steps gather from the OLD table address
    MCM_TRY(alloc_f32(c, &c->mod_all, R * c->mod_total));
    MCM_TRY(alloc_f32(c, &c->emb_all, R * E));
    MCM_TRY(dev_alloc(c, reinterpret_cast<void**>(&c->t_all), R * sizeof(long long)));
    if (c->step_buf == nullptr) MCM_TRY(dev_alloc(c, reinterpret_cast<void**>(&c->step_buf), 256));
    MCM_TRY(alloc_op(c, &c->te_all, R * D, D, true));
    MCM_TRY(alloc_op(c, &c->t1_all, R * E, E, true));
    MCM_TRY(alloc_op(c, &c->embop_all, R * E, E, true));
    c->hoist_rows = R;
  }
  std::vector<long long> th(R);
  for (int i = 0; i < S; ++i)
    for (int b = 0; b < B; ++b) th[(size_t)i * B + b] = s->timestep_map[i];
  MCM_CUDA(cudaMemcpyAsync(c->t_all, th.data(), R * sizeof(long long), cudaMemcpyHostToDevice, st));   // pageable: staged before return
  MCM_TRY(timestep_embedding_launch(c->t_all, 0, (int)R, D, c->te_all, fp, st));
  {
    GemmProblem g = linear_problem(c->te_all, (int)R, c->w_te0, E, D, fp);
    g.seg[0] = seg_default(E, 0);
    g.seg[0].bias = c->b_te0; g.seg[0].flags = EPI_SILU; g.seg[0].op = c->t1_all; g.seg[0].op_fmt = fp;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  {  // + xf_proj[sample]: one GEMM batch per step, the addend indexed by the row within the batch
    GemmProblem g = linear_problem(c->t1_all, B, c->w_te2, E, E, fp);
    g.a_batches = S; g.batches = S; g.out_rows_per_outer = B;
    g.seg[0] = seg_default(E, 0);
    g.seg[0].bias = c->b_te2; g.seg[0].addend = c->xfproj32; g.seg[0].out32 = c->emb_all; g.seg[0].ld32 = E;
    g.seg[0].flags = EPI_ADDEND_BCAST;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  MCM_TRY(pack_op_launch(c->emb_all, (int)R, E, E, true, c->embop_all, fp, st));
  {
    GemmProblem g = linear_problem(c->embop_all, (int)R, c->w_mod, c->mod_total, E, fp);
    g.seg[0] = seg_default(c->mod_total, 0);
    g.seg[0].bias = c->b_mod; g.seg[0].out32 = c->mod_all; g.seg[0].ld32 = c->mod_total;
    MCM_TRY(gemm_tc_launch(g, st));
  }
  return 0;
}

int run_sampler(mcm_ctx* c, const mcm_sampler* s, int B, const NoiseSource& ns, float* x_io, cudaStream_t st) {
  const size_t rows = (size_t)B * c->T;
  const size_t n = rows * c->IN;
  const bool stochastic = s->mode == 1 || s->eta != 0.f;
  if (stochastic && !ns.dev) MCM_TRY(ensure_noise_buf(c));
  // hoist the timestep-conditioned modulation out of the loop while the table of all steps stays small (<= 2 GB: t2m B = 256 x 50 steps = 1 GB)
  static const size_t hoist_cap = [] { const char* e = getenv("MCM_HOIST_MOD_MB"); return (size_t)(e ? atoi(e) : 2048) << 20; }();
  c->hoist_active = c->hoist_mod && !timing_enabled() && c->mod_total % 4 == 0 &&
                    (size_t)s->n_steps * B * c->mod_total * 4 <= hoist_cap;
  if (c->hoist_active) {
    const int rc = precompute_mod_all(c, s, B, st);
    if (rc != 0) { c->hoist_active = false; return rc; }
  }
  struct HoistOff { mcm_ctx* c; ~HoistOff() { c->hoist_active = false; } } hoist_off{c};
  // x_io holds x_T on entry and x_0 on exit; xop must already hold the operand copy of x_T
  for (int i = s->n_steps - 1; i >= 0; --i) {
    if (c->hoist_active) MCM_TRY(fill_timesteps_launch(c->step_buf, (long long)i, 1, st));
    MCM_TRY(run_denoiser_step(c, B, s->timestep_map[i], st));
    const float* noise = nullptr;
    if (stochastic && i != 0) MCM_TRY(step_noise_ptr(c, ns, i, n, st, &noise));
    if (s->mode == 0) {
      DdimCoefs k{s->sqrt_recip_alphas_cumprod[i], s->sqrt_recipm1_alphas_cumprod[i], s->alphas_cumprod[i],
                  s->alphas_cumprod_prev[i], s->eta, (s->eta != 0.f && i != 0) ? 1 : 0, s->model_mean_type == 1 ? 1 : 0};
      MCM_TRY(ddim_update_launch(x_io, c->eps32, noise, x_io, rows, c->IN, k, c->xop, c->fmt_prec(), st));
    } else {
      DdpmCoefs k{s->sqrt_recip_alphas_cumprod[i], s->sqrt_recipm1_alphas_cumprod[i], s->posterior_mean_coef1[i],
                  s->posterior_mean_coef2[i], s->posterior_log_variance_clipped[i], i != 0 ? 1 : 0,
                  s->model_mean_type == 1 ? 1 : 0};
      MCM_TRY(ddpm_update_launch(x_io, c->eps32, noise, x_io, rows, c->IN, k, c->xop, c->fmt_prec(), st));
    }
  }
  return 0;
}

int check_sampler(const mcm_sampler* s) {
  MCM_CHECK(s != nullptr && s->n_steps > 0 && s->timestep_map != nullptr, "bad sampler description");
  MCM_CHECK(s->mode == 0 || s->mode == 1, "sampler mode must be 0 (DDIM) or 1 (DDPM)");
  MCM_CHECK(s->model_mean_type == 0 || s->model_mean_type == 1, "model_mean_type must be 0 (epsilon) or 1 (start_x)");
  MCM_CHECK(s->sqrt_recip_alphas_cumprod && s->sqrt_recipm1_alphas_cumprod, "missing sampler tables");
  if (s->mode == 0) {
    MCM_CHECK(s->alphas_cumprod && s->alphas_cumprod_prev, "DDIM needs alphas_cumprod(_prev)");
  } else {
    MCM_CHECK(s->posterior_mean_coef1 && s->posterior_mean_coef2 && s->posterior_log_variance_clipped, "DDPM needs posterior tables");
  }
  return 0;
}

}  // namespace

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

const char* mcm_last_error(void) { return g_last_error.c_str(); }
const char* mcm_version(void) { return "motioncraft_b200 0.1.0 (sm_100a, tcgen05)"; }
unsigned long long mcm_gemm_launches(void) { return gemm_tc_launch_count() + fused_block_launch_count(); }
unsigned long long mcm_kernel_launches(void) { return gemm_tc_launch_count() + fused_block_launch_count() + elementwise_launch_count(); }

int mcm_debug_read(unsigned long long* out, int reset) {
  // out[0..16): GEMM epilogue phases (MCM_DEBUG_EPI=3); when MCM_FUSED_PROF=1 the fused kernel's phase clocks instead
  unsigned long long f[32];
  MCM_TRY(fused_block_prof_read(f, reset));
  bool any = false;
  for (int i = 0; i < 32; ++i) any = any || f[i] != 0;
  if (any) {
    for (int i = 0; i < 16; ++i) out[i] = f[i];
    return 0;
  }
  return gemm_tc_debug_read(out, reset);
}
int mcm_debug_read32(unsigned long long* out, int reset) { return fused_block_prof_read(out, reset); }
int mcm_debug_copy(mcm_ctx* c, int what, void* dst_dev, long long bytes) {
  MCM_CHECK(c != nullptr && dst_dev != nullptr && bytes > 0, "bad argument");
  const void* src = what == 0 ? c->fused_dbg : c->ws[0].hid;
  MCM_CHECK(src != nullptr, "no such debug buffer");
  MCM_CUDA(cudaDeviceSynchronize());
  MCM_CUDA(cudaMemcpy(dst_dev, src, (size_t)bytes, cudaMemcpyDeviceToDevice));
  return 0;
}
void mcm_timing_enable(int on) { timing_enable(on != 0); }
int mcm_timing_collect(double* ms, unsigned long long* launches, double* flops) {
  MCM_CHECK(ms && launches && flops, "null argument");
  return timing_collect(ms, launches, flops);
}

int mcm_create(const mcm_config* cfg, mcm_ctx** out) {
  MCM_CHECK(cfg != nullptr && out != nullptr, "null argument");
  *out = nullptr;
  {
    // Kernel attributes (opt-in shared memory size, cluster occupancy) are per DEVICE and are set up once per process for
    // the device that is current at the first mcm_create: the deployment model is one process per GPU (torchrun /
    // bench.py --gpus N).  A context on another device of the same process is refused instead of failing at launch.
    static int g_device = -1;
    int dev = -1;
    MCM_CUDA(cudaGetDevice(&dev));
    if (g_device < 0) g_device = dev;
    MCM_CHECK(dev == g_device, "this process already runs motioncraft_b200 contexts on another CUDA device: use one process per GPU");
  }
  MCM_TRY(gemm_tc_init());
  MCM_TRY(elementwise_init());
  MCM_CHECK(cfg->seq_len % cfg->num_heads == 0 && cfg->latent_dim % cfg->num_heads == 0, "heads must divide seq_len and latent_dim");
  MCM_CHECK(cfg->seq_len % 4 == 0 && cfg->seq_len <= 1024, "seq_len must be a multiple of 4 and <= 1024");
  MCM_CHECK(cfg->latent_dim % 32 == 0 && cfg->latent_dim <= 1024, "latent_dim must be a multiple of 32 and <= 1024");
  MCM_CHECK(cfg->text_latent_dim % 8 == 0 && cfg->text_latent_dim <= 1024, "text_latent_dim must be a multiple of 8");
  MCM_CHECK(cfg->time_embed_dim % 8 == 0 && cfg->ffn_dim % 8 == 0, "time_embed_dim / ffn_dim must be multiples of 8");
  MCM_CHECK((cfg->latent_dim / cfg->num_heads) % 8 == 0, "latent head dim must be a multiple of 8");
  MCM_CHECK(cfg->max_batch >= 1 && cfg->num_layers >= 1 && cfg->num_ctrl_blocks >= 0 && cfg->num_ctrl_blocks < cfg->num_layers, "bad sizes");
  mcm_ctx* c = new mcm_ctx();
  c->cfg = *cfg;
  c->T = cfg->seq_len; c->Tp = rup(c->T, 8); c->D = cfg->latent_dim; c->E = cfg->time_embed_dim; c->F = cfg->ffn_dim;
  c->L = cfg->text_latent_dim; c->H = cfg->num_heads; c->IN = cfg->input_feats; c->INp = rup(c->IN, 8);
  c->NTmax = cfg->max_text_tokens > 0 ? cfg->max_text_tokens : 77; c->NTp = rup(c->NTmax, 8);
  c->nL = cfg->num_layers; c->nC = cfg->num_ctrl_blocks; c->Cin = cfg->ctrl_cond_feats; c->Cinp = rup(c->Cin > 0 ? c->Cin : 8, 8);
  c->hdT = c->T / c->H; c->hdD = c->D / c->H; c->Bmax = cfg->max_batch;
  if (const char* e = getenv("MCM_CHUNK")) c->chunk = atoi(e);
  c->mod_total = (c->nL + c->nC) * (2 * c->T + 4 * c->D);

  const size_t B = c->Bmax, R1 = B * c->T, R2 = B * c->D;
  const bool lo = cfg->precise_all != 0;
  auto fail = [&](int) { delete c; return 1; };
  size_t szA = smax(smax(R2 * c->Tp, R1 * c->D), smax(B * c->NTmax * c->L, R2 * c->NTp));
  size_t szB = smax(smax(R1 * c->D, R1 * c->F), R2 * c->NTp);
  if (c->nC > 0) szB = smax(szB, R1 * c->Cinp);
  size_t szC = smax(smax(R2 * c->Tp, R1 * c->D), R2 * c->NTp);
  if (alloc_f32(c, &c->ws[0].h32, R1 * c->D)) return fail(0);
  if (alloc_f32(c, &c->ws[0].f32A, smax(smax(R2 * c->T, R1 * c->D), R2 * c->NTp))) return fail(0);
  if (alloc_f32(c, &c->ws[0].f32B, R1 * c->D)) return fail(0);
  if (alloc_f32(c, &c->mod32, B * c->mod_total)) return fail(0);
  if (alloc_f32(c, &c->emb32, B * c->E)) return fail(0);
  if (alloc_f32(c, &c->xfproj32, B * c->E)) return fail(0);
  if (alloc_f32(c, &c->eps32, R1 * c->IN)) return fail(0);
  if (alloc_f32(c, &c->x32, R1 * c->IN)) return fail(0);
  if (alloc_op(c, &c->ws[0].opA, szA, 8, lo)) return fail(0);
  if (alloc_op(c, &c->ws[0].opB, szB, 8, true)) return fail(0);      // lo: also stages the bf16x2 control condition
  if (alloc_op(c, &c->ws[0].opC, szC, 8, lo)) return fail(0);
  if (alloc_op(c, &c->ws[0].opD, R1 * c->D, 8, lo)) return fail(0);
  if (alloc_op(c, &c->ws[0].hop, R1 * c->D, c->D, true)) return fail(0);
  if (alloc_op(c, &c->ws[0].ctxT_sa, B * c->T * c->Tp, c->Tp, lo)) return fail(0);
  if (alloc_op(c, &c->xop, R1 * c->INp, c->INp, true)) return fail(0);
  if (alloc_op(c, &c->te_op, B * c->D, c->D, true)) return fail(0);
  if (alloc_op(c, &c->t1_op, B * c->E, c->E, true)) return fail(0);
  if (alloc_op(c, &c->emb_op, B * c->E, c->E, true)) return fail(0);
  if (c->nC > 0) {
    if (alloc_f32(c, &c->cc32, R1 * c->D)) return fail(0);
    if (alloc_f32(c, &c->ws[0].c32, R1 * c->D)) return fail(0);
    if (alloc_op(c, &c->cc_op, R1 * c->D, c->D, lo)) return fail(0);
    if (alloc_op(c, &c->ws[0].c_op, R1 * c->D, c->D, lo)) return fail(0);
  }
  if (const char* e = getenv("MCM_FUSED")) c->fused = atoi(e);
  if (const char* e = getenv("MCM_FUSED_SA")) c->fused_sa = atoi(e);
  if (const char* e = getenv("MCM_FUSED_MIN_ROWS")) c->fused_min_rows = atoi(e);
  if (const char* e = getenv("MCM_FUSED_SA_MIN_ROWS")) c->fused_sa_min_rows = atoi(e);
  c->ws[0].hid = nullptr;
  if (!lo && fused_block_supported(c->T, c->D, c->F, c->H)) {
    if (dev_alloc(c, &c->ws[0].hid, fused_block_hid_bytes())) return fail(0);
  }
  int dual_env = -1;                     // MCM_DUAL only picks the default; the second scratch set always exists
  if (const char* e = getenv("MCM_DUAL")) dual_env = atoi(e);
  if (const char* e = getenv("MCM_GRAPH")) c->use_graph = atoi(e);
  if (const char* e = getenv("MCM_SPLIT_SMS")) c->split_sms = atoi(e);
  if (const char* e = getenv("MCM_HOIST_MOD")) c->hoist_mod = atoi(e);
  if (dev_alloc(c, reinterpret_cast<void**>(&c->t_buf), (size_t)c->Bmax * sizeof(long long))) return fail(0);
  if (cudaStreamCreateWithFlags(&c->s0, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    set_error("could not create the graph stream / events");
    return fail(0);
  }
  if (c->Bmax < 2) c->dual = 0;
  if (c->dual) {
    // second scratch set for the second half of the batch (only what a pass through the layer stack touches)
    const size_t Bh = c->Bmax / 2, H1 = Bh * c->T, H2 = Bh * c->D;
    Scratch& w = c->ws[1];
    std::memset(&w, 0, sizeof(w));
    if (alloc_f32(c, &w.h32, H1 * c->D)) return fail(0);
    if (alloc_f32(c, &w.f32A, smax(H2 * c->T, H1 * c->D))) return fail(0);
    if (alloc_f32(c, &w.f32B, H1 * c->D)) return fail(0);
    if (alloc_op(c, &w.opA, smax(H2 * c->Tp, H1 * c->D), 8, lo)) return fail(0);
    if (alloc_op(c, &w.opB, smax(H1 * c->D, H1 * c->F), 8, lo)) return fail(0);
    if (alloc_op(c, &w.opC, smax(H2 * c->Tp, H1 * c->D), 8, lo)) return fail(0);
    if (alloc_op(c, &w.opD, H1 * c->D, 8, lo)) return fail(0);
    if (alloc_op(c, &w.hop, H1 * c->D, c->D, true)) return fail(0);
    if (alloc_op(c, &w.ctxT_sa, Bh * c->T * c->Tp, c->Tp, lo)) return fail(0);
    if (c->ws[0].hid != nullptr && dev_alloc(c, &w.hid, fused_block_hid_bytes())) return fail(0);
    if (c->nC > 0) {
      if (alloc_f32(c, &w.c32, H1 * c->D)) return fail(0);
      if (alloc_op(c, &w.c_op, H1 * c->D, c->D, lo)) return fail(0);
    }
    if (cudaStreamCreateWithFlags(&c->s1, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      set_error("could not create the second stream / events");
      return fail(0);
    }
  }
  if (c->dual && dual_env >= 0) c->dual = dual_env ? 1 : 0;
  *out = c;
  return 0;
}

void mcm_destroy(mcm_ctx* ctx) { delete ctx; }

int mcm_set_option(mcm_ctx* c, const char* name, int value) {
  MCM_CHECK(c != nullptr && name != nullptr, "bad argument");
  const std::string n(name);
  if (n == "dual") {
    MCM_CHECK(value == 0 || c->s1 != nullptr, "dual-stream mode was disabled at creation (no second scratch set)");
    c->dual = value;
  } else if (n == "graph") {
    c->use_graph = value;
  } else if (n == "split_sms") {
    c->split_sms = value;
  } else if (n == "hoist_mod") {
    c->hoist_mod = value;
  } else if (n == "chunk") {
    c->chunk = value;
  } else if (n == "fused") {
    c->fused = value;
  } else if (n == "fused_sa") {
    c->fused_sa = value;
  } else if (n == "fused_min_rows") {
    c->fused_min_rows = value;
  } else if (n == "fused_sa_min_rows") {
    c->fused_sa_min_rows = value;
  } else if (n == "fused_stop") {
    MCM_CHECK(value >= 0 && value <= 7, "fused_stop must be 0..7");
    c->fused_stop = value;
    if (value != 0 && c->fused_dbg == nullptr)
      MCM_TRY(dev_alloc(c, &c->fused_dbg, (size_t)c->Bmax * c->T * c->D * 2));
  } else {
    set_error("unknown option: " + n);
    return 1;
  }
  c->drop_graphs();      // captured step graphs embed the old schedule; the next sampler step re-captures
  return 0;
}

int mcm_set_param(mcm_ctx* ctx, const char* name, const float* dev_ptr, long long numel) {
  MCM_CHECK(ctx && name && dev_ptr && numel > 0, "bad argument");
  ctx->params[name] = {dev_ptr, numel};
  return 0;
}

int mcm_finalize_params(mcm_ctx* c, void* stream) {
  MCM_CHECK(c != nullptr, "null context");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // A repeated call (after load_state_dict: every parameter set again with mcm_set_param) re-packs into the SAME operand
  // buffers, so captured CUDA graphs and the workspace stay valid; only the step-invariant condition work, which depends
  // on the weights, must be redone by the caller (mcm_prepare_conditions).
  if (c->finalized) {
    c->finalized = false;
    c->cond_ready = false;
  }
  const int T = c->T, D = c->D, E = c->E, IN = c->IN;
  const int fp = c->fmt_prec();
  MCM_TRY(pack_weight(c, {{"joint_embed.weight", D}}, IN, fp, &c->w_joint, st));
  MCM_TRY(own_f32(c, {{"joint_embed.bias", D}}, &c->b_joint, st));
  MCM_TRY(pack_weight(c, {{"time_embed.0.weight", E}}, D, fp, &c->w_te0, st));
  MCM_TRY(own_f32(c, {{"time_embed.0.bias", E}}, &c->b_te0, st));
  MCM_TRY(pack_weight(c, {{"time_embed.2.weight", E}}, E, fp, &c->w_te2, st));
  MCM_TRY(own_f32(c, {{"time_embed.2.bias", E}}, &c->b_te2, st));
  MCM_TRY(pack_weight(c, {{"out.weight", IN}}, D, fp, &c->w_out, st));
  MCM_TRY(own_f32(c, {{"out.bias", IN}}, &c->b_out, st));
  MCM_TRY(own_f32(c, {{"sequence_embedding", (long long)T * D}}, &c->seq_emb, st));

  const int per = 2 * T + 4 * D;
  c->blocks.resize(c->nL + c->nC);
  std::vector<std::pair<std::string, int>> mod_w;
  std::vector<std::pair<std::string, long long>> mod_b;
  for (int i = 0; i < c->nL + c->nC; ++i) {
    const std::string pfx = i < c->nL ? "temporal_decoder_blocks." + std::to_string(i)
                                      : "controlnet." + std::to_string(i - c->nL) + ".copied_block";
    MCM_TRY(build_block(c, pfx, &c->blocks[i], i * per, st));
    mod_w.push_back({pfx + ".sa_block.proj_out.emb_layers.1.weight", 2 * T});
    mod_w.push_back({pfx + ".ca_block.proj_out.emb_layers.1.weight", 2 * D});
    mod_w.push_back({pfx + ".ffn_temporal.proj_out.emb_layers.1.weight", 2 * D});
    mod_b.push_back({pfx + ".sa_block.proj_out.emb_layers.1.bias", 2 * T});
    mod_b.push_back({pfx + ".ca_block.proj_out.emb_layers.1.bias", 2 * D});
    mod_b.push_back({pfx + ".ffn_temporal.proj_out.emb_layers.1.bias", 2 * D});
  }
  MCM_TRY(pack_weight(c, mod_w, E, fp, &c->w_mod, st));
  MCM_TRY(own_f32(c, mod_b, &c->b_mod, st));
  c->ctrls.resize(c->nC);
  for (int j = 0; j < c->nC; ++j) {
    const std::string pfx = "controlnet." + std::to_string(j);
    Ctrl& ct = c->ctrls[j];
    ct.before_w = OpPtr{nullptr, nullptr, 0};
    ct.before_b = nullptr;
    if (j == 0) {
      MCM_TRY(pack_weight(c, {{pfx + ".before_proj.weight", D}}, D, c->fmt_fast(), &ct.before_w, st));
      MCM_TRY(own_f32(c, {{pfx + ".before_proj.bias", D}}, &ct.before_b, st));
    }
    MCM_TRY(pack_weight(c, {{pfx + ".after_proj.weight", D}}, D, c->fmt_fast(), &ct.after_w, st));
    MCM_TRY(own_f32(c, {{pfx + ".after_proj.bias", D}}, &ct.after_b, st));
  }
  if (c->nC > 0) {
    MCM_TRY(pack_weight(c, {{"control_cond_input.weight", D}}, c->Cin, fp, &c->w_cci, st));
    MCM_TRY(own_f32(c, {{"control_cond_input.bias", D}}, &c->b_cci, st));
  }
  MCM_CUDA(cudaStreamSynchronize(st));
  c->params.clear();   // borrowed pointers are not used after this point
  c->finalized = true;
  return 0;
}

int mcm_prepare_conditions(mcm_ctx* c, int B, const float* xf_out, int n_tokens, const float* xf_proj, const float* cond,
                           int c_len, void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_CHECK(xf_out && xf_proj, "xf_out / xf_proj are required (the text encoder stays outside this library)");
  MCM_CHECK(n_tokens >= 1 && n_tokens <= c->NTmax, "n_tokens exceeds max_text_tokens");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int D = c->D, L = c->L, H = c->H, hdD = c->hdD, T = c->T;
  const int N = n_tokens, Np = rup(N, 8);
  const int ff = c->fmt_fast();
  MCM_CUDA(cudaMemcpyAsync(c->xfproj32, xf_proj, (size_t)B * c->E * 4, cudaMemcpyDeviceToDevice, st));
  for (size_t i = 0; i < c->blocks.size(); ++i) {
    const Block& k = c->blocks[i];
    const OpPtr xfn = view(c->ws[0].opA, L), vT = view(c->ws[0].opB, Np), pT = view(c->ws[0].opC, Np);
    // LN_L(xf) -> [B*N, L]
    MCM_TRY(ln_rows_launch(xf_out, B * N, L, L, k.ca_tn_w, k.ca_tn_b, nullptr, nullptr, 4, N, false, xfn, ff, st));
    {  // key | value = LN(xf) W^T + b, both written transposed: [B, D, Np] (tokens contiguous)
      GemmProblem g;
      std::memset(&g, 0, sizeof(g));
      g.a = xfn; g.a_rows = N; g.a_k = L; g.a_batches = B;
      g.b = k.ca_wkv; g.b_rows = 2 * D; g.b_k = L; g.b_batches = 1;
      g.fmt = ff; g.M = N; g.K = L; g.batches = B; g.inner = 1;
      g.out_rows_per_outer = N; g.trans_rows = D;
      g.nseg = 2;
      g.seg[0] = seg_default(D, 0);
      g.seg[0].bias = k.ca_bkv; g.seg[0].out32 = c->ws[0].f32A; g.seg[0].ld32 = Np; g.seg[0].flags = EPI_TRANSPOSED;
      g.seg[1] = seg_default(D, D);
      g.seg[1].bias = k.ca_bkv + D; g.seg[1].op = vT; g.seg[1].op_fmt = ff; g.seg[1].flags = EPI_TRANSPOSED;
      MCM_TRY(gemm_tc_launch(g, st));
    }
    // softmax over the N text tokens (dim=1 of [B, N, H, hd], efficient_attention.py:78)
    MCM_TRY(softmax_seg_launch(c->ws[0].f32A, B * D, N, Np, N, pT, ff, st));
    {  // ctx[b, h] = softmax(key)_h^T value_h  (hd x hd), stored transposed for the per-step q * ctx GEMM
      GemmProblem g;
      std::memset(&g, 0, sizeof(g));
      g.a = pT; g.a_rows = hdD; g.a_k = Np; g.a_batches = B * H;
      g.b = vT; g.b_rows = hdD; g.b_k = Np; g.b_batches = B * H; g.b_batched = 1;
      g.fmt = ff; g.M = hdD; g.K = N; g.batches = B * H; g.inner = 1;
      g.out_rows_per_outer = hdD; g.trans_rows = hdD;
      g.nseg = 1;
      g.seg[0] = seg_default(hdD, 0);
      g.seg[0].op = k.ca_ctxT; g.seg[0].op_fmt = ff; g.seg[0].flags = EPI_TRANSPOSED;
      MCM_TRY(gemm_tc_launch(g, st));
    }
  }
  c->have_c = false;
  if (cond != nullptr) {
    MCM_CHECK(c->nC > 0, "a control condition was given but the context has no control blocks");
    MCM_CHECK(c_len >= 1 && c_len <= T, "control condition longer than seq_len");
    // forward_c (controlnet_mcm.py:155-166): control_cond_input(c), zero-pad to T, + sequence_embedding[:len_c]
    const int fp = c->fmt_prec();
    const OpPtr cin = view(c->ws[0].opB, c->Cinp);
    MCM_TRY(pack_op_launch(cond, B * c_len, c->Cin, c->Cin, false, cin, fp, st));
    MCM_CUDA(cudaMemsetAsync(c->cc32, 0, (size_t)B * T * D * 4, st));
    MCM_CUDA(cudaMemsetAsync(c->cc_op.hi, 0, (size_t)B * T * D * 2, st));
    if (c->cc_op.lo) MCM_CUDA(cudaMemsetAsync(c->cc_op.lo, 0, (size_t)B * T * D * 2, st));
    GemmProblem g;
    std::memset(&g, 0, sizeof(g));
    g.a = cin; g.a_rows = c_len; g.a_k = c->Cinp; g.a_batches = B;
    g.b = c->w_cci; g.b_rows = D; g.b_k = c->Cinp; g.b_batches = 1;
    g.fmt = fp; g.M = c_len; g.K = c->Cin; g.batches = B; g.inner = 1;
    g.out_rows_per_outer = T;
    g.nseg = 1;
    g.seg[0] = seg_default(D, 0);
    g.seg[0].bias = c->b_cci; g.seg[0].addend = c->seq_emb; g.seg[0].flags = EPI_ADDEND_BCAST;
    g.seg[0].out32 = c->cc32; g.seg[0].ld32 = D;
    g.seg[0].op = view(c->cc_op, D); g.seg[0].op_fmt = c->fmt_fast();
    MCM_TRY(gemm_tc_launch(g, st));
    c->have_c = true;
  }
  c->cond_ready = true;
  c->cond_batch = B;
  return 0;
}

int mcm_denoise(mcm_ctx* c, int B, const float* x, const long long* timesteps, int t_uniform, float* eps_out,
                void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_CHECK(x && eps_out, "null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MCM_TRY(pack_op_launch(x, B * c->T, c->IN, c->IN, false, c->xop, c->fmt_prec(), st));
  return run_denoiser(c, B, timesteps, t_uniform, eps_out, st);
}

int mcm_block_forward(mcm_ctx* c, int kind, int index, int B, float* x_inout, const float* emb, void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_CHECK(x_inout && emb, "null tensor");
  MCM_CHECK(c->cond_ready && c->cond_batch >= B, "mcm_prepare_conditions must be called first");
  MCM_CHECK((kind == 0 && index >= 0 && index < c->nL) || (kind == 1 && index >= 0 && index < c->nC), "no such block");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bi = kind == 0 ? index : c->nL + index;
  MCM_TRY(run_mod(c, B, emb, bi, 1, st));
  const OpPtr none{nullptr, nullptr, 0};
  return run_block(c, c->ws[0], c->blocks[bi], B, x_inout, c->mod32, c->mod_total, none, 0, st);
}

int mcm_layers_forward(mcm_ctx* c, int B, const float* h, const float* emb, float* out, void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_CHECK(h && emb && out, "null tensor");
  MCM_CHECK(c->cond_ready && c->cond_batch >= B, "mcm_prepare_conditions must be called first");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MCM_CUDA(cudaMemcpyAsync(c->ws[0].h32, h, (size_t)B * c->T * c->D * 4, cudaMemcpyDeviceToDevice, st));
  MCM_TRY(run_mod(c, B, emb, 0, (int)c->blocks.size(), st));
  return run_stack(c, c->ws[0], 0, B, out, st);
}

int mcm_sample(mcm_ctx* c, const mcm_sampler* s, int B, const float* x_T, const float* step_noise, float* x0_out,
               void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_TRY(check_sampler(s));
  MCM_CHECK(x_T && x0_out, "null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * c->T * c->IN;
  if (x0_out != x_T) MCM_CUDA(cudaMemcpyAsync(x0_out, x_T, n * 4, cudaMemcpyDeviceToDevice, st));
  MCM_TRY(pack_op_launch(x0_out, B * c->T, c->IN, c->IN, false, c->xop, c->fmt_prec(), st));
  NoiseSource ns;
  ns.dev = step_noise; ns.generate = step_noise == nullptr; ns.seed = s->seed;
  return run_sampler(c, s, B, ns, x0_out, st);
}

int mcm_sample_repaint(mcm_ctx* c, const mcm_sampler* s, const mcm_repaint* r, int B, const float* x_T, float* x0_out,
                       void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_TRY(check_sampler(s));
  MCM_CHECK(s->mode == 0 && s->eta == 0.f, "mcm_sample_repaint: DDIM with eta = 0 only (what MotionDiffusion passes)");
  MCM_CHECK(r != nullptr && x_T && x0_out && r->gt && r->keep_mask, "mcm_sample_repaint: null argument");
  // noise_seq == NULL: every draw the loop actually READS (the blend noise of a denoise call, the undo noise) is generated
  // on the device, Philox keyed by (sampler seed, draw index); the unread eta draw of each denoise call costs nothing.
  NoiseSource ns;
  ns.dev = r->noise_seq; ns.generate = r->noise_seq == nullptr; ns.seed = s->seed;
  if (ns.generate) MCM_TRY(ensure_noise_buf(c));
  MCM_CHECK(r->n_times == 0 || (r->times != nullptr && r->betas != nullptr), "mcm_sample_repaint: schedule without times / betas");
  MCM_CHECK(r->overlap_len >= 0 && r->overlap_len <= c->T, "mcm_sample_repaint: overlap_len out of range");
  MCM_CHECK(!(r->add_blend && r->overlap_len > 0) || r->blend_w != nullptr, "mcm_sample_repaint: addBlend needs blend_w");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t rows = (size_t)B * c->T, n = rows * c->IN;
  if (x0_out != x_T) MCM_CUDA(cudaMemcpyAsync(x0_out, x_T, n * 4, cudaMemcpyDeviceToDevice, st));
  MCM_TRY(pack_op_launch(x0_out, B * c->T, c->IN, c->IN, false, c->xop, c->fmt_prec(), st));
  long long draw = 0;
  const OpPtr none{nullptr, nullptr, 0};
  auto denoise = [&](int i) -> int {
    MCM_CHECK(i >= 0 && i < s->n_steps, "mcm_sample_repaint: time index outside the sampler tables");
    MCM_CHECK(ns.generate || draw + 2 <= r->n_draws, "mcm_sample_repaint: noise_seq too short");
    MCM_TRY(run_denoiser_step(c, B, s->timestep_map[i], st));
    DdimCoefs k{s->sqrt_recip_alphas_cumprod[i], s->sqrt_recipm1_alphas_cumprod[i], s->alphas_cumprod[i],
                s->alphas_cumprod_prev[i], 0.f, 0, s->model_mean_type == 1 ? 1 : 0};
    MCM_TRY(ddim_update_launch(x0_out, c->eps32, nullptr, x0_out, rows, c->IN, k, none, c->fmt_prec(), st));
    const float abp = s->alphas_cumprod_prev[i];
    const float noise_w = sqrtf(1.f - abp), gt_w = sqrtf(abp);
    const bool blend = r->add_blend && r->overlap_len > 0 && noise_w < 0.2f;     // :872
    const float* bn = nullptr;
    MCM_TRY(step_noise_ptr(c, ns, draw + 1, n, st, &bn));
    MCM_TRY(repaint_blend_launch(x0_out, r->gt, r->keep_mask, bn, rows, c->IN, c->T, gt_w,
                                 noise_w, blend ? r->blend_w : nullptr, r->overlap_len, c->xop, c->fmt_prec(), st));
    draw += 2;
    return 0;
  };
  if (r->n_times == 0) {
    for (int i = s->n_steps - 1; i >= 0; --i) MCM_TRY(denoise(i));
    return 0;
  }
  for (int k = 0; k + 1 < r->n_times; ++k) {
    const int t_last = r->times[k], t_cur = r->times[k + 1];
    if (t_cur < t_last) {
      MCM_TRY(denoise(t_last));
    } else {
      MCM_CHECK(t_last >= 0 && t_last < s->n_steps, "mcm_sample_repaint: undo time outside the sampler tables");
      MCM_CHECK(ns.generate || draw + 1 <= r->n_draws, "mcm_sample_repaint: noise_seq too short");
      const float beta = r->betas[t_last];
      const float* un = nullptr;
      MCM_TRY(step_noise_ptr(c, ns, draw, n, st, &un));
      MCM_TRY(undo_launch(x0_out, un, rows, c->IN, sqrtf(1.f - beta), sqrtf(beta), c->xop,
                          c->fmt_prec(), st));
      draw += 1;
    }
  }
  return 0;
}

int mcm_sample_host(mcm_ctx* c, const mcm_sampler* s, int B, const float* x_T_host, const float* step_noise_host,
                    float* x0_out_host, void* stream) {
  MCM_TRY(check_batch(c, B));
  MCM_CHECK(x_T_host && x0_out_host, "null tensor");
  MCM_TRY(check_sampler(s));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * c->T * c->IN;
  MCM_CUDA(cudaMemcpyAsync(c->x32, x_T_host, n * 4, cudaMemcpyHostToDevice, st));
  MCM_TRY(pack_op_launch(c->x32, B * c->T, c->IN, c->IN, false, c->xop, c->fmt_prec(), st));
  NoiseSource ns;
  ns.host = step_noise_host; ns.generate = step_noise_host == nullptr; ns.seed = s->seed;
  MCM_TRY(run_sampler(c, s, B, ns, c->x32, st));
  MCM_CUDA(cudaMemcpyAsync(x0_out_host, c->x32, n * 4, cudaMemcpyDeviceToHost, st));
  MCM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int mcm_cfg_combine(const float* out_text, const float* out_none, double text_coef, double none_coef, float* out,
                    long long n, void* stream) {
  MCM_CHECK(out_text && out_none && out && n > 0, "bad argument");
  return axpby_launch(out_text, out_none, (float)text_coef, (float)none_coef, out, (size_t)n, reinterpret_cast<cudaStream_t>(stream));
}

int mcm_part_mix(const float* body_weight, const float* v, float* out, long long rows, int num_parts, int part_dim, void* stream) {
  MCM_CHECK(body_weight && v && out && rows > 0, "bad argument");
  return part_mix_launch(body_weight, v, out, (size_t)rows, num_parts, part_dim, reinterpret_cast<cudaStream_t>(stream));
}

int mcm_test_randn(float* out_dev, long long n, unsigned long long seed, unsigned long long sub, void* stream) {
  MCM_CHECK(out_dev != nullptr && n > 0, "bad argument");
  return randn_fill_launch(out_dev, (size_t)n, seed, sub, reinterpret_cast<cudaStream_t>(stream));
}

int mcm_test_linear(int M, int N, int K, const float* A, const float* W, const float* bias, float* C, int fmt,
                    void* stream) {
  MCM_CHECK(M > 0 && N > 0 && K > 0 && A && W && C, "bad argument");
  MCM_CHECK(fmt == OP_F16 || fmt == OP_BF16X2, "fmt must be 0 (fp16) or 1 (bf16x2)");
  MCM_TRY(gemm_tc_init());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int Kp = rup(K, 8);
  void *ah = nullptr, *al = nullptr, *wh = nullptr, *wl = nullptr;
  MCM_CUDA(cudaMalloc(&ah, (size_t)M * Kp * 2));
  MCM_CUDA(cudaMalloc(&al, (size_t)M * Kp * 2));
  MCM_CUDA(cudaMalloc(&wh, (size_t)N * Kp * 2));
  MCM_CUDA(cudaMalloc(&wl, (size_t)N * Kp * 2));
  OpPtr a{ah, al, Kp}, w{wh, wl, Kp};
  int rc = pack_op_launch(A, M, K, K, false, a, fmt, st);
  if (!rc) rc = pack_op_launch(W, N, K, K, false, w, fmt, st);
  if (!rc) {
    GemmProblem g = linear_problem(a, M, w, N, K, fmt);
    g.seg[0] = seg_default(N, 0);
    g.seg[0].bias = bias; g.seg[0].out32 = C; g.seg[0].ld32 = N;
    rc = gemm_tc_launch(g, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(ah); cudaFree(al); cudaFree(wh); cudaFree(wl);
  if (!rc && e != cudaSuccess) {
    set_error(std::string("mcm_test_linear: ") + cudaGetErrorString(e));
    rc = 1;
  }
  return rc;
}

}  // extern "C"
