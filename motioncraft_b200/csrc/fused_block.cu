// motioncraft_b200 -- fused cross-attention + FFN token kernel (see fused_block.cuh).
//
// One persistent CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256-row tile; each CTA owns 128 rows:
//   warp 0      TMA producer : streams weight / context / hidden slabs (128 rows x 64 fp16, 128B swizzle) through a
//                              4-slot ring; each CTA loads its half of every B tile, all bytes land on the LEADER's barrier
//   warp 1      MMA issuer   : leader CTA only, tcgen05.mma.cta_group::2 (M = 256, N = 256 or 128), accumulators in TMEM
//   warps 2..9  compute      : row phases.  "P" phases read h rows (warp per row, coalesced) and write the A operand
//                              tile OPA (128 x 512 fp16, 8 swizzled slabs) in shared memory; "E" phases read the TMEM
//                              accumulator (thread = row) and apply bias / softmax / LayerNorm + AdaLN + SiLU / GELU,
//                              writing either OPA (the next GEMM's A operand), the hidden scratch (TMA store) or the
//                              residual stream h (TMA reduce-add).
// Per tile (reference lines in brackets):
//   P0  OPA = LN(h)                          [efficient_attention.py:73 norm]
//   G1  q = OPA Wq^T                   E1  OPA = softmax_head(q + bq)               [:73, :75 query softmax]
//   per sample s of the tile:
//   G2  y = OPA blockdiag(ctx[s])      E2  rows of s: OPA = SiLU(LN(y) (1+scale_s) + shift_s)   [:88-90, stylization_block.py:38-39]
//   G3  d = OPA Wo^T                   E3  h += d + bo ; OPA = fp16(h)              [:39-40 ; residual efficient_attention.py:91]
//   G4  u = OPA W1^T (4 quarters)      E4  HID = GELU(u + b1)                       [diffusion_transformer.py:26]
//   G5  y = HID W2^T                   E5  OPA = SiLU(LN(y + b2) (1+scale) + shift) [:26-27]
//   G6  d = OPA Wo2^T                  E6  h += d + bo2
// The MMA <-> compute hand-off uses one "accumulator half drained / operand written" barrier per TMEM half (tempty,
// on the leader, 16 warp arrivals) and one "accumulator half complete" barrier (tfull, multicast commit).
#include "fused_block.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "gemm_tc.cuh"
#include "ptx_extra.cuh"
#include "timing.cuh"

namespace mcm {
namespace {

constexpr int D = 512, F = 1024, H = 4, HD = 128;
constexpr int ROWS = 128;                  // rows per CTA
constexpr int SLAB = 16384;                // 128 rows x 64 fp16
constexpr int NSLOT = 4;
#ifndef MCM_FB_NCW
#define MCM_FB_NCW 8
#endif
constexpr int NCW = MCM_FB_NCW;            // compute warps: 8 or 16 (2 or 4 per TMEM lane quadrant)
constexpr int WPQ = NCW / 4;               // warps sharing a lane quadrant; each owns CPW accumulator columns
constexpr int CPW = D / WPQ;               // 256 or 128
constexpr int NCH = CPW / 32;              // 32-column chunks per warp and phase
constexpr int QCH = 256 / WPQ / 32;        // chunks per warp of one 256-column quarter of the FFN hidden layer
constexpr int THREADS = 64 + 32 * NCW;
constexpr int OPA_BYTES = (D / 64) * SLAB;
constexpr int RING_BYTES = NSLOT * SLAB;
constexpr int STG_BYTES = 32768;
constexpr int STG_PER_WARP = STG_BYTES / NCW;        // 4096 or 2048: E4's fp16 store staging (2 KB tiles)
constexpr int RED_SLOTS = OPA_BYTES / NCW / 4096;    // 4 or 2 rotating fp32 staging tiles per warp for the reductions into h
constexpr int BIAS_OFF = 28672;            // per-warp 128-byte bias broadcast slots (not during E4, which owns the staging region)
constexpr int SMEM_BYTES = OPA_BYTES + RING_BYTES + STG_BYTES + 1024;
constexpr int MAX_SAMPLES = 6;             // samples a 256-row tile may span (4 KB of staged AdaLN parameters each)
constexpr int XCH_OFF = MAX_SAMPLES * 4096;   // LayerNorm partial statistics (NCW x 32 x 2 floats <= 4 KB), inside the staging region
constexpr int TMEM_COLS = 512;
constexpr float L2E = 1.4426950408889634f;

struct FbMaps {
  CUtensorMap wq, ctx, wo, w1, hida, w2, wo2, hred, hidst;
};

struct FbParams {
  const float* h;
  int rows, T, batch, n_tiles, mod_ld, stop;
  int stagger;          // clock cycles by which the pairs that get one tile fewer start late (de-synchronises the phases)
  const float *ca_ln_w, *ca_ln_b, *ca_bq, *ca_pn_w, *ca_pn_b, *ca_scale, *ca_shift, *ca_bo;
  const float *f_b1, *f_b2, *f_pn_w, *f_pn_b, *f_scale, *f_shift, *f_bo;
  uint16_t* dbg;
  unsigned long long* prof;   // MCM_FUSED_PROF=1: summed cycles per phase (compute warps: [0,16), MMA warp: [16,24))
};

__device__ __forceinline__ void st_shared_v2u(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float silu_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + ex2_fast(-x * L2E)));
  return x * r;
}
// ---- packed fp32 pairs (sm_100a `fma.rn.f32x2` = one FFMA2 issue slot for two FMAs): the GELU phase is issue-bound
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// silu_fast for two elements (same operations per element)
__device__ __forceinline__ f32x2 silu2(f32x2 x) {
  float n0, n1;
  upk2(mul2(x, pk2(-L2E, -L2E)), n0, n1);
  float d0, d1;
  upk2(add2(pk2(ex2_fast(n0), ex2_fast(n1)), pk2(1.f, 1.f)), d0, d1);
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return mul2(x, pk2(r0, r1));
}
// erf-GELU as relu(x) - 0.5 |x| erfc(|x| / sqrt 2) with the Abramowitz-Stegun 7.1.26 erfc (|abs err| < 5e-7, far below
// the fp16 rounding of the result): 13 FP32 ops + 2 MUFU, no select
__device__ __forceinline__ float gelu_relu_erfc(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float e = ex2_fast(ax * ax * (-0.5f * L2E));
  return fmaf(-0.5f * ax, poly * t * e, fmaxf(x, 0.f));
}
// gelu_relu_erfc for two elements at once: the same operations in the same order per element (bit-identical results), the
// FMA / MUL chain issued as packed pairs
__device__ __forceinline__ void gelu_relu_erfc2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const f32x2 ax = pk2(a0, a1);
  const f32x2 one = pk2(1.f, 1.f);
  float d0, d1;
  upk2(fma2(pk2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f), ax, one), d0, d1);
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const f32x2 t = pk2(t0, t1);
  f32x2 poly = fma2(t, pk2(1.061405429f, 1.061405429f), pk2(-1.453152027f, -1.453152027f));
  poly = fma2(t, poly, pk2(1.421413741f, 1.421413741f));
  poly = fma2(t, poly, pk2(-0.284496736f, -0.284496736f));
  poly = fma2(t, poly, pk2(0.254829592f, 0.254829592f));
  float q0, q1;
  upk2(mul2(mul2(ax, ax), pk2(-0.5f * L2E, -0.5f * L2E)), q0, q1);
  const f32x2 e = pk2(ex2_fast(q0), ex2_fast(q1));
  const f32x2 pte = mul2(mul2(poly, t), e);
  const f32x2 mh = mul2(pk2(-0.5f, -0.5f), ax);
  upk2(fma2(mh, pte, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}
// shared-memory address of the 16-byte chunk holding columns [col, col + 8) of row `row` of the operand tile
__device__ __forceinline__ uint32_t opa_addr(uint32_t opa, int row, int col) {
  return opa + (uint32_t)(col >> 6) * SLAB + (uint32_t)row * 128u + (((((uint32_t)col >> 3) & 7u) ^ ((uint32_t)row & 7u)) << 4);
}
// v[j] += b[j] where lane j holds b[j]: through the warp's 128-byte shared slot (1 STS + 8 broadcast LDS.128)
__device__ __forceinline__ void add_bias32(float* slot, int lane, float mine, float* v) {
  slot[lane] = mine;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b4 = *reinterpret_cast<const float4*>(slot + 4 * q);
    upk2(add2(pk2(v[4 * q], v[4 * q + 1]), pk2(b4.x, b4.y)), v[4 * q], v[4 * q + 1]);
    upk2(add2(pk2(v[4 * q + 2], v[4 * q + 3]), pk2(b4.z, b4.w)), v[4 * q + 2], v[4 * q + 3]);
  }
  __syncwarp();
}

// Column ownership of the compute warps in the thread-per-row E phases.  The accumulator is produced in two 256-column TMEM
// halves (the two N passes of a 512-wide GEMM), and the MMA warp publishes each half as soon as it is complete, so that
// the E phase of half 0 runs under the MMAs of half 1.  For that every warp owns columns in BOTH halves: with two warps
// per lane quadrant (NCW == 8), warp `part` owns [part * 128, +128) of half 0 and of half 1 -- i.e. heads {part, 2 + part};
// chunk c of a warp (32 columns) lives in half c >> 2.  With four warps per quadrant each warp owns 128 contiguous columns.
__device__ __forceinline__ int ccol(int part, int c) {
  if (NCW == 8) return ((c >> 2) << 8) + part * 128 + ((c & 3) << 5);
  return part * CPW + c * 32;
}

// ---- P phase: 128 rows of h -> (LayerNorm) -> fp16 -> OPA.  Warp per row, the reduction order of ln_rows_kernel.
// Two rows per iteration; the next iteration's rows are requested before the current ones are reduced (there is
// next to no L1 beside 225 KB of shared memory, so every row is an L2 / HBM round trip).
template <bool LN>
__device__ __forceinline__ void rows_to_opa(const float* __restrict__ h, int rows, long long g0, uint32_t opa, int ew,
                                            int lane, const float* __restrict__ lnw, const float* __restrict__ lnb) {
  float4 gw[4], gb[4];
  if (LN) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gw[j] = __ldg(reinterpret_cast<const float4*>(lnw) + j * 32 + lane);
      gb[j] = __ldg(reinterpret_cast<const float4*>(lnb) + j * 32 + lane);
    }
  }
  constexpr int RPI = NCW == 8 ? 2 : 1;      // rows in flight per iteration (register budget: 168 / 96 per thread)
  float4 nx[RPI][4];
  auto fetch = [&](int i) {
#pragma unroll
    for (int u = 0; u < RPI; ++u) {
      const long long g = g0 + ew + NCW * (i + u);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        nx[u][j] = g < rows ? __ldcg(reinterpret_cast<const float4*>(h + g * D) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
#pragma unroll 1
  for (int i = 0; i < ROWS / NCW; i += RPI) {
    float4 x[RPI][4];
#pragma unroll
    for (int u = 0; u < RPI; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) x[u][j] = nx[u][j];
    if (i + RPI < ROWS / NCW) fetch(i + RPI);
#pragma unroll
    for (int u = 0; u < RPI; ++u) {
      const int r = ew + NCW * (i + u);
      float mean = 0.f, rstd = 1.f;
      if (LN) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) s += (x[u][j].x + x[u][j].y) + (x[u][j].z + x[u][j].w);
        mean = warp_sum(s) / (float)D;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = x[u][j].x - mean, b = x[u][j].y - mean, c = x[u][j].z - mean, e = x[u][j].w - mean;
          ss += (a * a + b * b) + (c * c + e * e);
        }
        rstd = rsqrtf(warp_sum(ss) / (float)D + 1e-5f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y0 = x[u][j].x, y1 = x[u][j].y, y2 = x[u][j].z, y3 = x[u][j].w;
        if (LN) {
          y0 = fmaf((y0 - mean) * rstd, gw[j].x, gb[j].x);
          y1 = fmaf((y1 - mean) * rstd, gw[j].y, gb[j].y);
          y2 = fmaf((y2 - mean) * rstd, gw[j].z, gb[j].z);
          y3 = fmaf((y3 - mean) * rstd, gw[j].w, gb[j].w);
        }
        const int col = j * 128 + lane * 4;
        st_shared_v2u(opa_addr(opa, r, col) + (uint32_t)(lane & 1) * 8u, pack_f16x2_sat(y0, y1), pack_f16x2_sat(y2, y3));
      }
    }
  }
}

// The chunk loops of the E phases are deliberately NOT unrolled: the first version of this kernel unrolled them into
// ~300 KB of straight-line SASS that every warp executed once per tile, and ran instruction-fetch bound (10x slower).

// ---- E1: per-head softmax numerator of (acc + bq) over the 128 columns of each of this warp's two heads -> OPA.
// OPA receives the UN-normalised exp(q - max) (in (0, 1], the same relative fp16 rounding as the normalised value);
// the 1 / sum of each head is returned in inv0 / inv1 and applied to the fp32 accumulator of q * ctx by the same thread
// in E2 (y = softmax(q) ctx is linear in the scale of q per head).  Two streaming passes over TMEM, 32 live values.
template <typename WaitHalf>
__device__ __forceinline__ void epi_softmax(uint32_t trow, int part, int row, uint32_t opa, float* bslot, int lane,
                                            const float* __restrict__ bias, float& inv0, float& inv1, WaitHalf&& wait_half) {
#pragma unroll 1
  for (int i = 0; i < CPW / HD; ++i) {
    const int col0 = ccol(part, 4 * i);      // this warp's head in TMEM half i (NCW == 8) / its only head (NCW == 16)
    wait_half(col0 >> 8);
    float m = -INFINITY;
    float bcur = __ldg(bias + col0 + lane);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld_32x32(trow + (uint32_t)(col0 + c * 32), v);
      const float bnxt = __ldg(bias + col0 + ((c + 1) & 3) * 32 + lane);
      tmem_ld_wait();
      add_bias32(bslot, lane, bcur, v);
      bcur = bnxt;
      float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
      for (int j = 4; j < 32; j += 4) {
        m0 = fmaxf(m0, v[j]); m1 = fmaxf(m1, v[j + 1]); m2 = fmaxf(m2, v[j + 2]); m3 = fmaxf(m3, v[j + 3]);
      }
      m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
    }
    const float ml = m * L2E;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    // the numerators overwrite the operand tile, which the MMAs of BOTH halves read (K = all 512 columns): only the max
    // pass above may run under the second half's MMAs
    wait_half(0); wait_half(1);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {       // bcur holds chunk 0's bias again (the prefetch index wraps)
      float v[32];
      tmem_ld_32x32(trow + (uint32_t)(col0 + c * 32), v);
      const float bnxt = __ldg(bias + col0 + ((c + 1) & 3) * 32 + lane);
      tmem_ld_wait();
      add_bias32(bslot, lane, bcur, v);
      bcur = bnxt;
      const f32x2 l2 = pk2(L2E, L2E), nml = pk2(-ml, -ml);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float a0, a1, a2, a3;
        upk2(fma2(pk2(v[j], v[j + 1]), l2, nml), a0, a1);
        upk2(fma2(pk2(v[j + 2], v[j + 3]), l2, nml), a2, a3);
        v[j] = ex2_fast(a0); v[j + 1] = ex2_fast(a1); v[j + 2] = ex2_fast(a2); v[j + 3] = ex2_fast(a3);
        upk2(add2(pk2(s0, s1), pk2(v[j], v[j + 1])), s0, s1);
        upk2(add2(pk2(s2, s3), pk2(v[j + 2], v[j + 3])), s2, s3);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4u(opa_addr(opa, row, col0 + c * 32 + 8 * q), pack_f16x2_sat(v[8 * q], v[8 * q + 1]),
                      pack_f16x2_sat(v[8 * q + 2], v[8 * q + 3]), pack_f16x2_sat(v[8 * q + 4], v[8 * q + 5]),
                      pack_f16x2_sat(v[8 * q + 6], v[8 * q + 7]));
    }
    const float r = 1.f / ((s0 + s1) + (s2 + s3));
    if (i == 0) inv0 = r; else inv1 = r;
  }
}

// ---- E2 / E5: OPA = SiLU(LN_512(acc [* inv] [+ bias]) * gamma' + beta') for the rows with act = true.
// gamma' = w (1 + scale), beta' = b (1 + scale) + shift were staged per sample in shared memory (prm: [gamma' 512 | beta' 512]).
// The two warps of a lane quadrant each own 256 columns and exchange (mean, M2) partial statistics; the statistics
// are accumulated about the row's first element so that one pass suffices without cancellation.
template <bool BIAS, bool SCALE, typename WaitHalf>
__device__ __forceinline__ void epi_lnmod(uint32_t trow, int part, int row, bool act, const float* prm, float* xch, int ew,
                                          int quad, uint32_t opa, float* bslot, int lane, const float* __restrict__ bias,
                                          float inv0, float inv1, WaitHalf&& wait_half) {
  float K = 0.f, sd = 0.f, sq = 0.f;
  float bcur = BIAS ? __ldg(bias + ccol(part, 0) + lane) : 0.f;
  float bcur2 = BIAS ? __ldg(bias + ccol(part, 1) + lane) : 0.f;
  f32x2 sdp = pk2(0.f, 0.f), sqp = pk2(0.f, 0.f), sdp2 = pk2(0.f, 0.f), sqp2 = pk2(0.f, 0.f);   // (tile pairs: two TMEM loads in flight)
#pragma unroll 1
  for (int c = 0; c < NCH; c += 2) {
    float v[32], w[32];
    wait_half(ccol(part, c) >> 8);            // the statistics pass of TMEM half 0 runs under the MMAs of half 1
    tmem_ld_32x32(trow + (uint32_t)ccol(part, c), v);
    tmem_ld_32x32(trow + (uint32_t)ccol(part, c + 1), w);
    const int cn = (c + 2) & (NCH - 1);
    const float bnxt = BIAS ? __ldg(bias + ccol(part, cn) + lane) : 0.f;
    const float bnxt2 = BIAS ? __ldg(bias + ccol(part, cn + 1) + lane) : 0.f;
    tmem_ld_wait();
    if (BIAS) {
      add_bias32(bslot, lane, bcur, v);
      add_bias32(bslot, lane, bcur2, w);
    }
    bcur = bnxt; bcur2 = bnxt2;
    const float sc = SCALE ? ((c >> 2) ? inv1 : inv0) : 1.f;     // a pair of tiles never straddles a head (128 columns)
    if (c == 0) K = v[0] * sc;
    // packed fp32 pairs (FFMA2): the phase is issue-bound; four independent accumulator chains per statistic
    const f32x2 sc2 = pk2(sc, sc), nk2 = pk2(-K, -K);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const f32x2 d = fma2(pk2(v[j], v[j + 1]), sc2, nk2);
      const f32x2 d2 = fma2(pk2(w[j], w[j + 1]), sc2, nk2);
      sdp = add2(sdp, d);
      sqp = fma2(d, d, sqp);
      sdp2 = add2(sdp2, d2);
      sqp2 = fma2(d2, d2, sqp2);
    }
  }
  {
    float a0, a1, b0, b1;
    upk2(add2(sdp, sdp2), a0, a1);
    upk2(add2(sqp, sqp2), b0, b1);
    sd = a0 + a1;
    sq = b0 + b1;
  }
  bcur = BIAS ? __ldg(bias + ccol(part, 0) + lane) : 0.f;
  const float mean_w = fmaf(sd, 1.f / (float)CPW, K);
  const float m2_w = fmaf(-sd * (1.f / (float)CPW), sd, sq);
  xch[(ew * 32 + lane) * 2] = mean_w;
  xch[(ew * 32 + lane) * 2 + 1] = m2_w;
  bar_sync(1 + quad, 32 * WPQ);
  // combine the WPQ partial (mean, M2) pairs of this row (equal counts): Chan et al.
  float msum = 0.f, m2 = 0.f;
  float mj[WPQ];
#pragma unroll
  for (int j = 0; j < WPQ; ++j) {
    mj[j] = xch[(((ew & 3) + 4 * j) * 32 + lane) * 2];
    m2 += xch[(((ew & 3) + 4 * j) * 32 + lane) * 2 + 1];
    msum += mj[j];
  }
  const float mean = msum * (1.f / (float)WPQ);
#pragma unroll
  for (int j = 0; j < WPQ; ++j) m2 = fmaf((mj[j] - mean) * (mj[j] - mean), (float)CPW, m2);
  const float rstd = rsqrtf(m2 * (1.f / (float)D) + 1e-5f);
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {        // bcur holds chunk 0's bias again
    float v[32];
    tmem_ld_32x32(trow + (uint32_t)ccol(part, c), v);
    const float bnxt = BIAS ? __ldg(bias + ccol(part, (c + 1) & (NCH - 1)) + lane) : 0.f;
    tmem_ld_wait();
    if (BIAS) add_bias32(bslot, lane, bcur, v);
    bcur = bnxt;
    // (x sc - mean) rstd = x A + B
    const float A = SCALE ? ((c >> 2) ? inv1 : inv0) * rstd : rstd;
    const float Bc = -mean * rstd;
    const int col = ccol(part, c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 ga = *reinterpret_cast<const float4*>(prm + col + 8 * q);
      const float4 gb = *reinterpret_cast<const float4*>(prm + col + 8 * q + 4);
      const float4 ba = *reinterpret_cast<const float4*>(prm + D + col + 8 * q);
      const float4 bb = *reinterpret_cast<const float4*>(prm + D + col + 8 * q + 4);
      float y[8];
      const f32x2 A2 = pk2(A, A), B2 = pk2(Bc, Bc);
      upk2(silu2(fma2(fma2(pk2(v[8 * q], v[8 * q + 1]), A2, B2), pk2(ga.x, ga.y), pk2(ba.x, ba.y))), y[0], y[1]);
      upk2(silu2(fma2(fma2(pk2(v[8 * q + 2], v[8 * q + 3]), A2, B2), pk2(ga.z, ga.w), pk2(ba.z, ba.w))), y[2], y[3]);
      upk2(silu2(fma2(fma2(pk2(v[8 * q + 4], v[8 * q + 5]), A2, B2), pk2(gb.x, gb.y), pk2(bb.x, bb.y))), y[4], y[5]);
      upk2(silu2(fma2(fma2(pk2(v[8 * q + 6], v[8 * q + 7]), A2, B2), pk2(gb.z, gb.w), pk2(bb.z, bb.w))), y[6], y[7]);
      if (act)
        st_shared_v4u(opa_addr(opa, row, col + 8 * q), pack_f16x2_sat(y[0], y[1]), pack_f16x2_sat(y[2], y[3]),
                      pack_f16x2_sat(y[4], y[5]), pack_f16x2_sat(y[6], y[7]));
    }
  }
}

// ---- E3 / E6: h[rows, 256 columns of this warp] += acc + bias, 32 x 32 fp32 tiles, TMA reduce-add.
// Half 0 (chunks 0..3 of the warp) is reduced while the MMAs of half 1 still READ the operand tile, so its staging tile
// must live elsewhere: `stg1`, one 4 KB slot per warp in the parameter staging region, which is idle in E3 / E6 (the
// caller has synchronised the compute warps).  One tile at a time, each waiting for the previous reduction to have read
// the slot.  Half 1 follows once both halves are complete, two tiles per bulk group through `stg4` = 16 KB of the now idle
// operand tile private to this warp (4 rotating 4 KB slots).
template <typename WaitHalf>
__device__ __forceinline__ void epi_reduce_h(uint32_t trow, int part, uint32_t stg1, uint32_t stg4, const CUtensorMap* map,
                                             int grow0, float* bslot, int lane, const float* __restrict__ bias, bool& pending,
                                             WaitHalf&& wait_half) {
  static_assert(NCW == 8 && NCH == 8 && RED_SLOTS == 4, "the split reduction assumes two warps per lane quadrant");
  wait_half(0);
  {
    float b0 = __ldg(bias + ccol(part, 0) + lane);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld_32x32(trow + (uint32_t)ccol(part, c), v);
      const float n0 = __ldg(bias + ccol(part, (c + 1) & 3) + lane);
      tmem_ld_wait();
      add_bias32(bslot, lane, b0, v);
      b0 = n0;
      if (pending) {                          // the previous reduction (or an older bulk store) has read the slot
        if (lane == 0) tma_wait_read0();
        __syncwarp();
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st_shared_v4(stg1 + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_3d(map, stg1, ccol(part, c), grow0, 0);
        tma_commit();
      }
      pending = true;
    }
  }
  wait_half(1);
  // two 32-column tiles per iteration: both TMEM loads are in flight together and one bulk group carries both reductions
  float b0 = __ldg(bias + ccol(part, 4) + lane), b1 = __ldg(bias + ccol(part, 5) + lane);
#pragma unroll 1
  for (int c = 4; c < NCH; c += 2) {
    float v[32], w[32];
    tmem_ld_32x32(trow + (uint32_t)ccol(part, c), v);
    tmem_ld_32x32(trow + (uint32_t)ccol(part, c + 1), w);
    const int cn = 4 + ((c + 2) & 3);
    const float n0 = __ldg(bias + ccol(part, cn) + lane), n1 = __ldg(bias + ccol(part, cn + 1) + lane);
    tmem_ld_wait();
    add_bias32(bslot, lane, b0, v);
    add_bias32(bslot, lane, b1, w);
    b0 = n0; b1 = n1;
    const uint32_t sa = stg4 + (uint32_t)(c & (RED_SLOTS - 1)) * 4096u, sb = sa + 4096u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      st_shared_v4(sa + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      st_shared_v4(sb + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_reduce_add_3d(map, sa, ccol(part, c), grow0, 0);
      tma_reduce_add_3d(map, sb, ccol(part, c + 1), grow0, 0);
      tma_commit();
    }
    pending = true;
  }
}

__device__ __forceinline__ void stage_params(float* dst, int nsamp, int s_first, int batch, const float* __restrict__ w,
                                             const float* __restrict__ b, const float* __restrict__ scale,
                                             const float* __restrict__ shift, int mod_ld, int ctid) {
  for (int idx = ctid; idx < nsamp * D; idx += NCW * 32) {
    const int sl = idx >> 9, c = idx & (D - 1);
    const int s = min(s_first + sl, batch - 1);
    const float sc = 1.f + __ldg(scale + (size_t)s * mod_ld + c);
    const float sh = __ldg(shift + (size_t)s * mod_ld + c);
    dst[sl * 1024 + c] = __ldg(w + c) * sc;
    dst[sl * 1024 + D + c] = fmaf(__ldg(b + c), sc, sh);
  }
}

__global__ void __launch_bounds__(THREADS, 1)
fused_block_kernel(const __grid_constant__ FbMaps tm, const __grid_constant__ FbParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NSLOT];
  __shared__ __align__(8) uint64_t empty_bar[NSLOT];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t hid_bar;
  __shared__ __align__(8) uint64_t afull_bar[8];    // G5: hidden (A) slabs stream through the idle operand tile
  __shared__ __align__(8) uint64_t aempty_bar[8];
  __shared__ __align__(16) float bias4_s[NCW == 8 ? NCW : 1][32];   // E4's bias broadcast slots (the staging region is full then)
  __shared__ uint32_t tmem_slot;

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t opa = smem_base, ring = smem_base + OPA_BYTES, stg = ring + RING_BYTES;
  uint8_t* const stg_gen = smem_raw + (stg - smem_u32(smem_raw));
  const int rank = (int)cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const int last = p.stop ? p.stop : 7;      // number of compute phases executed per tile (debug truncation)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.wq); tma_prefetch_desc(&tm.ctx); tma_prefetch_desc(&tm.wo); tma_prefetch_desc(&tm.w1);
    tma_prefetch_desc(&tm.hida); tma_prefetch_desc(&tm.w2); tma_prefetch_desc(&tm.wo2);
  }
  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&tm.hred); tma_prefetch_desc(&tm.hidst);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSLOT; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&tfull_bar[s]), 1);
        mbar_init(smem_u32(&tempty_bar[s]), 2u * NCW);       // every compute warp of both CTAs
      }
      mbar_init(smem_u32(&hid_bar), NCW);
      for (int s = 0; s < 8; ++s) {
        mbar_init(smem_u32(&afull_bar[s]), 1);
        mbar_init(smem_u32(&aempty_bar[s]), 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm(smem_u32(&tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;
  // All pairs run the same phase program on same-sized tiles, so they hit the chip-level resources (L2 -> SM weight
  // stream in the MMA phases, L2 reductions in E3 / E6, h reads in the P phases) in lock-step.  Pairs that own one tile
  // fewer than the others have a whole tile time of slack: they start late, which shifts their phases for free.
  if (p.stagger > 0 && p.n_tiles > n_clusters && cluster_id >= p.n_tiles % n_clusters && p.n_tiles % n_clusters != 0 && warp != 1) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)p.stagger) __nanosleep(2000);
  }

  if (warp == 0) {
    // ================================================================== TMA producer (both CTAs)
    int slot = 0, it = 0;
    uint32_t ph = 0;
    auto push = [&](const CUtensorMap* m, int c0, int c1, int c2, uint32_t bytes) {
      mbar_wait(smem_u32(&empty_bar[slot]), ph ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[slot]);
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(bar, 2u * bytes);
        tma_load_3d_2sm(m, bar, ring + (uint32_t)slot * SLAB, c0, c1, c2);
      }
      __syncwarp();
      if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
    };
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++it) {
      const int g0 = tile * 2 * ROWS;
      const int s_first = g0 / p.T;
      const int s_last = (min(g0 + 2 * ROWS, p.rows) - 1) / p.T;
      if (last >= 2)
#pragma unroll 1
        for (int n = 0; n < 2; ++n)
#pragma unroll 1
          for (int kb = 0; kb < D / 64; ++kb) push(&tm.wq, kb * 64, n * 256 + rank * 128, 0, SLAB);
      if (last >= 3)
#pragma unroll 1
        for (int s = s_first; s <= s_last; ++s)
#pragma unroll 1
          for (int hd = 0; hd < H; ++hd)
#pragma unroll 1
            for (int kb = 0; kb < HD / 64; ++kb) push(&tm.ctx, kb * 64, rank * 64, s * H + hd, SLAB / 2);
      if (last >= 4)
#pragma unroll 1
        for (int n = 0; n < 2; ++n)
#pragma unroll 1
          for (int kb = 0; kb < D / 64; ++kb) push(&tm.wo, kb * 64, n * 256 + rank * 128, 0, SLAB);
      if (last >= 5)
#pragma unroll 1
        for (int q = 0; q < 4; ++q)
#pragma unroll 1
          for (int kb = 0; kb < D / 64; ++kb) push(&tm.w1, kb * 64, q * 256 + rank * 128, 0, SLAB);
      if (last >= 6) {
        mbar_wait(smem_u32(&hid_bar), (uint32_t)it & 1u);      // this CTA's hidden rows are in global memory
        fence_proxy_async_all();
        // the operand tile is idle between G4 and E5: the 16 hidden slabs stream through it as an 8-slot ring
        // (each slab feeds both N halves), the W2 slabs through the ordinary ring
#pragma unroll 1
        for (int kb = 0; kb < F / 64; ++kb) {
          const int sa = kb & 7;
          mbar_wait(smem_u32(&aempty_bar[sa]), (uint32_t)((kb >> 3) ^ 1));
          const uint32_t abar = smem_u32(&afull_bar[sa]);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(abar, 2u * SLAB);
            tma_load_3d_2sm(&tm.hida, abar, opa + (uint32_t)sa * SLAB, kb * 64, (int)blockIdx.x * ROWS, 0);
          }
          __syncwarp();
          push(&tm.w2, kb * 64, rank * 128, 0, SLAB);
          push(&tm.w2, kb * 64, 256 + rank * 128, 0, SLAB);
        }
      }
      if (last >= 7)
#pragma unroll 1
        for (int n = 0; n < 2; ++n)
#pragma unroll 1
          for (int kb = 0; kb < D / 64; ++kb) push(&tm.wo2, kb * 64, n * 256 + rank * 128, 0, SLAB);
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      int slot = 0;
      uint32_t ph = 0;
      uint32_t te_par[2] = {0u, 0u};
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      // instruction descriptors: fp32 accumulate, fp16 A/B, K-major, N >> 3, M = 256 >> 4
      const uint32_t idesc256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const bool prof = p.prof != nullptr;
      long long m_te = 0, m_full = 0;
      const long long m_t0 = prof ? clock64() : 0;
      auto wait_te = [&](int hh) {
        const long long t0 = prof ? clock64() : 0;
        mbar_wait(smem_u32(&tempty_bar[hh]), te_par[hh]);
        if (prof) m_te += clock64() - t0;
        te_par[hh] ^= 1u;
        tc_fence_after();
      };
      auto commit_tf = [&](int hh) {
        if (elect_one()) umma_commit_2sm(smem_u32(&tfull_bar[hh]), (uint16_t)3);
        __syncwarp();
      };
      // TWO consecutive ring slabs per trip.  The serial chain of one slab -- barrier probe (~295 cycles through the barrier unit
      // even when the phase completed long ago), fence, four MMA issues, commit, reconvergence -- measured ~700 cycles against
      // 512 cycles of tensor time for the four 256 x 256 x 16 pair MMAs: the G phases were ISSUE-bound (116 slabs x 700 = the
      // 80 k cycles per tile of the phase clocks; the 34 k "waiting for ring slabs" was the probe latency).  Here both probes are
      // in flight together, descriptors are formed by adding to the slab-0 descriptor (a swizzle-128B descriptor's low word is
      // address >> 4), and the warp reconverges once per trip.  Every caller consumes an even number of slabs starting at an
      // even slot, so a trip never wraps the ring.
      const uint64_t d_ring0 = make_smem_desc_sw128(ring), d_opa0 = make_smem_desc_sw128(opa);
      auto step2 = [&](uint32_t a_off0, uint32_t a_off1, uint32_t dcol0, uint32_t dcol1, uint32_t idesc, bool fresh0, bool fresh1) {
        const long long t0 = prof ? clock64() : 0;
        uint32_t r0, r1;
        mbar_test2(full0 + (uint32_t)slot * 8u, ph, full0 + (uint32_t)slot * 8u + 8u, ph, r0, r1);
        if (!r0) mbar_wait(full0 + (uint32_t)slot * 8u, ph);
        if (prof) m_full += clock64() - t0;
        tc_fence_after();
        const uint64_t dB0 = d_ring0 + (uint64_t)((uint32_t)slot * (SLAB >> 4));
        if (elect_one()) {
          const uint64_t dA = d_opa0 + (uint64_t)(a_off0 >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_2sm(tmem_base + dcol0, dA + 2 * k, dB0 + 2 * k, idesc, (fresh0 && k == 0) ? 0u : 1u);
          umma_commit_2sm(empty0 + (uint32_t)slot * 8u, (uint16_t)3);
        }
        if (!r1) {
          const long long t1 = prof ? clock64() : 0;
          mbar_wait(full0 + (uint32_t)slot * 8u + 8u, ph);
          if (prof) m_full += clock64() - t1;
          tc_fence_after();
        }
        if (elect_one()) {
          const uint64_t dA = d_opa0 + (uint64_t)(a_off1 >> 4), dB1 = dB0 + (uint64_t)(SLAB >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_2sm(tmem_base + dcol1, dA + 2 * k, dB1 + 2 * k, idesc, (fresh1 && k == 0) ? 0u : 1u);
          umma_commit_2sm(empty0 + (uint32_t)slot * 8u + 8u, (uint16_t)3);
        }
        __syncwarp();
        slot += 2;
        if (slot == NSLOT) { slot = 0; ph ^= 1u; }
      };
      // a full-width GEMM out of the resident operand tile: D[128(x2) x 512] = OPA[.. x 512] W[512 x 512]^T
      // (each 256-column half is published as soon as its MMAs are issued: the E phase of half 0 overlaps half 1)
      auto gemm_opa_512 = [&]() {
#pragma unroll 1
        for (int n = 0; n < 2; ++n) {
#pragma unroll 1
          for (int kb = 0; kb < D / 64; kb += 2)
            step2((uint32_t)kb * SLAB, (uint32_t)(kb + 1) * SLAB, (uint32_t)(n * 256), (uint32_t)(n * 256), idesc256, kb == 0, false);
          commit_tf(n);
        }
      };
#pragma unroll 1
      for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
        const int g0 = tile * 2 * ROWS;
        const int s_first = g0 / p.T;
        const int s_last = (min(g0 + 2 * ROWS, p.rows) - 1) / p.T;
        if (last >= 2) {                                     // G1
          wait_te(0); wait_te(1);
          gemm_opa_512();
        }
        if (last >= 3) {                                     // G2, one round per sample
#pragma unroll 1
          for (int s = s_first; s <= s_last; ++s) {
            wait_te(0); wait_te(1);
#pragma unroll 1
            for (int hd = 0; hd < H; ++hd) {
              static_assert(HD / 64 == 2, "G2 consumes one slab pair per head");
              step2((uint32_t)(hd * 2) * SLAB, (uint32_t)(hd * 2 + 1) * SLAB, (uint32_t)(hd * HD), (uint32_t)(hd * HD), idesc128, true, false);
              if (hd & 1) commit_tf(hd >> 1);                 // heads 0, 1 = TMEM half 0; heads 2, 3 = half 1
            }
          }
        }
        if (last >= 4) {                                     // G3
          wait_te(0); wait_te(1);
          gemm_opa_512();
        }
        if (last >= 5) {                                     // G4: four 256-column quarters, alternating TMEM halves
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            wait_te(q & 1);
#pragma unroll 1
            for (int kb = 0; kb < D / 64; kb += 2)
              step2((uint32_t)kb * SLAB, (uint32_t)(kb + 1) * SLAB, (uint32_t)((q & 1) * 256), (uint32_t)((q & 1) * 256), idesc256, kb == 0, false);
            commit_tf(q & 1);
          }
        }
        if (last >= 6) {                                     // G5: hidden slabs (A) from the operand-tile ring, W2 from the ring
          wait_te(0); wait_te(1);
#pragma unroll 1
          for (int kb = 0; kb < F / 64; ++kb) {
            const int sa = kb & 7;
            const long long t0 = prof ? clock64() : 0;
            mbar_wait(smem_u32(&afull_bar[sa]), (uint32_t)(kb >> 3));
            if (prof) m_full += clock64() - t0;
            step2((uint32_t)sa * SLAB, (uint32_t)sa * SLAB, 0u, 256u, idesc256, kb == 0, kb == 0);
            if (elect_one()) umma_commit_2sm(smem_u32(&aempty_bar[sa]), (uint16_t)3);
            __syncwarp();
          }
          commit_tf(0); commit_tf(1);
        }
        if (last >= 7) {                                     // G6
          wait_te(0); wait_te(1);
          gemm_opa_512();
        }
      }
      if (prof && lane == 0) {
        atomicAdd(p.prof + 16, (unsigned long long)m_te);
        atomicAdd(p.prof + 17, (unsigned long long)m_full);
        atomicAdd(p.prof + 18, (unsigned long long)(clock64() - m_t0));
      }
    }
  } else {
    // ================================================================== compute warps
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int part = ew >> 2;                  // which CPW-column slice of the accumulator this warp owns
    const int row = quad * 32 + lane;          // this thread's row in thread-per-row phases
    const int ctid = ew * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t stgw = stg + (uint32_t)ew * STG_PER_WARP;
    float* const bslot = reinterpret_cast<float*>(stg_gen + BIAS_OFF) + ew * 32;
    float* const prm_all = reinterpret_cast<float*>(stg_gen);
    float* const xch = reinterpret_cast<float*>(stg_gen + XCH_OFF);
    uint32_t tf_par[2] = {0u, 0u};
    bool pending = false;                      // a bulk store / reduction of this warp may still read its staging slot
    auto wait_tf = [&](int hh) {
      mbar_wait(smem_u32(&tfull_bar[hh]), tf_par[hh]);
      tf_par[hh] ^= 1u;
      tc_fence_after();
    };
    // a phase consumes each TMEM half exactly once, whichever chunk touches it first; `finish` closes the phase (a warp
    // that skipped its epilogue still has to follow the barrier phases)
    bool got_half[2] = {false, false};
    auto wait_half = [&](int hh) {
      if (!got_half[hh]) {
        wait_tf(hh);
        got_half[hh] = true;
      }
    };
    auto finish_halves = [&]() {
      wait_half(0); wait_half(1);
      got_half[0] = got_half[1] = false;
    };
    // operand tile written / accumulator drained: tell the leader's MMA warp
    auto arrive_te = [&](bool h0, bool h1) {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (h0) mbar_arrive_remote(smem_u32(&tempty_bar[0]), 0u);
        if (h1) mbar_arrive_remote(smem_u32(&tempty_bar[1]), 0u);
      }
    };
    auto drain_stores = [&](bool full) {       // staging slot reusable (reads done) or stores globally complete
      if (pending) {
        if (lane == 0) { if (full) tma_wait_all0(); else tma_wait_read0(); }
        __syncwarp();
        pending = false;
      }
    };

    long long tph[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) tph[i] = 0;
    const bool prof = p.prof != nullptr;
    long long tlast = prof ? clock64() : 0;
#define FB_TICK(i) do { if (prof) { const long long _n = clock64(); tph[i] += _n - tlast; tlast = _n; } } while (0)
    int it = 0;
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++it) {
      const int g0 = tile * 2 * ROWS;                       // first row of the pair's tile
      const int gc = g0 + rank * ROWS;                      // first row of this CTA
      const int s_first = g0 / p.T;
      const int s_last = (min(g0 + 2 * ROWS, p.rows) - 1) / p.T;
      const int nsamp = s_last - s_first + 1;
      const int grow = gc + row;
      const int s_row = grow / p.T;
      const int slot_row = min(max(s_row - s_first, 0), nsamp - 1);
      int done = 0;                                          // compute phases finished for this tile
      float qinv0 = 1.f, qinv1 = 1.f;                        // 1 / softmax denominators of this thread's two heads

      // ---- P0
      drain_stores(false);
      bar_sync(5, NCW * 32);                                 // every warp's E6 reductions have read their slice of the operand tile
      rows_to_opa<true>(p.h, p.rows, gc, opa, ew, lane, p.ca_ln_w, p.ca_ln_b);
      if (++done < last) arrive_te(true, true);
      {
        // pull the NEXT tile's rows of h (one contiguous 256 KB block per CTA) into L2 while this tile computes:
        // P0 is otherwise an HBM round trip per row pair, and all CTAs issue theirs at the same moment
        const long long gn = (long long)(tile + n_clusters) * 2 * ROWS + rank * ROWS;
        if (tile + n_clusters < p.n_tiles && lane == 0) {
          const long long r0 = gn + ew * (ROWS / NCW);
          const long long nrow = min((long long)(ROWS / NCW), (long long)p.rows - r0);
          if (nrow > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.h + r0 * D), "r"((uint32_t)(nrow * D * 4)) : "memory");
        }
      }
      if (it == 0) FB_TICK(15 + 4); else FB_TICK(0);
      if (done < last) {
        // ---- E1
        FB_TICK(1);
        epi_softmax(trow, part, row, opa, bslot, lane, p.ca_bq, qinv0, qinv1, wait_half);
        finish_halves();
        if (++done < last) arrive_te(true, true);
        FB_TICK(2);
      }
      if (done < last) {
        // ---- E2 (one round per sample of the tile)
        stage_params(prm_all, nsamp, s_first, p.batch, p.ca_pn_w, p.ca_pn_b, p.ca_scale, p.ca_shift, p.mod_ld, ctid);
        bar_sync(5, NCW * 32);
        ++done;
        for (int s = s_first; s <= s_last; ++s) {
          FB_TICK(3);
          const bool act = (s_row == s) && (grow < p.rows);
          if (__any_sync(0xffffffffu, act))
            epi_lnmod<false, true>(trow, part, row, act, prm_all + (s - s_first) * 1024, xch, ew, quad, opa, bslot, lane, nullptr, qinv0, qinv1, wait_half);
          finish_halves();
          if (s < s_last || done < last) arrive_te(true, true);
          FB_TICK(4);
        }
      }
      if (done < last) {
        // ---- E3: h += d + bo, then OPA = fp16(h)
        FB_TICK(5);
        bar_sync(5, NCW * 32);                               // every warp is past E2: the parameter staging region is free
        epi_reduce_h(trow, part, stg + (uint32_t)ew * 4096u, opa + (uint32_t)ew * (OPA_BYTES / NCW), &tm.hred, gc + quad * 32,
                     bias4_s[ew], lane, p.ca_bo, pending, wait_half);
        finish_halves();
        tc_fence_before();
        FB_TICK(6);
        drain_stores(true);
        bar_sync(5, NCW * 32);                               // every reduction into this CTA's rows has been performed
        FB_TICK(7);
        rows_to_opa<false>(p.h, p.rows, gc, opa, ew, lane, nullptr, nullptr);
        if (++done < last) arrive_te(true, true);
        FB_TICK(8);
      }
      if (done < last) {
        // ---- E4: hidden = GELU(u + b1) -> fp16 -> hidden scratch (this CTA's private rows)
        ++done;
        constexpr int QW = 256 / WPQ;          // columns of one quarter owned by this warp
        constexpr int E4_SLOTS = STG_PER_WARP / 2048;
        float bcur = __ldg(p.f_b1 + part * QW + lane);
#pragma unroll 1
        for (int idx = 0; idx < 4 * QCH; ++idx) {
          const int q = idx / QCH, c = idx % QCH;
          if (c == 0) {
            wait_tf(q & 1);
            FB_TICK(9);
          }
          float v[32];
          const long long e4a = prof ? clock64() : 0;
          tmem_ld_32x32(trow + (uint32_t)((q & 1) * 256 + part * QW + c * 32), v);
          const int nidx = (idx + 1) % (4 * QCH);
          const float bnxt = __ldg(p.f_b1 + (nidx / QCH) * 256 + part * QW + (nidx % QCH) * 32 + lane);
          tmem_ld_wait();
          const long long e4b = prof ? clock64() : 0;
          // the staging region belongs to the hidden-tile stores in this phase: bias goes through a static slot
          if (NCW == 8) {
            add_bias32(bias4_s[NCW == 8 ? ew : 0], lane, bcur, v);
#pragma unroll
            for (int j = 0; j < 32; j += 2) gelu_relu_erfc2(v[j], v[j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_relu_erfc(v[j] + __shfl_sync(0xffffffffu, bcur, j));
          }
          bcur = bnxt;
          const long long e4c = prof ? clock64() : 0;
          const uint32_t sl = stgw + (uint32_t)(idx & (E4_SLOTS - 1)) * 2048u;
          if (idx >= E4_SLOTS) {
            if (lane == 0) { if (E4_SLOTS == 2) tma_wait_read1(); else tma_wait_read0(); }
            __syncwarp();
          }
          const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            st_shared_v4u(sl + (uint32_t)lane * 64u + (((uint32_t)qq ^ sw) << 4), pack_f16x2_sat(v[8 * qq], v[8 * qq + 1]),
                          pack_f16x2_sat(v[8 * qq + 2], v[8 * qq + 3]), pack_f16x2_sat(v[8 * qq + 4], v[8 * qq + 5]),
                          pack_f16x2_sat(v[8 * qq + 6], v[8 * qq + 7]));
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tm.hidst, sl, q * 256 + part * QW + c * 32, (int)blockIdx.x * ROWS + quad * 32, 0);
            tma_commit();
          }
          if (prof) {
            const long long e4d = clock64();
            tph[16] += e4b - e4a; tph[17] += e4c - e4b; tph[18] += e4d - e4c;
          }
          if (c == QCH - 1) {
            if (q < 2 || done < last) arrive_te((q & 1) == 0, (q & 1) == 1);
            FB_TICK(10);
          }
        }
        pending = true;
        tc_fence_before();
        drain_stores(true);                                  // hidden rows are in global memory
        if (done < last && lane == 0) mbar_arrive(smem_u32(&hid_bar));
        FB_TICK(11);
      }
      if (done < last) {
        // ---- E5
        bar_sync(5, NCW * 32);                               // all warps' staging slots are free
        stage_params(prm_all, nsamp, s_first, p.batch, p.f_pn_w, p.f_pn_b, p.f_scale, p.f_shift, p.mod_ld, ctid);
        bar_sync(5, NCW * 32);
        FB_TICK(12);
        epi_lnmod<true, false>(trow, part, row, true, prm_all + slot_row * 1024, xch, ew, quad, opa, bslot, lane, p.f_b2, 1.f, 1.f, wait_half);
        finish_halves();
        if (++done < last) arrive_te(true, true);
        FB_TICK(13);
      }
      if (done < last) {
        // ---- E6
        bar_sync(5, NCW * 32);                               // nobody still reads the staged parameters / statistics
        FB_TICK(14);
        epi_reduce_h(trow, part, stg + (uint32_t)ew * 4096u, opa + (uint32_t)ew * (OPA_BYTES / NCW), &tm.hred, gc + quad * 32,
                     bias4_s[ew], lane, p.f_bo, pending, wait_half);
        finish_halves();
        tc_fence_before();
        ++done;
        FB_TICK(15);
      }
      if (p.stop != 0 && p.dbg != nullptr) {
        // debug: dump the operand tile (un-swizzled) after the last executed phase
        fence_async_smem();
        bar_sync(5, NCW * 32);
        for (int idx = ctid; idx < ROWS * 64; idx += NCW * 32) {
          const int r = idx >> 6, col = (idx & 63) * 8;
          if (gc + r < p.rows) {
            const float4 w4 = ld_shared_v4(opa_addr(opa, r, col));
            *reinterpret_cast<float4*>(p.dbg + (size_t)(gc + r) * D + col) = w4;
          }
        }
        bar_sync(5, NCW * 32);
      }
    }
    drain_stores(true);                                      // bulk operations complete before the CTA exits
    if (prof && lane == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) atomicAdd(p.prof + i, (unsigned long long)tph[i]);
      atomicAdd(p.prof + 19, (unsigned long long)tph[19]);
#pragma unroll
      for (int i = 0; i < 3; ++i) atomicAdd(p.prof + 20 + i, (unsigned long long)tph[16 + i]);
    }
#undef FB_TICK
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // the peer may still arrive on / multicast into this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// =================================================================================================================
// sa_tail_kernel -- the tail of the channel attention (EfficientSelfAttention on x^T, efficient_attention.py:40-45 with
// stylization_block.py:29-40, called from mcm.py:28-32) for 256 channel-tokens (half a sample) per CTA pair:
//   G_a  y = softmax(q) ctx[b]            (A: 256 rows of the fp16 softmax(q) operand, resident in the operand tile;
//                                          B: this sample's block-diagonal context, transposed)
//   E_a  OPA = SiLU(LN_T(y) (1 + scale_b) + shift_b)          thread = channel row, statistics over the T features
//   G_b  d = OPA Wo^T
//   E_b  h[b, t, n] += d[n, t] + bo[t]    transposed 32 x 32 tiles, TMA reduce-add into the [B, T, 512] residual stream
// G_a of the next tile runs while E_b of this one drains (different TMEM halves, A slabs are re-loaded after G_b).
// =================================================================================================================
struct StMaps {
  CUtensorMap qs, ctx, wo, hred;
};
struct StParams {
  int T, Np, nkb, nch, batch, n_tiles, mod_ld;
  const float *pn_w, *pn_b, *scale, *shift, *bo;
  unsigned long long* prof;   // MCM_FUSED_PROF=1 MCM_ST_PROF=1: summed cycles per phase of the compute warps
  const float* h;             // residual stream (L2 prefetch of the tile E_b reduces into)
  int pf;
};

__global__ void __launch_bounds__(THREADS, 1)
sa_tail_kernel(const __grid_constant__ StMaps tm, const __grid_constant__ StParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NSLOT];
  __shared__ __align__(8) uint64_t empty_bar[NSLOT];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t afull_bar, aempty_bar;
  __shared__ uint32_t tmem_slot;

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t opa = smem_base, ring = smem_base + OPA_BYTES, stg = ring + RING_BYTES;
  uint8_t* const stg_gen = smem_raw + (stg - smem_u32(smem_raw));
  const int rank = (int)cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const uint32_t slab_b = (uint32_t)(p.Np / 2) * 128u;       // bytes of this CTA's half of a B slab

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.qs); tma_prefetch_desc(&tm.ctx); tma_prefetch_desc(&tm.wo); tma_prefetch_desc(&tm.hred);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSLOT; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&tfull_bar[s]), 1);
        mbar_init(smem_u32(&tempty_bar[s]), 2u * NCW);
      }
      mbar_init(smem_u32(&afull_bar), 1);
      mbar_init(smem_u32(&aempty_bar), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm(smem_u32(&tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    int slot = 0, it = 0;
    uint32_t ph = 0;
    auto push = [&](const CUtensorMap* m, int c0, int c1, int c2, uint32_t bytes) {
      mbar_wait(smem_u32(&empty_bar[slot]), ph ^ 1u);
      const uint32_t bar = smem_u32(&full_bar[slot]);
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(bar, 2u * bytes);
        tma_load_3d_2sm(m, bar, ring + (uint32_t)slot * SLAB, c0, c1, c2);
      }
      __syncwarp();
      if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
    };
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++it) {
      const int b = tile >> 1;
      const int r0 = tile * 2 * ROWS + rank * ROWS;            // first row of this CTA in the [B*512, T] token space
      if (p.pf && elect_one()) {
        // E_b reduce-adds into h[b, :, n0 .. n0+128): a cold line makes the L2 reduction wait for HBM, so pull the tile into
        // L2 while G_a / E_a / G_b run (this warp is otherwise idle here)
        const float* src = p.h + (size_t)b * p.T * D + (tile & 1) * 2 * ROWS + rank * ROWS;
        for (int t = 0; t < p.T; ++t)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;" ::"l"(src + (size_t)t * D) : "memory");
      }
      __syncwarp();
      mbar_wait(smem_u32(&aempty_bar), ((uint32_t)it & 1u) ^ 1u);   // G_b of the previous tile has read the operand tile
      const uint32_t abar = smem_u32(&afull_bar);
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(abar, 2u * (uint32_t)p.nkb * SLAB);
        for (int kb = 0; kb < p.nkb; ++kb) tma_load_3d_2sm(&tm.qs, abar, opa + (uint32_t)kb * SLAB, kb * 64, r0, 0);
      }
      __syncwarp();
#pragma unroll 1
      for (int kb = 0; kb < p.nkb; ++kb) push(&tm.ctx, kb * 64, rank * (p.Np / 2), b, slab_b);
#pragma unroll 1
      for (int kb = 0; kb < p.nkb; ++kb) push(&tm.wo, kb * 64, rank * (p.Np / 2), 0, slab_b);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader)
    if (rank == 0) {
      int slot = 0, it = 0;
      uint32_t ph = 0;
      uint32_t te_par[2] = {0u, 0u};
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      auto step = [&](uint32_t a_addr, uint32_t dcol, bool fresh) {
        mbar_wait(smem_u32(&full_bar[slot]), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_addr = ring + (uint32_t)slot * SLAB;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_2sm(tmem_base + dcol, make_smem_desc_sw128(a_addr + k * 32), make_smem_desc_sw128(b_addr + k * 32), idesc,
                         (fresh && k == 0) ? 0u : 1u);
          umma_commit_2sm(smem_u32(&empty_bar[slot]), (uint16_t)3);
        }
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
      };
#pragma unroll 1
      for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++it) {
        // G_a: needs the A slabs and TMEM half 0 drained (E_a of the previous tile, already awaited by its G_b)
        mbar_wait(smem_u32(&afull_bar), (uint32_t)it & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < p.nkb; ++kb) step(opa + (uint32_t)kb * SLAB, 0u, kb == 0);
        if (elect_one()) umma_commit_2sm(smem_u32(&tfull_bar[0]), (uint16_t)3);
        __syncwarp();
        // G_b: operand tile rewritten by E_a (tempty 0), TMEM half 1 drained by E_b of the previous tile (tempty 1)
        mbar_wait(smem_u32(&tempty_bar[0]), te_par[0]); te_par[0] ^= 1u;
        mbar_wait(smem_u32(&tempty_bar[1]), te_par[1]); te_par[1] ^= 1u;
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < p.nkb; ++kb) step(opa + (uint32_t)kb * SLAB, 256u, kb == 0);
        if (elect_one()) {
          umma_commit_2sm(smem_u32(&tfull_bar[1]), (uint16_t)3);
          umma_commit_2sm(smem_u32(&aempty_bar), (uint16_t)3);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int part = ew >> 2;                                 // NCW == 8: two warps per lane quadrant
    const int row = quad * 32 + lane;
    const int ctid = ew * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* const prm = reinterpret_cast<float*>(stg_gen);              // gamma'[Tq] | beta'[Tq] | bo[Tq]   (Tq = nch * 32)
    float* const xch = reinterpret_cast<float*>(stg_gen + 16384);      // [NCW][32][2]
    float* const bslot = reinterpret_cast<float*>(stg_gen + 16384 + 4096) + ew * 32;
    const uint32_t stg2 = opa + 4u * SLAB + (uint32_t)ew * 8192u;      // two 4 KB transposed staging tiles in the unused slabs
    const int Tq = p.nch * 32;
    const int c_split = (p.nch + 1) / 2;                      // part 0: chunks [0, c_split), part 1: [c_split, nch)
    const int c_lo = part == 0 ? 0 : c_split, c_hi = part == 0 ? c_split : p.nch;
    const float n_mine = (float)(min(p.T, c_hi * 32) - c_lo * 32);
    uint32_t tf_par[2] = {0u, 0u};
    bool pending = false;
    int nstore = 0;
    auto arrive = [&](int hh) {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(smem_u32(&tempty_bar[hh]), 0u);
    };
    arrive(1);                                                // TMEM half 1 starts out free
    int prev_b = -1;
    long long tph[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) tph[i] = 0;
    const bool prof = p.prof != nullptr;
    long long tlast = prof ? clock64() : 0;
#define ST_TICK(i) do { if (prof) { const long long _n = clock64(); tph[i] += _n - tlast; tlast = _n; } } while (0)
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
      const int b = tile >> 1;
      const int n0 = (tile & 1) * 2 * ROWS + rank * ROWS;     // first channel of this CTA
      if (b != prev_b) {
        // stage this sample's folded AdaLN parameters over the T features (everybody is past E_a of the previous tile:
        // the barrier below is also crossed only after every warp arrived on tempty 0)
        bar_sync(5, NCW * 32);
        for (int t = ctid; t < Tq; t += NCW * 32) {
          float g = 0.f, be = 0.f, bo = 0.f;
          if (t < p.T) {
            const float sc = 1.f + __ldg(p.scale + (size_t)b * p.mod_ld + t);
            g = __ldg(p.pn_w + t) * sc;
            be = fmaf(__ldg(p.pn_b + t), sc, __ldg(p.shift + (size_t)b * p.mod_ld + t));
            bo = __ldg(p.bo + t);
          }
          prm[t] = g; prm[Tq + t] = be; prm[2 * Tq + t] = bo;
        }
        bar_sync(5, NCW * 32);
        prev_b = b;
      }
      ST_TICK(0);
      // ---- E_a
      mbar_wait(smem_u32(&tfull_bar[0]), tf_par[0]); tf_par[0] ^= 1u;
      tc_fence_after();
      ST_TICK(1);
      {
        float K = 0.f, sd = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32];
          tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          if (c == c_lo) K = v[0];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = (c * 32 + j < p.T) ? v[j] - K : 0.f;    // columns >= T hold stale TMEM contents
            sd += d;
            sq = fmaf(d, d, sq);
          }
        }
        const float mean_w = K + sd / n_mine;
        const float m2_w = sq - sd * sd / n_mine;
        xch[(ew * 32 + lane) * 2] = mean_w;
        xch[(ew * 32 + lane) * 2 + 1] = m2_w;
        bar_sync(1 + quad, 64);
        const float mean_o = xch[((ew ^ 4) * 32 + lane) * 2];
        const float m2_o = xch[((ew ^ 4) * 32 + lane) * 2 + 1];
        const float n_o = (float)p.T - n_mine;
        const float delta = mean_o - mean_w;
        const float mean = mean_w + delta * (n_o / (float)p.T);
        const float m2 = (m2_w + m2_o) + delta * delta * (n_mine * n_o / (float)p.T);
        const float rstd = rsqrtf(m2 / (float)p.T + 1e-5f);
        const float Bc = -mean * rstd;
        ST_TICK(2);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32];
          tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
          tmem_ld_wait();
          const int col = c * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ga = *reinterpret_cast<const float4*>(prm + col + 8 * q);
            const float4 gb = *reinterpret_cast<const float4*>(prm + col + 8 * q + 4);
            const float4 ba = *reinterpret_cast<const float4*>(prm + Tq + col + 8 * q);
            const float4 bb = *reinterpret_cast<const float4*>(prm + Tq + col + 8 * q + 4);
            float y[8];
            const f32x2 A2 = pk2(rstd, rstd), B2 = pk2(Bc, Bc);       // packed fp32 pairs (FFMA2): the phase is issue-bound
            upk2(silu2(fma2(fma2(pk2(v[8 * q], v[8 * q + 1]), A2, B2), pk2(ga.x, ga.y), pk2(ba.x, ba.y))), y[0], y[1]);
            upk2(silu2(fma2(fma2(pk2(v[8 * q + 2], v[8 * q + 3]), A2, B2), pk2(ga.z, ga.w), pk2(ba.z, ba.w))), y[2], y[3]);
            upk2(silu2(fma2(fma2(pk2(v[8 * q + 4], v[8 * q + 5]), A2, B2), pk2(gb.x, gb.y), pk2(bb.x, bb.y))), y[4], y[5]);
            upk2(silu2(fma2(fma2(pk2(v[8 * q + 6], v[8 * q + 7]), A2, B2), pk2(gb.z, gb.w), pk2(bb.z, bb.w))), y[6], y[7]);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (col + 8 * q + e >= p.T) y[e] = 0.f;           // K padding of the next GEMM must be exactly zero
            st_shared_v4u(opa_addr(opa, row, col + 8 * q), pack_f16x2_sat(y[0], y[1]), pack_f16x2_sat(y[2], y[3]),
                          pack_f16x2_sat(y[4], y[5]), pack_f16x2_sat(y[6], y[7]));
          }
        }
      }
      arrive(0);
      ST_TICK(3);
      // ---- E_b
      mbar_wait(smem_u32(&tfull_bar[1]), tf_par[1]); tf_par[1] ^= 1u;
      tc_fence_after();
      ST_TICK(4);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        float v[32];
        tmem_ld_32x32(trow + (uint32_t)(256 + c * 32), v);
        tmem_ld_wait();
        const uint32_t sl = stg2 + (uint32_t)(nstore & 1) * 4096u;
        if (nstore >= 2) {
          if (lane == 0) tma_wait_read1();
          __syncwarp();
        }
        // transposed tile: [t][n], lanes (= channels n) contiguous; the output bias bo[t] is a broadcast read
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b4 = *reinterpret_cast<const float4*>(prm + 2 * Tq + c * 32 + 4 * q4);
          const float o[4] = {v[4 * q4] + b4.x, v[4 * q4 + 1] + b4.y, v[4 * q4 + 2] + b4.z, v[4 * q4 + 3] + b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sl + (uint32_t)((4 * q4 + e) * 32 + lane) * 4u), "f"(o[e]) : "memory");
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_3d(&tm.hred, sl, n0 + quad * 32, c * 32, b);
          tma_commit();
        }
        ++nstore;
        pending = true;
      }
      arrive(1);
      ST_TICK(5);
    }
    if (pending) {
      if (lane == 0) tma_wait_all0();
      __syncwarp();
    }
    ST_TICK(6);
    if (prof && lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(p.prof + 8 * part + i, (unsigned long long)tph[i]);
    }
#undef ST_TICK
    (void)bslot;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}


// =================================================================================================================
// sa_front_kernel -- the head of the channel attention (EfficientSelfAttention on x^T, efficient_attention.py:29-36,
// called from mcm.py:28-32) for 256 channel-tokens (half a sample) per CTA pair:
//   P    OPA = LN_T(h^T)        thread = channel, statistics over the T frames read straight out of h[b, :, n] (coalesced
//                               across channels), so the transposed tensor is never materialised
//   G_q  q = OPA Wq^T    G_k  k = OPA Wk^T    (TMEM halves 0 / 1)      G_v  v = OPA Wv^T (half 0 again, after E_q)
//   E_q  softmax over each head's T/H features -> fp16 rows -> one TMA store per warp       (warps 0..3 of the 8)
//   E_k  k + bk -> fp32, stored TRANSPOSED into [B, T, 512] (the token softmax and k^T v stay separate kernels: they
//        reduce over all 512 channels of a sample, i.e. across tiles)                        (warps 4..7)
//   E_v  v + bv -> fp16, stored transposed into [B, T, 512]
// =================================================================================================================
struct SfMaps {
  CUtensorMap w, qs, k32, v16;
};
struct SfParams {
  const float* h;
  int T, Tp, Np, nkb, nch, hd, batch, n_tiles;
  const float *ln_w, *ln_b, *bqkv;
  unsigned long long* prof;
};

__global__ void __launch_bounds__(THREADS, 1)
sa_front_kernel(const __grid_constant__ SfMaps tm, const __grid_constant__ SfParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NSLOT];
  __shared__ __align__(8) uint64_t empty_bar[NSLOT];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t pa_bar, qd_bar;           // operand tile written (16 warps) / q accumulator drained (8 warps)
  __shared__ uint32_t tmem_slot;

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t opa = smem_base, ring = smem_base + OPA_BYTES, stg = ring + RING_BYTES;
  uint8_t* const stg_gen = smem_raw + (stg - smem_u32(smem_raw));
  uint8_t* const opa_gen = smem_raw + (opa - smem_u32(smem_raw));
  const int rank = (int)cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const uint32_t slab_b = (uint32_t)(p.Np / 2) * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.w); tma_prefetch_desc(&tm.qs); tma_prefetch_desc(&tm.k32); tma_prefetch_desc(&tm.v16);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSLOT; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      mbar_init(smem_u32(&tfull_bar[0]), 1);
      mbar_init(smem_u32(&tfull_bar[1]), 1);
      mbar_init(smem_u32(&pa_bar), 2u * NCW);
      mbar_init(smem_u32(&qd_bar), 2u * NCW);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm(smem_u32(&tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: q | k | v weight slabs
    int slot = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
#pragma unroll 1
      for (int seg = 0; seg < 3; ++seg)
#pragma unroll 1
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[slot]), ph ^ 1u);
          const uint32_t bar = smem_u32(&full_bar[slot]);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(bar, 2u * slab_b);
            tma_load_3d_2sm(&tm.w, bar, ring + (uint32_t)slot * SLAB, kb * 64, seg * p.T + rank * (p.Np / 2), 0);
          }
          __syncwarp();
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      {
        // pull the NEXT tile's [T x 128 channels] slab of h into L2 (one 512-byte row segment per lane and iteration).  Issued
        // by this otherwise idle warp: in the compute warps the 196 prefetch instructions cost ~2 k cycles of every tile.
        const int nt = tile + n_clusters;
        if (nt < p.n_tiles) {
          const float* nsrc = p.h + ((size_t)(nt >> 1) * p.T) * D + (nt & 1) * 2 * ROWS + rank * ROWS;
          for (int t = lane; t < p.T; t += 32)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;" ::"l"(nsrc + (size_t)t * D) : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader)
    if (rank == 0) {
      int slot = 0, it = 0;
      uint32_t ph = 0;
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      auto gemm = [&](uint32_t dcol) {
#pragma unroll 1
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(smem_u32(&full_bar[slot]), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = opa + (uint32_t)kb * SLAB, b_addr = ring + (uint32_t)slot * SLAB;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2sm(tmem_base + dcol, make_smem_desc_sw128(a_addr + k * 32), make_smem_desc_sw128(b_addr + k * 32), idesc,
                           (kb == 0 && k == 0) ? 0u : 1u);
            umma_commit_2sm(smem_u32(&empty_bar[slot]), (uint16_t)3);
          }
          __syncwarp();
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      };
#pragma unroll 1
      for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters, ++it) {
        mbar_wait(smem_u32(&pa_bar), (uint32_t)it & 1u);    // operand tile written; every warp is past E_v of the previous tile
        tc_fence_after();
        gemm(0u);
        if (elect_one()) umma_commit_2sm(smem_u32(&tfull_bar[0]), (uint16_t)3);
        __syncwarp();
        gemm(256u);
        if (elect_one()) umma_commit_2sm(smem_u32(&tfull_bar[1]), (uint16_t)3);
        __syncwarp();
        mbar_wait(smem_u32(&qd_bar), (uint32_t)it & 1u);    // E_q has drained TMEM half 0
        tc_fence_after();
        gemm(0u);
        if (elect_one()) umma_commit_2sm(smem_u32(&tfull_bar[0]), (uint16_t)3);
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int part = ew >> 2;
    const int row = quad * 32 + lane;
    const int ctid = ew * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int Tq = p.nch * 32;
    float* const prm = reinterpret_cast<float*>(stg_gen);             // ln_w | ln_b | bq | bk | bv, Tq floats each (<= 5 KB)
    float* const xch = reinterpret_cast<float*>(stg_gen + 6144);      // [NCW][32][2]
    float* const bslot = reinterpret_cast<float*>(stg_gen + 5120) + ew * 32;   // per-warp bias broadcast slot
    // staging: E_q rows (32 x Tp fp16 <= 16 KB) / E_v tiles of warps 0..3 in operand-tile slabs 4..7; E_k / E_v tiles of
    // warps 4..7 in the staging region
    const uint32_t stg_q = opa + 4u * SLAB + (uint32_t)(ew & 3) * SLAB;
    uint16_t* const stg_q_gen = reinterpret_cast<uint16_t*>(opa_gen + 4 * SLAB + (ew & 3) * SLAB);
    const uint32_t stg_k = stg + 8192u + (uint32_t)(ew & 3) * 4096u;
    const uint32_t stg_mine = part == 0 ? stg_q : stg_k;
    const int c_split = (p.nch + 1) / 2;
    uint32_t tf_par[2] = {0u, 0u};
    auto wait_tf = [&](int hh) {
      mbar_wait(smem_u32(&tfull_bar[hh]), tf_par[hh]);
      tf_par[hh] ^= 1u;
      tc_fence_after();
    };
    bool pending = false;
    auto drain = [&]() {
      if (pending) {
        if (lane == 0) tma_wait_read0();
        __syncwarp();
        pending = false;
      }
    };
    // parameters do not depend on the sample: staged once
    for (int t = ctid; t < Tq; t += NCW * 32) {
      const bool ok = t < p.T;
      prm[t] = ok ? __ldg(p.ln_w + t) : 0.f;
      prm[Tq + t] = ok ? __ldg(p.ln_b + t) : 0.f;
      prm[2 * Tq + t] = ok ? __ldg(p.bqkv + t) : 0.f;
      prm[3 * Tq + t] = ok ? __ldg(p.bqkv + p.T + t) : 0.f;
      prm[4 * Tq + t] = ok ? __ldg(p.bqkv + 2 * p.T + t) : 0.f;
    }
    bar_sync(5, NCW * 32);
    long long tph[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) tph[i] = 0;
    const bool prof = p.prof != nullptr;
    long long tlast = prof ? clock64() : 0;
#define SF_TICK(i) do { if (prof) { const long long _n = clock64(); tph[i] += _n - tlast; tlast = _n; } } while (0)
    const int T2 = ((p.T + 1) / 2 + 7) / 8 * 8;               // frames [0, T2) to warps 0..3, [T2, T) to warps 4..7
    const int t_lo = part * T2, t_hi = part == 0 ? T2 : p.T;
    const int Tr8 = (p.T + 7) / 8 * 8;
    const int t_end = part == 0 ? T2 : Tr8;                   // the zero K padding [Tr8, nkb * 64) is written by warps 0..3,
                                                              // which have the shorter half of the frames
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
      const int b = tile >> 1;
      const int n0 = (tile & 1) * 2 * ROWS + rank * ROWS;
      // ---- P: LayerNorm over the T frames of channel n0 + row, written as row `row` of the operand tile
      {
        const float* __restrict__ src = p.h + ((size_t)b * p.T) * D + n0 + row;
        // 32 independent loads in flight per thread: with ~no L1 every frame row is an L2 / HBM round trip
        float K = __ldcg(src + (size_t)t_lo * D), sd = 0.f, sq = 0.f;
#pragma unroll 1
        for (int t0 = t_lo; t0 < t_hi; t0 += 32) {
          float x[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) x[e] = (t0 + e < t_hi) ? __ldcg(src + (size_t)(t0 + e) * D) : K;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float d = x[e] - K;
            sd += d;
            sq = fmaf(d, d, sq);
          }
        }
        SF_TICK(0);
        const float n_mine = (float)(t_hi - t_lo), n_o = (float)p.T - n_mine;
        const float mean_w = K + sd / n_mine;
        const float m2_w = sq - sd * sd / n_mine;
        xch[(ew * 32 + lane) * 2] = mean_w;
        xch[(ew * 32 + lane) * 2 + 1] = m2_w;
        bar_sync(1 + quad, 64);
        const float mean_o = xch[((ew ^ 4) * 32 + lane) * 2];
        const float m2_o = xch[((ew ^ 4) * 32 + lane) * 2 + 1];
        const float delta = mean_o - mean_w;
        const float mean = mean_w + delta * (n_o / (float)p.T);
        const float m2 = (m2_w + m2_o) + delta * delta * (n_mine * n_o / (float)p.T);
        const float rstd = rsqrtf(m2 / (float)p.T + 1e-5f);
        if (part == 0)
          for (int t8 = Tr8; t8 < p.nkb * 64; t8 += 8) st_shared_v4u(opa_addr(opa, row, t8), 0u, 0u, 0u, 0u);
#pragma unroll 1
        for (int t0 = t_lo; t0 < t_end; t0 += 32) {
          float x[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) x[e] = (t0 + e < p.T) ? __ldcg(src + (size_t)(t0 + e) * D) : 0.f;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int t8 = t0 + 8 * g8;
            if (t8 < t_end) {
              // gamma / beta are zero beyond T (staged that way) and x is 0 there: (0 - mean) rstd * 0 + 0 = 0 exactly
              const float4 ga = *reinterpret_cast<const float4*>(prm + t8), gb = *reinterpret_cast<const float4*>(prm + t8 + 4);
              const float4 ba = *reinterpret_cast<const float4*>(prm + Tq + t8), bb = *reinterpret_cast<const float4*>(prm + Tq + t8 + 4);
              const float* xx = x + 8 * g8;
              const f32x2 nm2 = pk2(-mean, -mean), r2 = pk2(rstd, rstd);     // packed fp32 pairs, same operations per element
              float y0, y1, y2, y3, y4, y5, y6, y7;
              upk2(fma2(mul2(add2(pk2(xx[0], xx[1]), nm2), r2), pk2(ga.x, ga.y), pk2(ba.x, ba.y)), y0, y1);
              upk2(fma2(mul2(add2(pk2(xx[2], xx[3]), nm2), r2), pk2(ga.z, ga.w), pk2(ba.z, ba.w)), y2, y3);
              upk2(fma2(mul2(add2(pk2(xx[4], xx[5]), nm2), r2), pk2(gb.x, gb.y), pk2(bb.x, bb.y)), y4, y5);
              upk2(fma2(mul2(add2(pk2(xx[6], xx[7]), nm2), r2), pk2(gb.z, gb.w), pk2(bb.z, bb.w)), y6, y7);
              st_shared_v4u(opa_addr(opa, row, t8), pack_f16x2_sat(y0, y1), pack_f16x2_sat(y2, y3), pack_f16x2_sat(y4, y5),
                            pack_f16x2_sat(y6, y7));
            }
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(smem_u32(&pa_bar), 0u);
        SF_TICK(1);
      }
      wait_tf(0);                                             // q complete
      SF_TICK(2);
      {
        // ---- E_q: per-head softmax of q + bq.  The quadrant's two warps take two heads each; a head's T/H <= 64 columns are
        // read as TWO chunks ALIGNED TO THE HEAD START (TMEM columns are addressable one by one), both loads in flight
        // together and kept in registers: one TMEM pass, one exp per element (max -> exp -> sum -> scale out of the same 64
        // registers), only per-chunk valid counts mask elements.  Normalised fp16 rows go to the staging tile as packed pairs.
        if (part == 0) drain();                                // the staging rows may still be read by this warp's last store
        bar_sync(1 + quad, 64);
        const int nv0 = min(32, p.hd), nv1 = p.hd - nv0;       // valid columns of the head's two chunks (hd <= 64)
#pragma unroll 1
        for (int hh = 2 * part; hh < 2 * part + 2; ++hh) {
          const int lo = hh * p.hd;
          float v0[32], v1[32];
          tmem_ld_32x32(trow + (uint32_t)lo, v0);
          tmem_ld_32x32(trow + (uint32_t)(lo + 32), v1);       // may reach into the next head / stale columns: masked by nv1
          const float b0 = lane < nv0 ? prm[2 * Tq + lo + lane] : 0.f;
          const float b1 = lane < nv1 ? prm[2 * Tq + lo + 32 + lane] : 0.f;
          tmem_ld_wait();
          add_bias32(bslot, lane, b0, v0);
          add_bias32(bslot, lane, b1, v1);
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            m = fmaxf(m, j < nv0 ? v0[j] : -INFINITY);
            m = fmaxf(m, j < nv1 ? v1[j] : -INFINITY);
          }
          const float ml = m * L2E;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v0[j] = j < nv0 ? ex2_fast(fmaf(v0[j], L2E, -ml)) : 0.f;
            v1[j] = j < nv1 ? ex2_fast(fmaf(v1[j], L2E, -ml)) : 0.f;
            s0 += v0[j];
            s1 += v1[j];
          }
          const float inv = 1.f / (s0 + s1);
          // row `lane` of the staging tile, columns [lo, lo + hd): 4-byte stores of packed pairs where the halfword index
          // is even (the pitch Tp is a multiple of 8, so the parity is that of lo), a single halfword at the ragged ends
          uint16_t* const dst = stg_q_gen + lane * p.Tp + lo;
          auto store_chunk = [&](const float* v, int nv, uint16_t* d) {
            if ((lo & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                if (j + 1 < nv) *reinterpret_cast<uint32_t*>(d + j) = pack_f16x2_sat(v[j] * inv, v[j + 1] * inv);
                else if (j < nv) d[j] = f32_to_f16_bits(v[j] * inv);
              }
            } else {
              if (0 < nv) d[0] = f32_to_f16_bits(v[0] * inv);
#pragma unroll
              for (int j = 1; j < 32; j += 2) {
                if (j + 1 < 32 && j + 1 < nv) *reinterpret_cast<uint32_t*>(d + j) = pack_f16x2_sat(v[j] * inv, v[j + 1] * inv);
                else if (j < nv) d[j] = f32_to_f16_bits(v[j] * inv);
              }
            }
          };
          store_chunk(v0, nv0, dst);
          if (nv1 > 0) store_chunk(v1, nv1, dst + 32);
        }
        if (part == 1)
          for (int col = p.T; col < p.Tp; ++col) stg_q_gen[lane * p.Tp + col] = 0;   // operand pad columns
        fence_async_smem();
        tc_fence_before();
        bar_sync(1 + quad, 64);                               // both warps' halves of the 32 staged rows are written
        if (lane == 0) mbar_arrive_remote(smem_u32(&qd_bar), 0u);
        if (part == 0) {
          if (lane == 0) {
            tma_store_3d(&tm.qs, stg_q, 0, b * D + n0 + quad * 32, 0);
            tma_commit();
          }
          pending = true;
        }
        SF_TICK(3);
      }
      {
        // ---- E_k: k + bk, fp32, transposed 32 x 32 tiles into [B, T, 512]; chunks split between the quadrant's two warps
        wait_tf(1);
#pragma unroll 1
        for (int c = (part == 0 ? 0 : c_split); c < (part == 0 ? c_split : p.nch); ++c) {
          float v[32];
          tmem_ld_32x32(trow + (uint32_t)(256 + c * 32), v);
          tmem_ld_wait();
          drain();
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(prm + 3 * Tq + c * 32 + 4 * q4);
            const float o[4] = {v[4 * q4] + b4.x, v[4 * q4 + 1] + b4.y, v[4 * q4 + 2] + b4.z, v[4 * q4 + 3] + b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg_mine + (uint32_t)((4 * q4 + e) * 32 + lane) * 4u), "f"(o[e]) : "memory");
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tm.k32, stg_mine, n0 + quad * 32, c * 32, b);
            tma_commit();
          }
          pending = true;
        }
        tc_fence_before();
        SF_TICK(4);
      }
      // ---- E_v: v + bv, fp16, transposed 32 x 32 tiles into [B, T, 512]; chunks split between the quadrant's two warps
      wait_tf(0);
      SF_TICK(5);
#pragma unroll 1
      for (int c = (part == 0 ? 0 : c_split); c < (part == 0 ? c_split : p.nch); ++c) {
        float v[32];
        tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        drain();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b4 = *reinterpret_cast<const float4*>(prm + 4 * Tq + c * 32 + 4 * q4);
          const float o[4] = {v[4 * q4] + b4.x, v[4 * q4 + 1] + b4.y, v[4 * q4 + 2] + b4.z, v[4 * q4 + 3] + b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint16_t h16 = f32_to_f16_bits(o[e]);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg_mine + (uint32_t)((4 * q4 + e) * 32 + lane) * 2u), "h"(h16) : "memory");
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tm.v16, stg_mine, n0 + quad * 32, c * 32, b);
          tma_commit();
        }
        pending = true;
      }
      tc_fence_before();
      SF_TICK(6);
    }
    if (lane == 0) tma_wait_all0();
    __syncwarp();
    if (prof && lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(p.prof + 8 * part + i, (unsigned long long)tph[i]);
    }
#undef SF_TICK
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}


// =================================================================================================================
// sa_ctx_kernel -- the per-head context of the channel attention (efficient_attention.py:36-39 on x^T): for a tile of
// 256 rows of k [B*T, 512] (fp32; row = (sample, feature d), columns = the 512 channel-tokens):
//   P   OPA = softmax over the 512 tokens of each row (warp per row)          [key softmax, dim=1 of (B, N, H, d)]
//   per sample s of the tile:
//   G   c = OPA v[s]^T        (B operand: the sample's value rows [T, 512], K-major)      ctx[d, l] = sum_n ks[n, d] v[n, l]
//   E   rows of s: keep the block-diagonal (head(d) == head(l)), fp16, transposed 32 x 32 tiles -> ctxT[s][l][d]
// Rounds alternate between the two TMEM halves, so the epilogue of one sample overlaps the MMAs of the next.
// =================================================================================================================
struct ScMaps {
  CUtensorMap v;
};
struct ScParams {
  const float* k32;
  uint16_t* ctxT;
  int rows, T, Tp, Np, hd, nch, batch, n_tiles;
};

__global__ void __launch_bounds__(THREADS, 1)
sa_ctx_kernel(const __grid_constant__ ScMaps tm, const __grid_constant__ ScParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NSLOT];
  __shared__ __align__(8) uint64_t empty_bar[NSLOT];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_slot;

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t opa = smem_base, ring = smem_base + OPA_BYTES;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const uint32_t slab_b = (uint32_t)(p.Np / 2) * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.v);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSLOT; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&tfull_bar[s]), 1);
        mbar_init(smem_u32(&tempty_bar[s]), 2u * NCW);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm(smem_u32(&tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: value slabs of every sample of the tile
    int slot = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
      const int g0 = tile * 2 * ROWS;
      const int s_first = g0 / p.T, s_last = (min(g0 + 2 * ROWS, p.rows) - 1) / p.T;
#pragma unroll 1
      for (int s = s_first; s <= s_last; ++s)
#pragma unroll 1
        for (int kb = 0; kb < D / 64; ++kb) {
          mbar_wait(smem_u32(&empty_bar[slot]), ph ^ 1u);
          const uint32_t bar = smem_u32(&full_bar[slot]);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(bar, 2u * slab_b);
            tma_load_3d_2sm(&tm.v, bar, ring + (uint32_t)slot * SLAB, kb * 64, rank * (p.Np / 2), s);
          }
          __syncwarp();
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader)
    if (rank == 0) {
      int slot = 0;
      uint32_t ph = 0;
      uint32_t te_par[2] = {0u, 0u};
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
#pragma unroll 1
      for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
        const int g0 = tile * 2 * ROWS;
        const int s_first = g0 / p.T, s_last = (min(g0 + 2 * ROWS, p.rows) - 1) / p.T;
#pragma unroll 1
        for (int s = s_first; s <= s_last; ++s) {
          const int hh = (s - s_first) & 1;
          mbar_wait(smem_u32(&tempty_bar[hh]), te_par[hh]);
          te_par[hh] ^= 1u;
          tc_fence_after();
#pragma unroll 1
          for (int kb = 0; kb < D / 64; ++kb) {
            mbar_wait(smem_u32(&full_bar[slot]), ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_addr = opa + (uint32_t)kb * SLAB, b_addr = ring + (uint32_t)slot * SLAB;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_2sm(tmem_base + (uint32_t)(hh * 256), make_smem_desc_sw128(a_addr + k * 32),
                             make_smem_desc_sw128(b_addr + k * 32), idesc, (kb == 0 && k == 0) ? 0u : 1u);
              umma_commit_2sm(smem_u32(&empty_bar[slot]), (uint16_t)3);
            }
            __syncwarp();
            if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
          }
          if (elect_one()) umma_commit_2sm(smem_u32(&tfull_bar[hh]), (uint16_t)3);
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int part = ew >> 2;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int c_split = (p.nch + 1) / 2;
    const int c_lo = part == 0 ? 0 : c_split, c_hi = part == 0 ? c_split : p.nch;
    uint32_t tf_par[2] = {0u, 0u};
    auto arrive = [&](int hh) {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(smem_u32(&tempty_bar[hh]), 0u);
    };
#pragma unroll 1
    for (int tile = cluster_id; tile < p.n_tiles; tile += n_clusters) {
      const int g0 = tile * 2 * ROWS, gc = g0 + rank * ROWS;
      const int s_first = g0 / p.T, s_last = (min(g0 + 2 * ROWS, p.rows) - 1) / p.T;
      const int nsamp = s_last - s_first + 1;
      // ---- P: softmax over the 512 tokens of every row -> fp16 operand tile (the arithmetic of softmax_seg_kernel)
#pragma unroll 1
      for (int i = 0; i < ROWS / NCW; i += 2) {
        float4 x[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const long long g = (long long)gc + ew + NCW * (i + u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            x[u][j] = g < p.rows ? __ldcg(reinterpret_cast<const float4*>(p.k32 + g * D) + j * 32 + lane)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = ew + NCW * (i + u);
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) m = fmaxf(m, fmaxf(fmaxf(x[u][j].x, x[u][j].y), fmaxf(x[u][j].z, x[u][j].w)));
          m = warp_max(m);
          float ssum = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            x[u][j].x = ex2_fast((x[u][j].x - m) * L2E); x[u][j].y = ex2_fast((x[u][j].y - m) * L2E);
            x[u][j].z = ex2_fast((x[u][j].z - m) * L2E); x[u][j].w = ex2_fast((x[u][j].w - m) * L2E);
            ssum += (x[u][j].x + x[u][j].y) + (x[u][j].z + x[u][j].w);
          }
          const float inv = 1.f / warp_sum(ssum);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = j * 128 + lane * 4;
            st_shared_v2u(opa_addr(opa, r, col) + (uint32_t)(lane & 1) * 8u, pack_f16x2_sat(x[u][j].x * inv, x[u][j].y * inv),
                          pack_f16x2_sat(x[u][j].z * inv, x[u][j].w * inv));
          }
        }
      }
      arrive(0);
      if (nsamp >= 2) arrive(1);
      // ---- rounds
#pragma unroll 1
      for (int s = s_first; s <= s_last; ++s) {
        const int hh = (s - s_first) & 1;
        mbar_wait(smem_u32(&tfull_bar[hh]), tf_par[hh]);
        tf_par[hh] ^= 1u;
        tc_fence_after();
        const int d = gc + quad * 32 + lane - s * p.T;          // this thread's feature index within sample s
        const bool mine = d >= 0 && d < p.T;
        if (__any_sync(0xffffffffu, mine)) {
          // ctxT[s][l][d]: for a fixed l the 32 lanes (consecutive d) write 64 contiguous bytes.  (A transposed TMA tile
          // store cannot be used: its start coordinate d0 = row - s T is not 16-byte aligned in general.)
          const int dh = mine ? d / p.hd : -1;
          uint16_t* const dst = p.ctxT + ((size_t)s * p.T) * p.Tp + (mine ? d : 0);
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            float v[32];
            tmem_ld_32x32(trow + (uint32_t)(hh * 256 + c * 32), v);
            tmem_ld_wait();
            if (mine) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int l = c * 32 + j;
                if (l < p.T) dst[(size_t)l * p.Tp] = (l / p.hd == dh) ? f32_to_f16_bits(v[j]) : (uint16_t)0;
              }
            }
          }
        }
        if (s + 2 <= s_last) arrive(hh);                         // a later round reuses this TMEM half
        else tc_fence_before();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}


std::atomic<unsigned long long> g_fb_launches{0};
int g_fb_max_pairs = -1;
unsigned long long* g_fb_prof = nullptr;

// All four kernels of this file are persistent CTA-pair kernels with the same launch shape: cluster of 2, THREADS
// threads, SMEM_BYTES of dynamic shared memory, one pair per SM pair, programmatic dependent launch.
template <typename K>
int pair_capacity(K kernel, int* cached, const char* name) {
  if (*cached > 0) return 0;
  MCM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tc_num_sms() / 2 * 2);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  MCM_CUDA(cudaOccupancyMaxActiveClusters(&n, kernel, &cfg));
  if (n <= 0) {
    set_error(std::string(name) + " does not fit on this device");
    return 1;
  }
  *cached = n;
  return 0;
}

template <typename K, typename M, typename P>
int launch_pairs(K kernel, int n_pairs, int kind, double flops, const M& maps, const P& params, cudaStream_t stream) {
  {
    LaunchTimer lt(kind, stream, flops);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_pairs);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MCM_CUDA(cudaLaunchKernelEx(&cfg, kernel, maps, params));
  }
  MCM_CUDA(cudaGetLastError());
  g_fb_launches.fetch_add(1);
  return 0;
}

int fb_init() {
  if (g_fb_max_pairs >= 0) return 0;
  MCM_TRY(gemm_tc_init());
  int n = 0;
  MCM_TRY(pair_capacity(fused_block_kernel, &n, "fused_block_kernel"));
  g_fb_max_pairs = n;
  if (const char* e = getenv("MCM_FUSED_PROF")) {
    if (e[0] == '1') {
      MCM_CUDA(cudaMalloc(&g_fb_prof, 32 * sizeof(unsigned long long)));
      MCM_CUDA(cudaMemset(g_fb_prof, 0, 32 * sizeof(unsigned long long)));
    }
  }
  return 0;
}

}  // namespace

bool fused_block_supported(int T, int Dm, int Fm, int Hm) {
  return Dm == D && Fm == F && Hm == H && (2 * ROWS - 1) / T + 2 <= MAX_SAMPLES;
}
size_t fused_block_hid_bytes() { return (size_t)160 * ROWS * F * 2; }
int fused_block_max_pairs() { return fb_init() == 0 ? g_fb_max_pairs : 0; }
unsigned long long fused_block_launch_count() { return g_fb_launches.load(); }
int fused_block_prof_read(unsigned long long* out, int reset) {
  for (int i = 0; i < 32; ++i) out[i] = 0;
  if (g_fb_prof == nullptr) return 0;
  MCM_CUDA(cudaDeviceSynchronize());
  MCM_CUDA(cudaMemcpy(out, g_fb_prof, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) MCM_CUDA(cudaMemset(g_fb_prof, 0, 32 * sizeof(unsigned long long)));
  return 0;
}
void fused_block_count_replayed(unsigned long long n) { g_fb_launches.fetch_add(n); }

int fused_block_launch(const FusedBlockArgs& a, cudaStream_t stream) {
  MCM_TRY(fb_init());
  MCM_CHECK(a.h && a.hid && a.rows > 0 && a.T > 0 && a.batch > 0 && a.rows == a.batch * a.T, "fused block: bad arguments");
  MCM_CHECK((2 * ROWS - 1) / a.T + 2 <= MAX_SAMPLES, "fused block: sequence too short for the fused kernel");
  MCM_CHECK(a.ca_wq.ld == D && a.ca_wo.ld == D && a.f_w1.ld == D && a.f_w2.ld == F && a.f_wo.ld == D && a.ca_ctxT.ld == HD,
            "fused block: unexpected operand pitch");
  MCM_CHECK(a.mod_ld % 4 == 0, "fused block: modulation pitch");
  const int n_tiles = (a.rows + 2 * ROWS - 1) / (2 * ROWS);
  const int n_pairs = std::min(n_tiles, std::max(1, g_fb_max_pairs / tc_sm_share()));
  const int grid = 2 * n_pairs;
  MCM_CHECK((size_t)grid * ROWS * F * 2 <= fused_block_hid_bytes(), "fused block: hidden scratch too small");

  FbMaps tm;
  std::memset(&tm, 0, sizeof(tm));
  MCM_TRY(tc_make_operand_map(&tm.wq, a.ca_wq.hi, OP_F16, D, D, 1, D, 128));
  MCM_TRY(tc_make_operand_map(&tm.ctx, a.ca_ctxT.hi, OP_F16, HD, HD, a.batch * H, HD, 64));
  MCM_TRY(tc_make_operand_map(&tm.wo, a.ca_wo.hi, OP_F16, D, D, 1, D, 128));
  MCM_TRY(tc_make_operand_map(&tm.w1, a.f_w1.hi, OP_F16, D, F, 1, D, 128));
  MCM_TRY(tc_make_operand_map(&tm.hida, a.hid, OP_F16, F, grid * ROWS, 1, F, 128));
  MCM_TRY(tc_make_operand_map(&tm.w2, a.f_w2.hi, OP_F16, F, D, 1, F, 128));
  MCM_TRY(tc_make_operand_map(&tm.wo2, a.f_wo.hi, OP_F16, D, D, 1, D, 128));
  MCM_TRY(tc_make_tile_map(&tm.hred, a.h, 0, D, a.rows, 1, D, (long long)a.rows * D, 128));
  MCM_TRY(tc_make_tile_map(&tm.hidst, a.hid, 1, F, (long long)grid * ROWS, 1, F, (long long)grid * ROWS * F, 64));

  FbParams p;
  std::memset(&p, 0, sizeof(p));
  p.h = a.h; p.rows = a.rows; p.T = a.T; p.batch = a.batch; p.n_tiles = n_tiles; p.mod_ld = a.mod_ld; p.stop = a.stop;
  p.ca_ln_w = a.ca_ln_w; p.ca_ln_b = a.ca_ln_b; p.ca_bq = a.ca_bq; p.ca_pn_w = a.ca_pn_w; p.ca_pn_b = a.ca_pn_b;
  p.ca_scale = a.ca_scale; p.ca_shift = a.ca_shift; p.ca_bo = a.ca_bo;
  p.f_b1 = a.f_b1; p.f_b2 = a.f_b2; p.f_pn_w = a.f_pn_w; p.f_pn_b = a.f_pn_b; p.f_scale = a.f_scale; p.f_shift = a.f_shift;
  p.f_bo = a.f_bo;
  p.dbg = reinterpret_cast<uint16_t*>(a.dbg);
  static const int stagger = [] { const char* e = getenv("MCM_FB_STAGGER"); return e ? atoi(e) : 0; }();
  p.stagger = stagger;
  p.prof = (getenv("MCM_SF_PROF") != nullptr || getenv("MCM_ST_PROF") != nullptr) ? nullptr : g_fb_prof;

  // algorithmic flops of the two sub-blocks (AdaLN emb GEMM is not in this kernel)
  const double flops = 2.0 * (double)a.rows * ((double)D * D * 2 + (double)D * HD + 2.0 * D * F + (double)D * D);
  return launch_pairs(fused_block_kernel, n_pairs, LK_FUSED, flops, tm, p, stream);
}

bool sa_tail_supported(int T, int Dm) { return NCW == 8 && Dm == D && T >= 40 && T <= 256 && T % 4 == 0; }

int sa_tail_launch(const SaTailArgs& a, cudaStream_t stream) {
  MCM_TRY(fb_init());
  MCM_CHECK(a.h && a.qs.hi && a.ctxT.hi && a.wo.hi && a.batch > 0 && sa_tail_supported(a.T, D), "sa tail: bad arguments");
  static int max_pairs = -1;
  MCM_TRY(pair_capacity(sa_tail_kernel, &max_pairs, "sa_tail_kernel"));
  const int T = a.T, Tp = a.qs.ld;
  const int Np = (T + 15) / 16 * 16;
  MCM_CHECK(a.ctxT.ld == Tp && a.wo.ld == Tp && Tp % 8 == 0 && Tp >= T, "sa tail: operand pitch");
  StMaps tm;
  std::memset(&tm, 0, sizeof(tm));
  MCM_TRY(tc_make_operand_map(&tm.qs, a.qs.hi, OP_F16, Tp, a.batch * D, 1, Tp, 128));
  MCM_TRY(tc_make_operand_map(&tm.ctx, a.ctxT.hi, OP_F16, Tp, T, a.batch, Tp, Np / 2));
  MCM_TRY(tc_make_operand_map(&tm.wo, a.wo.hi, OP_F16, Tp, T, 1, Tp, Np / 2));
  MCM_TRY(tc_make_tile_map(&tm.hred, a.h, 0, D, T, a.batch, D, (long long)T * D, 0));
  StParams p;
  std::memset(&p, 0, sizeof(p));
  p.T = T; p.Np = Np; p.nkb = (T + 63) / 64; p.nch = (T + 31) / 32; p.batch = a.batch; p.n_tiles = a.batch * 2;
  p.mod_ld = a.mod_ld; p.pn_w = a.pn_w; p.pn_b = a.pn_b; p.scale = a.scale; p.shift = a.shift; p.bo = a.bo;
  p.prof = (g_fb_prof != nullptr && getenv("MCM_ST_PROF") != nullptr) ? g_fb_prof : nullptr;
  p.h = a.h;
  static const int pf = [] { const char* e = getenv("MCM_ST_PREFETCH"); return e ? atoi(e) : 1; }();
  p.pf = pf;
  const int n_pairs = std::min(p.n_tiles, std::max(1, max_pairs / tc_sm_share()));
  const double flops = 2.0 * (double)a.batch * D * ((double)T * (T / 4) + (double)T * T);   // q ctx (per head) + out
  return launch_pairs(sa_tail_kernel, n_pairs, LK_GEMM, flops, tm, p, stream);
}

// generic [d0 contiguous, d1, d2] tensor map with an explicit box (the E_q row store: box {Tp, 32, 1}, no swizzle)
int sa_front_launch(const SaFrontArgs& a, cudaStream_t stream) {
  MCM_TRY(fb_init());
  MCM_CHECK(a.h && a.w.hi && a.qs.hi && a.k32 && a.v16.hi && a.batch > 0 && sa_tail_supported(a.T, D), "sa front: bad arguments");
  static int max_pairs = -1;
  MCM_TRY(pair_capacity(sa_front_kernel, &max_pairs, "sa_front_kernel"));
  const int T = a.T, Tp = a.w.ld;
  const int Np = (T + 15) / 16 * 16;
  MCM_CHECK(a.qs.ld == Tp && Tp % 8 == 0 && Tp >= T && a.v16.ld == D && a.heads == H && T % H == 0 && 32 * Tp * 2 <= SLAB,
            "sa front: operand layout");
  SfMaps tm;
  std::memset(&tm, 0, sizeof(tm));
  MCM_TRY(tc_make_operand_map(&tm.w, a.w.hi, OP_F16, Tp, 3 * T, 1, Tp, Np / 2));
  MCM_TRY(tc_make_box_map(&tm.qs, a.qs.hi, 1, Tp, (long long)a.batch * D, 1, Tp, (long long)a.batch * D * Tp, Tp, 32));
  MCM_TRY(tc_make_tile_map(&tm.k32, a.k32, 0, D, T, a.batch, D, (long long)T * D, 0));
  MCM_TRY(tc_make_tile_map(&tm.v16, a.v16.hi, 1, D, T, a.batch, D, (long long)T * D, 0));
  SfParams p;
  std::memset(&p, 0, sizeof(p));
  p.h = a.h; p.T = T; p.Tp = Tp; p.Np = Np; p.nkb = (T + 63) / 64; p.nch = (T + 31) / 32; p.hd = T / H; p.batch = a.batch;
  p.n_tiles = a.batch * 2; p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.bqkv = a.bqkv;
  p.prof = (g_fb_prof != nullptr && getenv("MCM_SF_PROF") != nullptr) ? g_fb_prof : nullptr;
  const int n_pairs = std::min(p.n_tiles, std::max(1, max_pairs / tc_sm_share()));
  const double flops = 2.0 * (double)a.batch * D * (double)T * 3.0 * T;
  return launch_pairs(sa_front_kernel, n_pairs, LK_GEMM, flops, tm, p, stream);
}

int sa_ctx_launch(const SaCtxArgs& a, cudaStream_t stream) {
  MCM_TRY(fb_init());
  MCM_CHECK(a.k32 && a.v16.hi && a.ctxT.hi && a.batch > 0 && sa_tail_supported(a.T, D) && a.heads > 0 && a.T % a.heads == 0,
            "sa ctx: bad arguments");
  static int max_pairs = -1;
  MCM_TRY(pair_capacity(sa_ctx_kernel, &max_pairs, "sa_ctx_kernel"));
  const int T = a.T, Tp = a.ctxT.ld;
  const int Np = (T + 15) / 16 * 16;
  MCM_CHECK(a.v16.ld == D && Tp % 8 == 0 && Tp >= T, "sa ctx: operand layout");
  ScMaps tm;
  std::memset(&tm, 0, sizeof(tm));
  MCM_TRY(tc_make_operand_map(&tm.v, a.v16.hi, OP_F16, D, T, a.batch, D, Np / 2));
  ScParams p;
  std::memset(&p, 0, sizeof(p));
  p.k32 = a.k32; p.ctxT = reinterpret_cast<uint16_t*>(a.ctxT.hi); p.Tp = Tp; p.rows = a.batch * T; p.T = T; p.Np = Np; p.hd = T / a.heads; p.nch = (T + 31) / 32; p.batch = a.batch;
  p.n_tiles = (p.rows + 2 * ROWS - 1) / (2 * ROWS);
  const int n_pairs = std::min(p.n_tiles, std::max(1, max_pairs / tc_sm_share()));
  const double flops = 2.0 * (double)a.batch * T * (T / a.heads) * D;      // only the per-head diagonal blocks are algorithmic
  return launch_pairs(sa_ctx_kernel, n_pairs, LK_GEMM, flops, tm, p, stream);
}

}  // namespace mcm
