// motioncraft_b200 -- row kernels (see elementwise.cuh).  One warp per row, float4 loads, warp-shuffle
// reductions, 8-byte packed 16-bit stores; the transposing LayerNorm stages a [T x 32] tile in shared
// memory so both its reads (along d) and its writes (along t) are coalesced.
#include "elementwise.cuh"
#include "timing.cuh"

#include <atomic>

namespace mcm {
namespace {
std::atomic<unsigned long long> g_ew_launches{0};
constexpr int MAXV_LIMIT = 8;   // float4 per lane -> rows of up to 1024 elements

// The row kernels were instruction-issue bound (ncu: 72-90 % issue-active at 2-2.5 TB/s), so the fp16
// ("fast") operand path uses single-instruction saturating converts and MUFU-based exp / reciprocal
// (relative error ~1e-6, far below the 2^-11 rounding of the fp16 value being produced); the bf16x2
// ("precise") path keeps the IEEE expf / division so that mode stays a clean numerical reference.
template <int FMT> __device__ __forceinline__ float exp_sel(float x) {
  if (FMT == OP_F16) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
  }
  return expf(x);
}
template <int FMT> __device__ __forceinline__ float silu_sel(float x) {
  if (FMT == OP_F16) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * 1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
    return x * r;
  }
  return x / (1.f + expf(-x));
}
__device__ __forceinline__ uint32_t pk_f16(float a, float b) {   // lo half = a; saturating
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint16_t cv_f16(float a) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pk_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// store 4 consecutive operand elements at element index idx (idx % 4 == 0)
template <int FMT>
__device__ __forceinline__ void st4(uint16_t* hi, uint16_t* lo, size_t idx, float a, float b, float c, float d) {
  if (FMT == OP_F16) {
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2(pk_f16(a, b), pk_f16(c, d));
  } else {
    const uint32_t h0 = pk_bf16(a, b), h1 = pk_bf16(c, d);
    const uint32_t l0 = pk_bf16(a - __uint_as_float(h0 << 16), b - __uint_as_float(h0 & 0xffff0000u));
    const uint32_t l1 = pk_bf16(c - __uint_as_float(h1 << 16), d - __uint_as_float(h1 & 0xffff0000u));
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l0, l1);
  }
}
template <int FMT>
__device__ __forceinline__ void st1(uint16_t* hi, uint16_t* lo, size_t idx, float a) {
  if (FMT == OP_F16) {
    hi[idx] = cv_f16(a);
  } else {
    uint16_t h, l;
    f32_to_bf16x2_bits(a, h, l);
    hi[idx] = h;
    lo[idx] = l;
  }
}
template <int FMT>
__device__ __forceinline__ void st2(uint16_t* hi, uint16_t* lo, size_t idx, float a, float b) {   // idx % 2 == 0
  if (FMT == OP_F16) {
    *reinterpret_cast<uint32_t*>(hi + idx) = pk_f16(a, b);
  } else {
    const uint32_t h0 = pk_bf16(a, b);
    *reinterpret_cast<uint32_t*>(hi + idx) = h0;
    *reinterpret_cast<uint32_t*>(lo + idx) = pk_bf16(a - __uint_as_float(h0 << 16), b - __uint_as_float(h0 & 0xffff0000u));
  }
}

// ------------------------------------------------------------------------------------------ ln_rows
template <int MAXV, int FMT>
__global__ void __launch_bounds__(256, MAXV >= 8 ? 4 : 1)
ln_rows_kernel(const float* __restrict__ in, int rows, int d, int ld_in, const float* __restrict__ w,
               const float* __restrict__ b, const float* __restrict__ scale, const float* __restrict__ shift,
               int mod_ld, int rows_per_batch, int act_silu, uint16_t* __restrict__ ohi, uint16_t* __restrict__ olo,
               int out_ld) {
  pdl_trigger();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = in + (size_t)row * ld_in;
  const int nv = d >> 2;                       // float4 count (d % 4 == 0)
  const int batch = row / rows_per_batch;
  const float* sc = scale ? scale + (size_t)batch * mod_ld : nullptr;
  const float* sh = shift ? shift + (size_t)batch * mod_ld : nullptr;
  const int nvo = out_ld >> 2;
  const size_t obase = (size_t)row * out_ld;
  float4 v[MAXV];
  float s = 0.f;
  pdl_wait();                                  // as late as possible: every kernel parameter has been fetched by now
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = i * 32 + lane;
    if (q < nv) {
      v[i] = __ldcs(reinterpret_cast<const float4*>(x + 4 * q));     // streamed once: keep L1 for the parameter vectors
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)d;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = i * 32 + lane;
    if (q < nv) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      ss += (a * a + bb * bb) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = i * 32 + lane;
    if (q < nv) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + 4 * q));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + 4 * q));
      float y[4] = {fmaf((v[i].x - mean) * rstd, ww.x, bb.x), fmaf((v[i].y - mean) * rstd, ww.y, bb.y),
                    fmaf((v[i].z - mean) * rstd, ww.z, bb.z), fmaf((v[i].w - mean) * rstd, ww.w, bb.w)};
      if (sc) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + 4 * q));
        const float4 h4 = __ldg(reinterpret_cast<const float4*>(sh + 4 * q));
        y[0] = fmaf(y[0], 1.f + s4.x, h4.x);
        y[1] = fmaf(y[1], 1.f + s4.y, h4.y);
        y[2] = fmaf(y[2], 1.f + s4.z, h4.z);
        y[3] = fmaf(y[3], 1.f + s4.w, h4.w);
      }
      if (act_silu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = silu_sel<FMT>(y[j]);
      }
      st4<FMT>(ohi, olo, obase + 4 * q, y[0], y[1], y[2], y[3]);
    } else if (q < nvo) {
      st4<FMT>(ohi, olo, obase + 4 * q, 0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------ softmax_seg
// LPS lanes cooperate on one segment (EPL elements per lane, seg <= LPS * EPL); a warp handles 32 / LPS
// consecutive segments.  Short segments (the 49-wide heads of the channel attention) use 8 lanes each so a
// warp covers a whole 196-float row with four 32-byte-granular streams instead of wasting 3/4 of its lanes.
template <int LPS, int EPL, int FMT>
__global__ void __launch_bounds__(256)
softmax_seg_kernel(const float* __restrict__ in, long long total_segs, int ncols, int ld_in, int seg, int nseg,
                   uint16_t* __restrict__ ohi, uint16_t* __restrict__ olo, int out_ld) {
  pdl_trigger();
  constexpr int SPW = 32 / LPS;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPS;
  const long long gs = wid * SPW + lane / LPS;
  const bool active = gs < total_segs;
  const long long row = active ? gs / nseg : 0;
  const int sidx = active ? (int)(gs - row * nseg) : 0;
  const float* x = in + (size_t)row * ld_in + (size_t)sidx * seg + sub;
  const size_t obase = (size_t)row * out_ld + (size_t)sidx * seg + sub;
  float v[EPL];
  float m = -INFINITY;
  pdl_wait();                                  // after the index arithmetic (which fetches every kernel parameter)
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    v[i] = (active && i * LPS + sub < seg) ? x[i * LPS] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    v[i] = exp_sel<FMT>(v[i] - m);        // exp(-inf) = 0 for the lanes past the segment end
    s += v[i];
  }
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (!active) return;
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < EPL; ++i)
    if (i * LPS + sub < seg) st1<FMT>(ohi, olo, obase + i * LPS, v[i] * inv);
  if (sidx == nseg - 1) {
    for (int c = ncols + sub; c < out_ld; c += LPS) st1<FMT>(ohi, olo, (size_t)row * out_ld + c, 0.f);
  }
}

// ------------------------------------------------------------------------------------------ ln_transpose
// grid (D/32, B), block (32, NW).  tile[t][dx] (+1 pad) holds h[b, t, d0 + dx].  NW = 8 warps for short sequences (several
// blocks per SM); long ones (T > 256: the tile is up to 135 KB, ONE block per SM) take 32 warps so that the strided loads
// of the column statistics still have ~32 x 8 requests in flight per SM.
template <int FMT, int NW>
__global__ void __launch_bounds__(32 * NW)
ln_transpose_kernel(const float* __restrict__ h, int T, int D, const float* __restrict__ w,
                    const float* __restrict__ b, uint16_t* __restrict__ ohi, uint16_t* __restrict__ olo, int out_ld) {
  pdl_trigger();
  extern __shared__ float tile[];            // T * 33
  __shared__ float red[NW][33];
  __shared__ float mean_s[32], rstd_s[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int d0 = blockIdx.x * 32;
  const int bidx = blockIdx.y;
  const float* src = h + (size_t)bidx * T * D + d0 + tx;
  float s = 0.f;
  pdl_wait();
#pragma unroll 8
  for (int t = ty; t < T; t += NW) {
    const float val = src[(size_t)t * D];
    tile[t * 33 + tx] = val;
    s += val;
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += red[i][tx];
    mean_s[tx] = tot / (float)T;
  }
  __syncthreads();
  const float mean = mean_s[tx];
  float ss = 0.f;
  for (int t = ty; t < T; t += NW) {
    const float dlt = tile[t * 33 + tx] - mean;
    ss += dlt * dlt;
  }
  red[ty][tx] = ss;
  __syncthreads();
  if (ty == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += red[i][tx];
    rstd_s[tx] = rsqrtf(tot / (float)T + 1e-5f);
  }
  __syncthreads();
  // write: warp ty handles columns dl = ty, ty+8, ...; lanes run along t (conflict-free smem reads, stride 33);
  // neighbouring lanes exchange values so every even lane emits one packed 32-bit store of two consecutive t
  for (int dl = ty; dl < 32; dl += NW) {
    const float mu = mean_s[dl], rs = rstd_s[dl];
    const size_t obase = ((size_t)bidx * D + d0 + dl) * out_ld;
    for (int t0 = 0; t0 < out_ld; t0 += 32) {
      const int t = t0 + tx;
      const float y = (t < T) ? fmaf((tile[t * 33 + dl] - mu) * rs, __ldg(w + t), __ldg(b + t)) : 0.f;
      const float yn = __shfl_down_sync(0xffffffffu, y, 1);
      if ((tx & 1) == 0 && t < out_ld) st2<FMT>(ohi, olo, obase + t, y, yn);
    }
  }
}

// Long sequences (T > 512): 16 channels per block -> a 17-float pitch tile of <= 70 KB, so THREE 512-thread blocks share
// an SM and one block's strided loads overlap another's write phase (with the 135 KB tile of the 32-channel version one
// block owns the SM and its load / statistics / write phases run strictly one after the other).  A half-warp covers the 16
// channels (64 contiguous bytes) of one frame; in the write phase a warp owns one channel and every lane two frames.
template <int FMT>
__global__ void __launch_bounds__(512)
ln_transpose16_kernel(const float* __restrict__ h, int T, int D, const float* __restrict__ w,
                      const float* __restrict__ b, uint16_t* __restrict__ ohi, uint16_t* __restrict__ olo, int out_ld) {
  pdl_trigger();
  pdl_wait();
  constexpr int CH = 16, NW = 16, P = CH + 1;
  extern __shared__ float tile[];            // T * 17
  __shared__ float red[NW][33];
  __shared__ float mean_s[CH], rstd_s[CH];
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int cx = lane & (CH - 1), tsub = lane >> 4;
  const int d0 = blockIdx.x * CH;
  const int bidx = blockIdx.y;
  const float* src = h + (size_t)bidx * T * D + d0 + cx;
  float s = 0.f;
#pragma unroll 8
  for (int t = ty * 2 + tsub; t < T; t += NW * 2) {
    const float val = src[(size_t)t * D];
    tile[t * P + cx] = val;
    s += val;
  }
  red[ty][lane] = s;
  __syncthreads();
  if (ty == 0 && lane < CH) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += red[i][lane] + red[i][lane + CH];
    mean_s[lane] = tot / (float)T;
  }
  __syncthreads();
  const float mean = mean_s[cx];
  float ss = 0.f;
  for (int t = ty * 2 + tsub; t < T; t += NW * 2) {
    const float dlt = tile[t * P + cx] - mean;
    ss += dlt * dlt;
  }
  red[ty][lane] = ss;
  __syncthreads();
  if (ty == 0 && lane < CH) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += red[i][lane] + red[i][lane + CH];
    rstd_s[lane] = rsqrtf(tot / (float)T + 1e-5f);
  }
  __syncthreads();
  {
    const int dl = ty;
    const float mu = mean_s[dl], rs = rstd_s[dl];
    const size_t obase = ((size_t)bidx * D + d0 + dl) * out_ld;
    for (int t0 = 0; t0 < out_ld; t0 += 64) {
      const int t = t0 + 2 * lane;
      if (t >= out_ld) break;
      const float y0 = (t < T) ? fmaf((tile[t * P + dl] - mu) * rs, __ldg(w + t), __ldg(b + t)) : 0.f;
      const float y1 = (t + 1 < T) ? fmaf((tile[(t + 1) * P + dl] - mu) * rs, __ldg(w + t + 1), __ldg(b + t + 1)) : 0.f;
      st2<FMT>(ohi, olo, obase + t, y0, y1);
    }
  }
}

// ------------------------------------------------------------------------------------------ pack_op
__global__ void __launch_bounds__(256)
pack_op_kernel(const float* __restrict__ in, size_t rows, int cols, int ld_in, int act_silu, OpPtr out, int out_fmt) {
  pdl_trigger();
  const size_t total = rows * (size_t)out.ld;
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / out.ld;
    const int c = (int)(i - r * out.ld);
    float v = 0.f;
    if (c < cols) {
      v = in[r * ld_in + c];
      if (act_silu) v = silu(v);
    }
    op_store1(out, out_fmt, i, v);
  }
}

// 8 elements per thread (two 16-byte loads, one 16-byte store per half): cols, both pitches and the bases 8-element aligned
__global__ void __launch_bounds__(256)
pack_op8_kernel(const float* __restrict__ in, size_t rows, int cols, int ld_in, int act_silu, OpPtr out, int out_fmt) {
  pdl_trigger();
  const int gpr = out.ld >> 3;
  const size_t total = rows * (size_t)gpr;
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / gpr;
    const int c = (int)(i - r * gpr) * 8;
    float v[8];
    if (c < cols) {
      const float4 a = *reinterpret_cast<const float4*>(in + r * ld_in + c);
      const float4 b = *reinterpret_cast<const float4*>(in + r * ld_in + c + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      if (act_silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = silu(v[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    op_store8(out, out_fmt, r * out.ld + c, v);
  }
}

// ------------------------------------------------------------------------------------------ timestep embedding
__global__ void timestep_embedding_kernel(const long long* __restrict__ t_dev, int t_uniform, int B, int dim,
                                          OpPtr out, int out_fmt) {
  pdl_trigger();
  pdl_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * out.ld) return;
  const int bidx = i / out.ld;
  const int c = i - bidx * out.ld;
  float v = 0.f;
  if (c < 2 * half) {
    const float t = (float)(t_dev ? t_dev[bidx] : (long long)t_uniform);
    const int k = c < half ? c : c - half;
    // freqs = exp(-ln(1e4) * k / half) in fp32, exactly as position_encoding.py:53-55
    const float freq = expf(-9.210340371976184f * (float)k / (float)half);
    const float arg = t * freq;
    v = c < half ? cosf(arg) : sinf(arg);
  }
  op_store1(out, out_fmt, (size_t)i, v);
}

// ------------------------------------------------------------------------------------------ sampler updates
__global__ void __launch_bounds__(256)
ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eps_model, const float* __restrict__ noise,
                   float* __restrict__ x_out, size_t rows, int cols, DdimCoefs k, OpPtr xop, int op_fmt) {
  pdl_trigger();
  pdl_wait();
  // gaussian_diffusion.py:572-577 (x0 from eps), :587-591 (eps re-derived), :839-852 (Equation 12);
  // same fp32 operation order as the reference, no fused multiply-adds
  const float one_m_abp = __fsub_rn(1.f, k.alpha_bar_prev);
  const float sigma = __fmul_rn(__fmul_rn(k.eta, sqrtf(__fdiv_rn(one_m_abp, __fsub_rn(1.f, k.alpha_bar)))),
                                sqrtf(__fsub_rn(1.f, __fdiv_rn(k.alpha_bar, k.alpha_bar_prev))));
  const float s_abp = sqrtf(k.alpha_bar_prev);
  const float s_dir = sqrtf(__fsub_rn(one_m_abp, __fmul_rn(sigma, sigma)));
  const int ldo = xop.hi ? xop.ld : cols;
  const size_t total = rows * (size_t)ldo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ldo;
    const int c = (int)(i - r * ldo);
    float xn = 0.f;
    if (c < cols) {
      const size_t j = r * cols + c;
      const float c1x = __fmul_rn(k.c1, x[j]);
      const float x0 = k.start_x ? eps_model[j] : __fsub_rn(c1x, __fmul_rn(k.c2, eps_model[j]));
      const float eps = __fdiv_rn(__fsub_rn(c1x, x0), k.c2);
      xn = __fadd_rn(__fmul_rn(x0, s_abp), __fmul_rn(s_dir, eps));
      if (k.add_noise) xn = __fadd_rn(xn, __fmul_rn(sigma, noise[j]));
      x_out[j] = xn;
    }
    if (xop.hi) op_store1(xop, op_fmt, i, xn);
  }
}

__global__ void __launch_bounds__(256)
ddpm_update_kernel(const float* __restrict__ x, const float* __restrict__ eps_model, const float* __restrict__ noise,
                   float* __restrict__ x_out, size_t rows, int cols, DdpmCoefs k, OpPtr xop, int op_fmt) {
  pdl_trigger();
  pdl_wait();
  // gaussian_diffusion.py:572-577, :445-449 (posterior mean), :694 (mean + exp(0.5 logvar) * noise)
  const float sd = expf(__fmul_rn(0.5f, k.log_var));
  const int ldo = xop.hi ? xop.ld : cols;
  const size_t total = rows * (size_t)ldo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ldo;
    const int c = (int)(i - r * ldo);
    float xn = 0.f;
    if (c < cols) {
      const size_t j = r * cols + c;
      const float x0 = k.start_x ? eps_model[j] : __fsub_rn(__fmul_rn(k.c1, x[j]), __fmul_rn(k.c2, eps_model[j]));
      xn = __fadd_rn(__fmul_rn(k.pm1, x0), __fmul_rn(k.pm2, x[j]));
      if (k.add_noise) xn = __fadd_rn(xn, __fmul_rn(sd, noise[j]));
      x_out[j] = xn;
    }
    if (xop.hi) op_store1(xop, op_fmt, i, xn);
  }
}

// RePaint / outpainting blend after a DDIM update (ddim_sample, gaussian_diffusion.py:855-879, same_overlap_noisy=False):
//   weighed_gt = sqrt(abar_prev) gt + sqrt(1 - abar_prev) noise ; over the first `overlap` frames, once the noise weight is
//   below 0.2 (and opt.addBlend), weighed_gt = weighed_gt (1 - lw[t]) + x lw[t] ; x <- keep_mask ? weighed_gt : x
// in the reference's fp32 op order (no FMA contraction).  Also rewrites the operand copy of x.
__global__ void __launch_bounds__(256)
repaint_blend_kernel(float* __restrict__ x, const float* __restrict__ gt, const unsigned char* __restrict__ keep,
                     const float* __restrict__ noise, size_t rows, int cols, int T, float gt_w, float noise_w,
                     const float* __restrict__ blend_w, int overlap, OpPtr xop, int op_fmt) {
  pdl_trigger();
  pdl_wait();
  const int ldo = xop.hi ? xop.ld : cols;
  const size_t total = rows * (size_t)ldo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ldo;
    const int c = (int)(i - r * ldo);
    float xn = 0.f;
    if (c < cols) {
      const size_t j = r * cols + c;
      xn = x[j];
      if (keep[j]) {
        float wg = __fadd_rn(__fmul_rn(gt_w, gt[j]), __fmul_rn(noise_w, noise[j]));
        const int t = (int)(r % (size_t)T);
        if (blend_w != nullptr && t < overlap) {
          const float lw = blend_w[t];
          wg = __fadd_rn(__fmul_rn(wg, __fsub_rn(1.f, lw)), __fmul_rn(xn, lw));
        }
        xn = wg;
        x[j] = xn;
      }
    }
    if (xop.hi) op_store1(xop, op_fmt, i, xn);
  }
}

// RePaint "undo" (gaussian_diffusion.py:426-435): x <- sqrt(1 - beta) x + sqrt(beta) noise ; rewrites the operand copy
__global__ void __launch_bounds__(256)
undo_kernel(float* __restrict__ x, const float* __restrict__ noise, size_t rows, int cols, float a, float b, OpPtr xop,
            int op_fmt) {
  pdl_trigger();
  pdl_wait();
  const int ldo = xop.hi ? xop.ld : cols;
  const size_t total = rows * (size_t)ldo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ldo;
    const int c = (int)(i - r * ldo);
    float xn = 0.f;
    if (c < cols) {
      const size_t j = r * cols + c;
      xn = __fadd_rn(__fmul_rn(a, x[j]), __fmul_rn(b, noise[j]));
      x[j] = xn;
    }
    if (xop.hi) op_store1(xop, op_fmt, i, xn);
  }
}

// Standard-normal noise generated on the device: Philox4x32-10 (Salmon et al., SC'11) keyed by `seed`, counter =
// (element quad index, stream id `sub`), Box-Muller on the four outputs.  Replaces the per-step torch.randn_like draws
// of the reference samplers (gaussian_diffusion.py:685, :847, :867, :432) when the caller supplies no explicit noise:
// one buffer of B*T*F floats is refilled per step instead of n_steps of them being allocated up front.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}
__global__ void __launch_bounds__(256)
randn_fill_kernel(float* __restrict__ out, size_t n, unsigned long long seed, unsigned long long sub) {
  pdl_trigger();
  pdl_wait();
  const size_t nq = (n + 3) / 4;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = ((float)(c[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);       // (0, 1]
      const float u2 = (float)(c[2 * h + 1] >> 8) * (1.0f / 16777216.0f);            // [0, 1)
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      z[2 * h] = rad * cs; z[2 * h + 1] = rad * sn;
    }
    const size_t i = q * 4;
    if (i + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      *reinterpret_cast<float4*>(out + i) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
      for (int e = 0; e < 4; ++e) if (i + e < n) out[i + e] = z[e];
    }
  }
}

__global__ void __launch_bounds__(256)
axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float wa, float wb, float* __restrict__ out, size_t n) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a[i], wa), __fmul_rn(b[i], wb));
}

// Static human-topology mixing of STMA (st_attention.py:123-128): out[r, h, :] = sum_l softmax(body_weight, dim=1)[h, l] *
// v[r, l, :] for the H (= 12) body parts of every token row r.  HBM-bound: one read and one write of the (rows, H, L) tensor;
// the H x H graph is soft-maxed once per block in shared memory.
__global__ void __launch_bounds__(256)
part_mix_kernel(const float* __restrict__ body_weight, const float* __restrict__ v, float* __restrict__ out, size_t rows,
                int H, int L, int pitch) {
  __shared__ float w[32 * 32];
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x < H) {
    float m = -INFINITY;
    for (int l = 0; l < H; ++l) m = fmaxf(m, body_weight[threadIdx.x * H + l]);
    float sum = 0.f;
    for (int l = 0; l < H; ++l) {
      const float e = expf(body_weight[threadIdx.x * H + l] - m);
      w[threadIdx.x * H + l] = e;
      sum += e;
    }
    for (int l = 0; l < H; ++l) w[threadIdx.x * H + l] /= sum;
  }
  __syncthreads();
  const size_t total = rows * (size_t)L;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / L;
    const int d = (int)(i - r * L);
    const float* src = v + r * (size_t)H * pitch + d;      // part l of row r starts at (r * H + l) * pitch
    float x[32];
    for (int l = 0; l < H; ++l) x[l] = src[(size_t)l * pitch];
    for (int h = 0; h < H; ++h) {
      float acc = 0.f;
      for (int l = 0; l < H; ++l) acc = fmaf(w[h * H + l], x[l], acc);
      out[r * (size_t)H * L + (size_t)h * L + d] = acc;
    }
  }
}

__global__ void fill_timesteps_kernel(long long* __restrict__ t_buf, long long t, int B) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) t_buf[i] = t;
}

// out[0..n) = src[idx[0] * n + (0..n)): one slice of a per-step table picked by a DEVICE index, so that the node can live in a
// CUDA graph that is replayed for every step
__global__ void __launch_bounds__(256)
gather_slice_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ out, size_t n4) {
  pdl_trigger();
  pdl_wait();
  const float4* s4 = reinterpret_cast<const float4*>(src) + (size_t)idx[0] * n4;
  float4* o4 = reinterpret_cast<float4*>(out);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) o4[i] = s4[i];
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  return (int)std::min<size_t>(g, 148 * 16);
}
}  // namespace

int fill_timesteps_launch(long long* t_buf, long long t, int B, cudaStream_t stream) {
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(fill_timesteps_kernel, dim3((B + 255) / 256), dim3(256), (size_t)(0), stream, t_buf, t, B));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int gather_slice_launch(const float* src, const long long* idx_dev, float* out, size_t n, cudaStream_t stream) {
  MCM_CHECK(n % 4 == 0, "gather_slice: slice length must be a multiple of 4");
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(gather_slice_kernel, dim3(grid_for(n / 4, 256)), dim3(256), (size_t)(0), stream, src, idx_dev, out, n / 4));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int axpby_launch(const float* a, const float* b, float wa, float wb, float* out, size_t n, cudaStream_t stream) {
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(axpby_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), stream, a, b, wa, wb, out, n));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int part_mix_launch(const float* body_weight, const float* v, float* out, size_t rows, int H, int L, cudaStream_t stream,
                    int pitch) {
  MCM_CHECK(H >= 1 && H <= 32 && L >= 1, "part_mix: at most 32 parts");
  if (pitch <= 0) pitch = L;
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(part_mix_kernel, dim3(grid_for(rows * (size_t)L, 256)), dim3(256), (size_t)(0), stream, body_weight, v, out,
                      rows, H, L, pitch));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int randn_fill_launch(float* out, size_t n, unsigned long long seed, unsigned long long sub, cudaStream_t stream) {
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(randn_fill_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), stream, out, n, seed, sub));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int elementwise_init() {
  MCM_CUDA(cudaFuncSetAttribute(ln_transpose_kernel<OP_F16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 33 * 4));
  MCM_CUDA(cudaFuncSetAttribute(ln_transpose_kernel<OP_BF16X2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 33 * 4));
  MCM_CUDA(cudaFuncSetAttribute(ln_transpose_kernel<OP_F16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 33 * 4));
  MCM_CUDA(cudaFuncSetAttribute(ln_transpose_kernel<OP_BF16X2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 33 * 4));
  MCM_CUDA(cudaFuncSetAttribute(ln_transpose16_kernel<OP_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 17 * 4));
  MCM_CUDA(cudaFuncSetAttribute(ln_transpose16_kernel<OP_BF16X2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 17 * 4));
  return 0;
}
unsigned long long elementwise_launch_count() { return g_ew_launches.load(); }
void elementwise_count_replayed(unsigned long long n) { g_ew_launches.fetch_add(n); }

int ln_rows_launch(const float* in, int rows, int d, int ld_in, const float* w, const float* b, const float* scale,
                   const float* shift, int mod_ld, int rows_per_batch, bool act_silu, OpPtr out, int out_fmt,
                   cudaStream_t stream) {
  MCM_CHECK(d % 4 == 0 && d <= 32 * 4 * MAXV_LIMIT, "ln_rows: row length must be a multiple of 4 and <= 1024");
  MCM_CHECK(ld_in % 4 == 0 && out.ld % 4 == 0 && out.ld >= d && out.ld <= 32 * 4 * MAXV_LIMIT, "ln_rows: bad pitch");
  MCM_CHECK(mod_ld % 4 == 0, "ln_rows: modulation pitch must be a multiple of 4");
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  const int rpb = rows_per_batch > 0 ? rows_per_batch : 1;
  const int nv = (out.ld / 4 + 31) / 32;      // float4 per lane needed to cover the (padded) row
  uint16_t* hi = reinterpret_cast<uint16_t*>(out.hi);
  uint16_t* lo = reinterpret_cast<uint16_t*>(out.lo);
  LaunchTimer lt(LK_ROW, stream);
#define MCM_LN_GO(V, F)                                                                                          \
  MCM_CUDA(launch_pdl(ln_rows_kernel<V, F>, dim3(grid), dim3(wpb * 32), (size_t)0, stream, in, rows, d, ld_in, w, b,   \
                      scale, shift, mod_ld, rpb, act_silu ? 1 : 0, hi, lo, out.ld))
#define MCM_LN_CASE(V)                                                                                           \
  case V:                                                                                                        \
    if (out_fmt == OP_F16) { MCM_LN_GO(V, OP_F16); } else { MCM_LN_GO(V, OP_BF16X2); }                           \
    break;
  switch (nv) {
    MCM_LN_CASE(1) MCM_LN_CASE(2) MCM_LN_CASE(3) MCM_LN_CASE(4) MCM_LN_CASE(5) MCM_LN_CASE(6) MCM_LN_CASE(7)
    default:
      if (out_fmt == OP_F16) { MCM_LN_GO(8, OP_F16); } else { MCM_LN_GO(8, OP_BF16X2); }
  }
#undef MCM_LN_CASE
#undef MCM_LN_GO
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int softmax_seg_launch(const float* in, int rows, int ncols, int ld_in, int seg, OpPtr out, int out_fmt,
                       cudaStream_t stream) {
  MCM_CHECK(seg > 0 && ncols % seg == 0 && seg <= 1024, "softmax_seg: segment must divide ncols and be <= 1024");
  MCM_CHECK(out.ld >= ncols, "softmax_seg: output pitch too small");
  const int nseg = ncols / seg;
  const long long total = (long long)rows * nseg;
  const int wpb = 8;
  uint16_t* hi = reinterpret_cast<uint16_t*>(out.hi);
  uint16_t* lo = reinterpret_cast<uint16_t*>(out.lo);
  LaunchTimer lt(LK_ROW, stream);
#define MCM_SM_GO(LPS, EPL, F)                                                                                   \
  {                                                                                                              \
    const long long warps = (total + (32 / LPS) - 1) / (32 / LPS);                                               \
    MCM_CUDA(launch_pdl(softmax_seg_kernel<LPS, EPL, F>, dim3((unsigned)((warps + wpb - 1) / wpb)), dim3(wpb * 32), (size_t)(0), stream, \
        in, total, ncols, ld_in, seg, nseg, hi, lo, out.ld));                                                     \
  }
#define MCM_SM_LAUNCH(LPS, EPL)                                                                                  \
  { if (out_fmt == OP_F16) MCM_SM_GO(LPS, EPL, OP_F16) else MCM_SM_GO(LPS, EPL, OP_BF16X2) }
  if (seg <= 16) MCM_SM_LAUNCH(8, 2)
  else if (seg <= 32) MCM_SM_LAUNCH(8, 4)
  else if (seg <= 64) MCM_SM_LAUNCH(8, 8)
  else if (seg < 128) MCM_SM_LAUNCH(8, 16)
  else if (seg <= 128) MCM_SM_LAUNCH(32, 4)
  else if (seg <= 256) MCM_SM_LAUNCH(32, 8)
  else if (seg <= 512) MCM_SM_LAUNCH(32, 16)
  else MCM_SM_LAUNCH(32, 32)
#undef MCM_SM_LAUNCH
#undef MCM_SM_GO
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int ln_transpose_launch(const float* h, int B, int T, int D, const float* w, const float* b, OpPtr out, int out_fmt,
                        cudaStream_t stream) {
  MCM_CHECK(D % 32 == 0 && T <= 1024 && out.ld >= T && out.ld % 2 == 0, "ln_transpose: need D % 32 == 0, T <= 1024");
  static const int wide_min = [] { const char* e = getenv("MCM_LNT_WIDE_MIN_T"); return e ? atoi(e) : 257; }();
  static const int narrow_min = [] { const char* e = getenv("MCM_LNT16_MIN_T"); return e ? atoi(e) : 513; }();
  const bool wide = T >= wide_min;
  dim3 grid(D / 32, B), block(32, wide ? 32 : 8);
  uint16_t* hi = reinterpret_cast<uint16_t*>(out.hi);
  uint16_t* lo = reinterpret_cast<uint16_t*>(out.lo);
  const size_t smem = (size_t)T * 33 * sizeof(float);
  LaunchTimer lt(LK_ROW, stream);
  if (T >= narrow_min) {
    const dim3 g16(D / 16, B), b16(32, 16);
    const size_t sm16 = (size_t)T * 17 * sizeof(float);
    if (out_fmt == OP_F16) MCM_CUDA(launch_pdl(ln_transpose16_kernel<OP_F16>, dim3(g16), dim3(b16), sm16, stream, h, T, D, w, b, hi, lo, out.ld));
    else MCM_CUDA(launch_pdl(ln_transpose16_kernel<OP_BF16X2>, dim3(g16), dim3(b16), sm16, stream, h, T, D, w, b, hi, lo, out.ld));
  } else if (out_fmt == OP_F16) {
    if (wide) MCM_CUDA(launch_pdl(ln_transpose_kernel<OP_F16, 32>, dim3(grid), dim3(block), smem, stream, h, T, D, w, b, hi, lo, out.ld));
    else MCM_CUDA(launch_pdl(ln_transpose_kernel<OP_F16, 8>, dim3(grid), dim3(block), smem, stream, h, T, D, w, b, hi, lo, out.ld));
  } else {
    if (wide) MCM_CUDA(launch_pdl(ln_transpose_kernel<OP_BF16X2, 32>, dim3(grid), dim3(block), smem, stream, h, T, D, w, b, hi, lo, out.ld));
    else MCM_CUDA(launch_pdl(ln_transpose_kernel<OP_BF16X2, 8>, dim3(grid), dim3(block), smem, stream, h, T, D, w, b, hi, lo, out.ld));
  }
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int pack_op_launch(const float* in, int rows, int cols, int ld_in, bool act_silu, OpPtr out, int out_fmt,
                   cudaStream_t stream) {
  MCM_CHECK(out.ld >= cols, "pack_op: output pitch too small");
  const size_t total = (size_t)rows * out.ld;
  LaunchTimer lt(LK_ROW, stream);
  const bool vec8 = cols % 8 == 0 && ld_in % 4 == 0 && out.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out.hi) & 15) == 0 && (out.lo == nullptr || (reinterpret_cast<uintptr_t>(out.lo) & 15) == 0);
  if (vec8)
    MCM_CUDA(launch_pdl(pack_op8_kernel, dim3(grid_for(total / 8, 256)), dim3(256), (size_t)0, stream, in, (size_t)rows, cols, ld_in, act_silu ? 1 : 0, out, out_fmt));
  else
  MCM_CUDA(launch_pdl(pack_op_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, stream, in, (size_t)rows, cols, ld_in, act_silu ? 1 : 0, out, out_fmt));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int timestep_embedding_launch(const long long* t_dev, int t_uniform, int B, int dim, OpPtr out, int out_fmt,
                              cudaStream_t stream) {
  const int total = B * out.ld;
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(timestep_embedding_kernel, dim3((total + 255) / 256), dim3(256), (size_t)(0), stream, t_dev, t_uniform, B, dim, out, out_fmt));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int ddim_update_launch(const float* x, const float* eps, const float* noise, float* x_out, size_t rows, int cols,
                       DdimCoefs c, OpPtr xop, int op_fmt, cudaStream_t stream) {
  MCM_CHECK(!c.add_noise || noise != nullptr, "ddim_update: eta != 0 needs step noise");
  const size_t total = rows * (size_t)(xop.hi ? xop.ld : cols);
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(ddim_update_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, stream, x, eps, noise, x_out, rows, cols, c, xop, op_fmt));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int repaint_blend_launch(float* x, const float* gt, const unsigned char* keep, const float* noise, size_t rows, int cols,
                         int T, float gt_w, float noise_w, const float* blend_w, int overlap, OpPtr xop, int op_fmt,
                         cudaStream_t stream) {
  MCM_CHECK(x && gt && keep && noise && T > 0, "repaint_blend: null argument");
  const size_t total = rows * (size_t)(xop.hi ? xop.ld : cols);
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(repaint_blend_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, stream, x, gt, keep, noise, rows, cols,
                      T, gt_w, noise_w, blend_w, overlap, xop, op_fmt));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int undo_launch(float* x, const float* noise, size_t rows, int cols, float a, float b, OpPtr xop, int op_fmt,
                cudaStream_t stream) {
  MCM_CHECK(x && noise, "undo: null argument");
  const size_t total = rows * (size_t)(xop.hi ? xop.ld : cols);
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(undo_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, stream, x, noise, rows, cols, a, b, xop, op_fmt));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

int ddpm_update_launch(const float* x, const float* eps, const float* noise, float* x_out, size_t rows, int cols,
                       DdpmCoefs c, OpPtr xop, int op_fmt, cudaStream_t stream) {
  MCM_CHECK(!c.add_noise || noise != nullptr, "ddpm_update: needs step noise");
  const size_t total = rows * (size_t)(xop.hi ? xop.ld : cols);
  LaunchTimer lt(LK_ROW, stream);
  MCM_CUDA(launch_pdl(ddpm_update_kernel, dim3(grid_for(total, 256)), dim3(256), (size_t)0, stream, x, eps, noise, x_out, rows, cols, c, xop, op_fmt));
  MCM_CUDA(cudaGetLastError());
  g_ew_launches.fetch_add(1);
  return 0;
}

}  // namespace mcm
