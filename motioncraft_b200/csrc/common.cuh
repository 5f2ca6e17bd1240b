// motioncraft_b200 -- shared device helpers: error plumbing, mbarrier / TMA / tcgen05 PTX wrappers,
// 16-bit operand packing.  sm_100a only (tcgen05 + TMEM + TMA); there is no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace mcm {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing: every C-ABI entry returns an int status and records a message that
// mcm_last_error() hands back; nothing throws across the ABI and nothing calls exit().
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
std::string mcm_last_error_string();
#define MCM_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::mcm::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                       ":" + std::to_string(__LINE__) + ")");                                       \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)
#define MCM_CHECK(cond, msg)                                                             \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ::mcm::set_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" +            \
                       std::to_string(__LINE__) + ")");                                  \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)
#define MCM_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != 0) return _s;    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// operand formats.  A GEMM operand lives in HBM as 16-bit K-major rows (row pitch `ld` elements,
// ld % 8 == 0 so TMA's 16-byte pitch rule holds, pad columns zero):
//   OP_F16   : one fp16 tensor (11 significant bits, the fast 1-pass mode)
//   OP_BF16X2: hi + lo bf16 tensors (hi = bf16(v), lo = bf16(v - hi), 16 significant bits); the
//              GEMM issues hi*hi + hi*lo + lo*hi (3 passes) -- used where fp16 rounding would eat the
//              1e-3 parity budget (embed, out, AdaLN emb_layers; DESIGN.md "precision")
// ---------------------------------------------------------------------------------------------
enum OpFormat : int { OP_F16 = 0, OP_BF16X2 = 1 };

struct OpPtr {       // a 16-bit operand tensor (hi, optional lo)
  void* hi;
  void* lo;
  int ld;            // row pitch in elements
};

#ifdef __CUDACC__
__device__ __forceinline__ uint16_t f32_to_f16_bits(float v) {
  // saturating (one F2FP.SATFINITE): fp16 has a 65504 ceiling; anything larger would otherwise turn into inf -> NaN
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void f32_to_bf16x2_bits(float v, uint16_t& hi, uint16_t& lo) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}
// store one operand element
__device__ __forceinline__ void op_store1(const OpPtr& o, int fmt, size_t idx, float v) {
  if (fmt == OP_F16) {
    reinterpret_cast<uint16_t*>(o.hi)[idx] = f32_to_f16_bits(v);
  } else {
    uint16_t h, l;
    f32_to_bf16x2_bits(v, h, l);
    reinterpret_cast<uint16_t*>(o.hi)[idx] = h;
    reinterpret_cast<uint16_t*>(o.lo)[idx] = l;
  }
}
// store 8 consecutive operand elements (idx % 8 == 0, 16-byte aligned)
__device__ __forceinline__ void op_store8(const OpPtr& o, int fmt, size_t idx, const float* v) {
  if (fmt == OP_F16) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = (uint32_t)f32_to_f16_bits(v[2 * i]) | ((uint32_t)f32_to_f16_bits(v[2 * i + 1]) << 16);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o.hi) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint16_t h0, l0, h1, l1;
      f32_to_bf16x2_bits(v[2 * i], h0, l0);
      f32_to_bf16x2_bits(v[2 * i + 1], h1, l1);
      wh[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
      wl[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o.hi) + idx) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o.lo) + idx) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error the host reports), never
// as a hung GPU.  try_wait suspends in hardware for a while per call, so the bound is seconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 22)) __trap();
  }
}
// Non-blocking probe of a barrier phase (1 = the phase with this parity has completed).  Issued one ring stage AHEAD of its use,
// the ~300-cycle round trip of the barrier unit overlaps the instructions in between instead of heading every k-block.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// two probes in flight at once (their round trips overlap)
__device__ __forceinline__ void mbar_test2(uint32_t bar0, uint32_t par0, uint32_t bar1, uint32_t par1, uint32_t& d0, uint32_t& d1) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "selp.b32 %1, 1, 0, q;\n\t}"
      : "=r"(d0), "=r"(d1)
      : "r"(bar0), "r"(par0), "r"(bar1), "r"(par1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (fp16 or bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete (implies fence::before)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// [0,14) addr>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024 B between
// 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Programmatic dependent launch: every kernel of this library is launched with the programmatic-stream-serialization
// attribute.  pdl_trigger() (first statement) lets the NEXT kernel of the stream be scheduled while this one still
// runs; pdl_wait() blocks until the PREVIOUS kernel has completed and its writes are visible, and must precede the
// first global-memory access.  What overlaps is the launch latency and the memory-free prologue of each kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

#ifdef __CUDACC__
// host-side launch with the PDL attribute (MCM_PDL=0 disables it)
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

}  // namespace mcm
