// motioncraft_b200 -- tcgen05 / TMEM / TMA GEMM (sm_100a).  See gemm_tc.cuh for the problem model.
//
// One persistent CTA per SM, 6 warps, warp-specialised:
//   warp 0      TMA producer   : cp.async.bulk.tensor (128B-swizzled boxes) -> smem ring, mbarrier tx
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM);
//                                tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2..5  epilogue       : tcgen05.ld (one TMEM lane = one output row per thread), bias / addend /
//                                activation / mask, fp32 and/or 16-bit operand stores
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include "gemm_tc.cuh"
#include "ptx_extra.cuh"
#include "timing.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace mcm {

namespace {
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;    // 64 x 16-bit = 128 B = one swizzle-128B atom row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 8;
constexpr int NUM_EPI_WARPS = 8;          // multiple of 4 (TMEM lane quadrants)
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;
constexpr int SMEM_TOTAL = 230400;         // dynamic shared memory request (+ ~1.3 KB static <= 227 KB)
// epilogue staging per warp: 4 KB (an fp32 tile, or 16-bit hi [+ lo]) or 8 KB when a segment writes fp32 AND an operand

struct KParams {
  int M, M_pad, K, batches, inner, a_k_inner, b_batched, out_col_inner, out_rows_per_outer,
      trans_rows, head_dim, b_k_inner, a_batched, out_batched, bias_inner;
  int block_n, m_tiles, n_tiles, num_kb, stages, split, total_tiles;
  int stg_bytes;        // per-warp epilogue staging bytes (4096 or 8192)
  int pair;             // 1: cta_group::2 -- a CTA pair computes a 256 x block_n tile, each CTA holds half of B
  int cs, m_supers;     // cluster size (CTAs sharing one multicast B tile) and ceil(m_tiles / cs)
  uint32_t idesc, stage_bytes, a_bytes, b_bytes;
  int nseg;
  EpiSeg seg[3];
  unsigned long long* dbg;   // MCM_DEBUG_EPI=3: per-phase clock totals of the epilogue warps
  int prefetch;         // producer issues L2 prefetches for the next work item's A tile
  int debug;            // MCM_DEBUG_EPI: 1 = skip staging + stores, 2 = also skip the TMEM read (timing experiments only)
  unsigned long long* trace;   // MCM_GEMM_TRACE=<file>: stamps of CTA 0 of this launch (8 x globaltimer ns, 8 x clock64, then
                               // clock64 per k-block: [16, 32) MMA warp after its full-barrier wait, [32, 48) producer after issue)
  int tma_mode[3];      // per segment: 0 = generic epilogue, else bit0 TMA epilogue, bit1 residual via TMA reduce-add,
                        // bit2 addend TMA-loaded, bit3 addend broadcast over batches
};

struct EpiMaps {        // [segment][0 = fp32 out, 1 = fp32 addend, 2 = op hi, 3 = op lo]
  CUtensorMap m[3][4];
};

enum { TM_TMA = 1, TM_RED = 2, TM_LDADD = 4, TM_BCAST = 8 };

template <bool WITH_GENERIC, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const __grid_constant__ EpiMaps em, const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t abar[NUM_EPI_WARPS];   // addend-tile arrival, one per epilogue warp
  __shared__ __align__(16) float bias_s[NUM_EPI_WARPS][32];

  pdl_trigger();                                      // the next kernel of the stream may start its prologue
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
#define MCM_STAMP(i)                                                                                     \
  do {                                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) {                                            \
      unsigned long long _g;                                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_g));                                             \
      p.trace[i] = _g;                                                                                   \
      p.trace[8 + (i)] = (unsigned long long)clock64();                                                  \
    }                                                                                                    \
  } while (0)
  if (warp == 0) MCM_STAMP(0);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.split) {
      tma_prefetch_desc(&tmAlo);
      tma_prefetch_desc(&tmBlo);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        // multicast mode: one tcgen05.commit arrival from every CTA of the cluster; pair mode: one (the leader's)
        mbar_init(smem_u32(&empty_bar[s]), PAIR ? 1u : (uint32_t)p.cs);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&tfull_bar[s]), 1);
        mbar_init(smem_u32(&tempty_bar[s]), PAIR ? 2u * NUM_EPI_WARPS : (uint32_t)NUM_EPI_WARPS);   // per epilogue warp (of both CTAs)
      }
      for (int s = 0; s < NUM_EPI_WARPS; ++s) mbar_init(smem_u32(&abar[s]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (PAIR) tmem_alloc_2sm(smem_u32(&tmem_slot), TMEM_COLS);
    else tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();     // peers' barriers are initialised before anyone multicasts / arrives remotely
  tc_fence_after();
  // griddepcontrol.wait is issued by each role as LATE as possible (with programmatic dependent launch the kernel typically
  // starts tens of microseconds before its inputs exist): the producer right before its first TMA load, after the tile
  // decode and every parameter fetch -- cold constant-bank / instruction-cache misses cost ~2.5 k cycles there, which the
  // time line showed between the wait and the first load; the epilogue warps after their first accumulator arrives (their
  // bias prefetch reads parameters only); the MMA warp touches no global memory at all.
  const uint32_t tmem_base = tmem_slot;
  const int crank = p.cs > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / p.cs;
  const int n_clusters = gridDim.x / p.cs;
  const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The WHOLE warp runs the loop (so addresses / coordinates stay warp-uniform and live in uniform registers);
    // only the issuing instructions are predicated on one elected lane.  A divergent `if (lane == 0)` around the
    // loop makes the compiler wrap every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall.
    {
      int stage = 0;
      uint32_t phase = 0;
      bool pdl_done = false;
      long long pr_wait = 0, pr_kb = 0, pr_t0 = (p.debug == 3) ? clock64() : 0;
      const uint32_t tx = p.stage_bytes;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const int n_tile = tile % p.n_tiles;
        const int t2 = tile / p.n_tiles;
        const int m_blk = (t2 % p.m_supers) * p.cs + crank;
        const int batch = t2 / p.m_supers;
        const int outer = batch / p.inner;
        const int inner = batch - outer * p.inner;
        int si = 0;
        while (si + 1 < p.nseg && n_tile >= p.seg[si + 1].tile0) ++si;
        const int b_row = p.seg[si].w_row0 + (n_tile - p.seg[si].tile0) * p.block_n;
        const int a_k0 = inner * p.a_k_inner;
        const int b_k0 = inner * p.b_k_inner;
        const int a_z = p.a_batched ? batch : outer;
        const int b_z = p.b_batched ? batch : 0;
        if (p.prefetch && !pdl_done) {           // (experiment switch) the prefetches below read activations
          pdl_wait();
          pdl_done = true;
        }
        if (p.prefetch && lane == 0) {
          // pull the NEXT work item's A rows (activations, usually DRAM-resident) into L2 while this one is loaded
          const int nt = tile + n_clusters;
          if (nt < p.total_tiles) {
            const int nt2 = nt / p.n_tiles;
            const int nm = (nt2 % p.m_supers) * p.cs + crank;
            const int nbatch = nt2 / p.m_supers;
            const int nouter = nbatch / p.inner;
            const int nk0 = (nbatch - nouter * p.inner) * p.a_k_inner;
            const int nz = p.a_batched ? nbatch : nouter;
            for (int kb = 0; kb < p.num_kb; ++kb) {
              tma_prefetch_3d(&tmA, nk0 + kb * BLOCK_K, nm * BLOCK_M, nz);
              if (p.split) tma_prefetch_3d(&tmAlo, nk0 + kb * BLOCK_K, nm * BLOCK_M, nz);
            }
          }
        }
        // Everything that does not depend on kb is computed here: the k loop is ONE warp issuing dependent scalar instructions
        // (measured with MCM_GEMM_TRACE: 710 cycles per k-block with the address / mode arithmetic inside the loop).
        const int num_kb = p.num_kb, stages = p.stages;
        const uint32_t stage_bytes = p.stage_bytes;
        const int a_row = m_blk * BLOCK_M;
        const bool b_mc = p.cs > 1;
        const int b_row_l = PAIR ? b_row + crank * (p.block_n / 2) : (b_mc ? b_row + crank * (p.block_n / p.cs) : b_row);
        const uint32_t b_dst = (p.split ? 2u * p.a_bytes : p.a_bytes) +
                               ((!PAIR && b_mc) ? (uint32_t)(crank * (p.block_n / p.cs)) * (BLOCK_K * 2) : 0u);
        const uint32_t alo_dst = p.a_bytes, blo_dst = b_dst + p.b_bytes;
        const bool split = p.split != 0;
        const bool dbg3 = p.debug == 3;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && tile == cluster_id;
        auto load_kb = [&](int st, int kb) {
          const uint32_t bar = full0 + (uint32_t)st * 8u;
          const uint32_t sa = smem_base + (uint32_t)st * stage_bytes;
          const int ka = a_k0 + kb * BLOCK_K, kbk = b_k0 + kb * BLOCK_K;
          if (elect_one()) {
            if constexpr (PAIR) {
              // CTA pair: each CTA loads its own A rows and its half of the B rows into its OWN smem; all bytes are
              // accounted on the LEADER's barrier, which its MMA thread waits on.
              if (crank == 0) mbar_expect_tx(bar, 2u * tx);
              tma_load_3d_2sm(&tmA, bar, sa, ka, a_row, a_z);
              if (split) tma_load_3d_2sm(&tmAlo, bar, sa + alo_dst, ka, a_row, a_z);
              tma_load_3d_2sm(&tmB, bar, sa + b_dst, kbk, b_row_l, b_z);
              if (split) tma_load_3d_2sm(&tmBlo, bar, sa + blo_dst, kbk, b_row_l, b_z);
            } else {
              // A: this CTA's own 128 rows.  B: this CTA fetches rows [crank, crank+1) * block_n / cs of the tile and
              // multicasts them to every CTA of the cluster (each CTA's barrier counts the whole tile's bytes).
              mbar_expect_tx(bar, tx);
              tma_load_3d(&tmA, bar, sa, ka, a_row, a_z);
              if (split) tma_load_3d(&tmAlo, bar, sa + alo_dst, ka, a_row, a_z);
              if (b_mc) {
                tma_load_3d_mc(&tmB, bar, sa + b_dst, kbk, b_row_l, b_z, cmask);
                if (split) tma_load_3d_mc(&tmBlo, bar, sa + blo_dst, kbk, b_row_l, b_z, cmask);
              } else {
                tma_load_3d(&tmB, bar, sa + b_dst, kbk, b_row_l, b_z);
                if (split) tma_load_3d(&tmBlo, bar, sa + blo_dst, kbk, b_row_l, b_z);
              }
              if (tr && kb < 16) p.trace[32 + kb] = (unsigned long long)clock64();
            }
          }   // elect_one
        };
        // two k-blocks per trip, their empty-barrier probes in flight together (see the MMA warp; four per trip measured slower)
        for (int kb = 0; kb < num_kb; kb += 2) {
          long long tp0 = 0;
          if (dbg3) tp0 = clock64();
          const bool two = kb + 1 < num_kb;
          const int s0 = stage;
          const uint32_t ph0 = phase;
          const int s1 = (s0 + 1 == stages) ? 0 : s0 + 1;
          const uint32_t ph1 = (s0 + 1 == stages) ? ph0 ^ 1u : ph0;
          uint32_t r0, r1;
          mbar_test2(empty0 + (uint32_t)s0 * 8u, ph0 ^ 1u, empty0 + (uint32_t)s1 * 8u, ph1 ^ 1u, r0, r1);
          if (stages < 2) r1 = 0u;
          if (!r0) mbar_wait(empty0 + (uint32_t)s0 * 8u, ph0 ^ 1u);
          if (dbg3) { const long long n = clock64(); pr_wait += n - tp0; pr_kb += two ? 2 : 1; }
          if (!pdl_done) {                       // first load of this CTA: the inputs must exist from here on
            pdl_wait();
            pdl_done = true;
            MCM_STAMP(1);
          }
          load_kb(s0, kb);
          if (two) {
            if (!r1) {
              const long long tp1 = dbg3 ? clock64() : 0;
              mbar_wait(empty0 + (uint32_t)s1 * 8u, ph1 ^ 1u);
              if (dbg3) pr_wait += clock64() - tp1;
            }
            load_kb(s1, kb + 1);
          }
          __syncwarp();
          stage = two ? ((s1 + 1 == stages) ? 0 : s1 + 1) : s1;
          phase = two ? ((s1 + 1 == stages) ? ph1 ^ 1u : ph1) : ph1;
        }
      }
      if (p.debug == 3 && p.dbg != nullptr && lane == 0) {
        atomicAdd(p.dbg + 8, (unsigned long long)pr_wait);
        atomicAdd(p.dbg + 9, (unsigned long long)(clock64() - pr_t0));
        atomicAdd(p.dbg + 10, (unsigned long long)pr_kb);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair mode: the leader CTA only)
    if (!(PAIR && crank != 0)) {      // whole warp, converged; tcgen05.mma / commit issued by one elected lane
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long mm_wacc = 0, mm_wfull = 0, mm_t0 = (p.debug == 3) ? clock64() : 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        long long tm0 = (p.debug == 3) ? clock64() : 0;
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
        if (p.debug == 3) mm_wacc += clock64() - tm0;
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(acc * ACC_STRIDE);
        // Descriptors by arithmetic: the low word of a swizzle-128B descriptor is (address >> 4) | const and shared-memory
        // addresses stay below 2^18, so the descriptor of (stage, k) is the stage-0 descriptor plus (offset >> 4).  The loop is one
        // warp of dependent scalar instructions: per k-block it must stay well under the tensor time of the four MMAs.
        const int num_kb = p.num_kb, stages = p.stages;
        const uint32_t idesc = p.idesc;
        const uint64_t dA0 = make_smem_desc_sw128(smem_base);
        const uint64_t st_step = (uint64_t)(p.stage_bytes >> 4);
        const uint64_t offAlo = (uint64_t)(p.a_bytes >> 4);
        const uint64_t offB = (uint64_t)((p.split ? 2u * p.a_bytes : p.a_bytes) >> 4);
        const uint64_t offBlo = offB + (uint64_t)(p.b_bytes >> 4);
        const bool split = p.split != 0, dbg3 = p.debug == 3, mc = p.cs > 1;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && it == 0;
        // One k-block = the MMAs of ring stage `st` + the commit that hands the stage back to the producer.
        auto issue_kb = [&](int st, int kb, int kleft) {
          const int nk = kleft >= BLOCK_K ? BLOCK_K / UMMA_K : (kleft + UMMA_K - 1) / UMMA_K;
          const uint64_t dA = dA0 + (uint64_t)st * st_step;
          if (elect_one()) {
            if (!split) {
              const uint64_t dB = dA + offB;
              if (nk == BLOCK_K / UMMA_K) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  if constexpr (PAIR) umma_f16_2sm(taddr, dA + 2 * k, dB + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                  else umma_f16(taddr, dA + 2 * k, dB + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                }
              } else {
                for (int k = 0; k < nk; ++k) {
                  if constexpr (PAIR) umma_f16_2sm(taddr, dA + 2 * k, dB + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                  else umma_f16(taddr, dA + 2 * k, dB + 2 * k, idesc, (uint32_t)((kb | k) != 0));
                }
              }
            } else {
              for (int k = 0; k < nk; ++k) {
                const uint64_t ah = dA + 2 * k, al = ah + offAlo, bh = ah + offB, bl = ah + offBlo;
                if constexpr (PAIR) {
                  umma_f16_2sm(taddr, al, bh, idesc, (uint32_t)((kb | k) != 0));   // small terms first
                  umma_f16_2sm(taddr, ah, bl, idesc, 1u);
                  umma_f16_2sm(taddr, ah, bh, idesc, 1u);
                } else {
                  umma_f16(taddr, al, bh, idesc, (uint32_t)((kb | k) != 0));
                  umma_f16(taddr, ah, bl, idesc, 1u);
                  umma_f16(taddr, ah, bh, idesc, 1u);
                }
              }
            }
            // release the stage: pair mode in BOTH CTAs; multicast mode on every CTA of the cluster (they all write into it)
            if constexpr (PAIR) umma_commit_2sm(empty0 + (uint32_t)st * 8u, (uint16_t)3);
            else if (mc) umma_commit_mc(empty0 + (uint32_t)st * 8u, cmask);
            else umma_commit(empty0 + (uint32_t)st * 8u);
          }   // elect_one
        };
        // Two k-blocks per trip: the barrier probes of both stages are in flight together (a probe's round trip through the
        // barrier unit is ~230 cycles even when the phase completed long ago; per k-block the serial chain probe -> fence ->
        // 4 MMA issues -> commit -> reconverge measured 700+ cycles against 512 cycles of tensor time at N = 256).
        int kleft = p.K;
        for (int kb = 0; kb < num_kb; kb += 2, kleft -= 2 * BLOCK_K) {
          long long tf0 = dbg3 ? clock64() : 0;
          const bool two = kb + 1 < num_kb;
          const int s0 = stage;
          const uint32_t ph0 = phase;
          const int s1 = (s0 + 1 == stages) ? 0 : s0 + 1;
          const uint32_t ph1 = (s0 + 1 == stages) ? ph0 ^ 1u : ph0;
          uint32_t r0, r1;
          mbar_test2(full0 + (uint32_t)s0 * 8u, ph0, full0 + (uint32_t)s1 * 8u, ph1, r0, r1);
          if (stages < 2) r1 = 0u;            // a one-stage ring wraps inside the trip: that probe's parity would be ambiguous
          if (!r0) mbar_wait(full0 + (uint32_t)s0 * 8u, ph0);
          if (dbg3) mm_wfull += clock64() - tf0;
          if (it == 0 && kb == 0) MCM_STAMP(2);
          if (tr && kb < 16 && lane == 0) p.trace[16 + kb] = (unsigned long long)clock64();
          tc_fence_after();
          issue_kb(s0, kb, kleft);
          if (two) {
            if (!r1) {
              const long long tf1 = dbg3 ? clock64() : 0;
              mbar_wait(full0 + (uint32_t)s1 * 8u, ph1);
              if (dbg3) mm_wfull += clock64() - tf1;
              tc_fence_after();
            }
            if (tr && kb + 1 < 16 && lane == 0) p.trace[16 + kb + 1] = (unsigned long long)clock64();
            issue_kb(s1, kb + 1, kleft - BLOCK_K);
          }
          __syncwarp();
          stage = two ? ((s1 + 1 == stages) ? 0 : s1 + 1) : s1;
          phase = two ? ((s1 + 1 == stages) ? ph1 ^ 1u : ph1) : ph1;
        }
        if (elect_one()) {
          if constexpr (PAIR) umma_commit_2sm(smem_u32(&tfull_bar[acc]), (uint16_t)3);   // both CTAs' epilogues
          else umma_commit(smem_u32(&tfull_bar[acc]));       // accumulator complete
        }
        __syncwarp();
      }
      MCM_STAMP(3);
      if (p.debug == 3 && p.dbg != nullptr && lane == 0) {
        atomicAdd(p.dbg + 11, (unsigned long long)mm_wacc);
        atomicAdd(p.dbg + 12, (unsigned long long)mm_wfull);
        atomicAdd(p.dbg + 13, (unsigned long long)(clock64() - mm_t0));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2 .. 2+NUM_EPI_WARPS)
    // TMEM lane quadrant q is readable only by warps with warp % 4 == q; the NUM_EPI_WARPS / 4 warps that
    // share a quadrant take alternate 32-column chunks of the tile.
    const int ew = warp - 2;                          // 0 .. NUM_EPI_WARPS-1
    const int quad = warp & 3;
    const int chunk_phase = ew >> 2;                  // which of the interleaved chunk sets this warp takes
    // per-warp staging region behind the operand ring (1024-byte aligned: TMA swizzle patterns are
    // functions of the absolute shared-memory address)
    const uint32_t stg_base = smem_base + (uint32_t)p.stages * p.stage_bytes + (uint32_t)(ew * p.stg_bytes);
    float* stg = reinterpret_cast<float*>(smem_raw + (stg_base - smem_u32(smem_raw)));   // generic path: [32][33] floats
    const uint32_t my_abar = smem_u32(&abar[ew]);
    uint32_t abar_phase = 0;
    int pending_groups = 0;      // bulk-store groups of this warp that may still be reading the staging region
    long long tph[6] = {0, 0, 0, 0, 0, 0};   // debug: cycles in [wait acc | wait staging | tmem ld | math | stage+issue | chunks]
    const bool prof = p.debug == 3;
#define MCM_TICK(i) do { if (prof) { const long long _n = clock64(); tph[i] += _n - tlast; tlast = _n; } } while (0)
    long long tlast = prof ? clock64() : 0;
    int it = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int n_tile = tile % p.n_tiles;
      const int t2 = tile / p.n_tiles;
      const int m_blk = (t2 % p.m_supers) * p.cs + crank;
      const int batch = t2 / p.m_supers;
      const int outer_i = batch / p.inner;
      const int inner = batch - outer_i * p.inner;
      const int outer = p.out_batched ? batch : outer_i;       // z coordinate / row block of the OUTPUT
      int si = 0;
      while (si + 1 < p.nseg && n_tile >= p.seg[si + 1].tile0) ++si;
      const EpiSeg& sg = p.seg[si];
      const float* const bias_p = sg.bias != nullptr ? sg.bias + (size_t)inner * p.bias_inner : nullptr;
      const int nbase = (n_tile - sg.tile0) * p.block_n;
      const int row0 = m_blk * BLOCK_M + quad * 32;             // first row (within the batch) of this warp
      const int r = row0 + lane;                                // this thread's row in the TMEM layout
      const bool has_op = sg.op.hi != nullptr;
      const int flags = sg.flags;
      const int op_fmt = sg.op_fmt;

      // bias of this warp's (up to 4) chunks, fetched before waiting for the accumulator
      constexpr int CSTRIDE = 32 * (NUM_EPI_WARPS / 4);
      float bias_pre[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cb = nbase + chunk_phase * 32 + k * CSTRIDE + lane;
        bias_pre[k] = (bias_p != nullptr && chunk_phase * 32 + k * CSTRIDE + lane < p.block_n && cb < sg.n) ? __ldg(bias_p + cb) : 0.f;
      }
      const int tmode_t = p.tma_mode[si];
      // staging layout: a chunk that needs <= 4 KB (fp32 only, or 16-bit only) alternates between two 4 KB slots
      const bool dbl = !(sg.out32 != nullptr && has_op);

      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      if (it == 0) pdl_wait();                 // returns at once: the accumulator exists, so the producer's wait has passed
      if (ew == 0 && it == 0) MCM_STAMP(4);
      MCM_TICK(0);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);

      int kchunk = -1;
      for (int c0 = chunk_phase * 32; c0 < p.block_n; c0 += CSTRIDE) {
        ++kchunk;
        const int cbase = nbase + c0;                 // column within the segment of v[0]
        if (cbase >= sg.n_pad) break;                 // warp-uniform
        const int ncover = min(32, p.block_n - c0);   // columns of this chunk that belong to this tile
        const int nvalid = min(ncover, sg.n - cbase); // may be <= 0 (pad-only chunk)
        const int colg0 = sg.col0 + inner * p.out_col_inner + cbase;
        const int tmode = tmode_t;
        float v[32];

        // The staging slot may still be read by an earlier bulk store of this warp.
        if (pending_groups > 0) {
          if (lane == 0) tma_wait_read0();
          __syncwarp();
          pending_groups = 0;
        }
        MCM_TICK(1);
        const uint32_t stgA = stg_base;
        const uint32_t stgH = dbl ? stg_base : stg_base + 4096u;          // `dbl`: the segment has ONE kind of output
        const uint32_t stgL = dbl ? stg_base + 2048u : stg_base + 6144u;
        if (tmode & TM_TMA) {
          // ================= TMA epilogue: registers -> swizzled smem tile -> bulk tensor store =================
          if (p.debug == 2) continue;
          if ((tmode & TM_LDADD) && elect_one()) {     // fetch the addend tile while the accumulator is read
            mbar_expect_tx(my_abar, 4096);
            tma_load_3d(&em.m[si][1], my_abar, stgA, colg0, row0, (tmode & TM_BCAST) ? 0 : outer);
          }
          tmem_ld_32x32(taddr + (uint32_t)c0, v);
          bias_s[ew][lane] = kchunk == 0 ? bias_pre[0] : (kchunk == 1 ? bias_pre[1] : (kchunk == 2 ? bias_pre[2] : bias_pre[3]));
          __syncwarp();
          tmem_ld_wait();
          MCM_TICK(2);
          const bool transposed = (flags & EPI_TRANSPOSED) != 0;
          if (sg.bias != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[ew][4 * q]);
              v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
            }
          }
          if (tmode & TM_LDADD) {
            mbar_wait(my_abar, abar_phase);
            abar_phase ^= 1u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a4 = ld_shared_v4(stgA + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4));
              v[4 * q] += a4.x; v[4 * q + 1] += a4.y; v[4 * q + 2] += a4.z; v[4 * q + 3] += a4.w;
            }
          }
          if (flags & EPI_GELU) {
            if (op_fmt == OP_F16) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            }
          } else if (flags & EPI_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
          }
          if (flags & EPI_MASK_BLOCKDIAG) {
            const int rh = r / p.head_dim;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((cbase + j) / p.head_dim != rh) v[j] = 0.f;
          }
          if (nvalid < 32 || r >= p.M) {               // pad columns / pad rows of an operand must be exactly zero
            const bool rv = r < p.M;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j >= nvalid || !rv) v[j] = 0.f;
          }
          __syncwarp();                                // every lane is done reading the addend tile
          MCM_TICK(3);
          if (p.debug == 1) {
            if (v[0] == 1.2345e-30f && v[17] == 3.3e-33f) bias_s[ew][lane] = v[5];   // keep the loads alive
            continue;
          }
          if (!transposed) {
            if (sg.out32 != nullptr) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                st_shared_v4(stgA + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4), v[4 * q], v[4 * q + 1],
                             v[4 * q + 2], v[4 * q + 3]);
            }
            if (has_op) {
              const uint32_t sw = (uint32_t)((lane >> 1) & 3);
              if (op_fmt == OP_F16) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  st_shared_v4u(stgH + (uint32_t)lane * 64u + (((uint32_t)q ^ sw) << 4), pack_f16x2_sat(v[8 * q], v[8 * q + 1]),
                                pack_f16x2_sat(v[8 * q + 2], v[8 * q + 3]), pack_f16x2_sat(v[8 * q + 4], v[8 * q + 5]),
                                pack_f16x2_sat(v[8 * q + 6], v[8 * q + 7]));
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t h[4], l[4];
#pragma unroll
                  for (int e2 = 0; e2 < 4; ++e2) {
                    const float a0 = v[8 * q + 2 * e2], a1 = v[8 * q + 2 * e2 + 1];
                    h[e2] = pack_bf16x2(a0, a1);
                    l[e2] = pack_bf16x2(a0 - __uint_as_float(h[e2] << 16), a1 - __uint_as_float(h[e2] & 0xffff0000u));
                  }
                  st_shared_v4u(stgH + (uint32_t)lane * 64u + (((uint32_t)q ^ sw) << 4), h[0], h[1], h[2], h[3]);
                  st_shared_v4u(stgL + (uint32_t)lane * 64u + (((uint32_t)q ^ sw) << 4), l[0], l[1], l[2], l[3]);
                }
              }
            }
          } else {
            // transposed tile in smem: [c][r] dense rows (no swizzle), lanes (= r) contiguous
            if (sg.out32 != nullptr) {
              float* a = reinterpret_cast<float*>(smem_raw + (stgA - smem_u32(smem_raw)));
#pragma unroll
              for (int j = 0; j < 32; ++j) a[j * 32 + lane] = v[j];
            }
            if (has_op) {
              uint16_t* hh = reinterpret_cast<uint16_t*>(smem_raw + (stgH - smem_u32(smem_raw)));
              uint16_t* ll = reinterpret_cast<uint16_t*>(smem_raw + (stgL - smem_u32(smem_raw)));
              if (op_fmt == OP_F16) {
#pragma unroll
                for (int j = 0; j < 32; ++j) hh[j * 32 + lane] = f32_to_f16_bits(v[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  uint16_t h16, l16;
                  f32_to_bf16x2_bits(v[j], h16, l16);
                  hh[j * 32 + lane] = h16;
                  ll[j * 32 + lane] = l16;
                }
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (elect_one()) {
            const int x0 = transposed ? row0 : colg0;
            const int x1 = transposed ? colg0 : row0;
            if (sg.out32 != nullptr) {
              if (tmode & TM_RED) tma_reduce_add_3d(&em.m[si][0], stgA, x0, x1, outer);
              else tma_store_3d(&em.m[si][0], stgA, x0, x1, outer);
            }
            if (has_op) {
              tma_store_3d(&em.m[si][2], stgH, x0, x1, outer);
              if (op_fmt != OP_F16) tma_store_3d(&em.m[si][3], stgL, x0, x1, outer);
            }
            tma_commit();
          }
          ++pending_groups;
          MCM_TICK(4);
          if (prof) tph[5] += 1;
          continue;
        }
        if (!WITH_GENERIC) continue;   // (never reached: the host launches the WITH_GENERIC instantiation if needed)

        tmem_ld_32x32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!(flags & EPI_TRANSPOSED)) {
          // ---- row-major destination: transpose the 32x32 block through smem so that a warp
          // instruction touches ONE row (128 contiguous bytes) instead of 32 rows.
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = stg[i * 33 + lane];     // now: lane = column, i = row
          __syncwarp();
          const int nrows = min(32, p.M - row0);      // valid rows of this warp (<= 0: none)
          const bool col_ok = lane < nvalid;
          const bool col_pad = !col_ok && lane < ncover && (cbase + lane) < sg.n_pad;
          const float bias_v = (bias_p != nullptr && col_ok) ? __ldg(bias_p + cbase + lane) : 0.f;
          const size_t rowg0 = (size_t)outer * p.out_rows_per_outer + row0;
          const int colg = colg0 + lane;
          if (sg.addend != nullptr && col_ok) {
            const float* ap = sg.addend + ((flags & EPI_ADDEND_BCAST) ? (size_t)row0 : rowg0) * (size_t)sg.ld32 + colg;
            float a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = (i < nrows) ? ap[(size_t)i * sg.ld32] : 0.f;   // all loads in flight
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += a[i];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += bias_v;
          if (flags & EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          } else if (flags & EPI_SILU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = silu(v[i]);
          }
          if (flags & EPI_MASK_BLOCKDIAG) {
            const int ch = (cbase + lane) / p.head_dim;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if ((row0 + i) / p.head_dim != ch) v[i] = 0.f;
          }
          if (sg.out32 != nullptr && col_ok) {
            float* op32 = sg.out32 + rowg0 * (size_t)sg.ld32 + colg;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nrows) op32[(size_t)i * sg.ld32] = v[i];
          }
          if (has_op && (col_ok || col_pad)) {
            const size_t ob = rowg0 * (size_t)sg.op.ld + colg;
            if (op_fmt == OP_F16) {
              uint16_t* oh = reinterpret_cast<uint16_t*>(sg.op.hi) + ob;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nrows) oh[(size_t)i * sg.op.ld] = col_ok ? f32_to_f16_bits(v[i]) : (uint16_t)0;
            } else {
              uint16_t* oh = reinterpret_cast<uint16_t*>(sg.op.hi) + ob;
              uint16_t* ol = reinterpret_cast<uint16_t*>(sg.op.lo) + ob;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (i < nrows) {
                  uint16_t h16 = 0, l16 = 0;
                  if (col_ok) f32_to_bf16x2_bits(v[i], h16, l16);
                  oh[(size_t)i * sg.op.ld] = h16;
                  ol[(size_t)i * sg.op.ld] = l16;
                }
              }
            }
          }
        } else {
          // ---- transposed destination: element (r, c) -> [(outer*trans_rows + c) * ld + r]; lanes
          // (consecutive r) are already contiguous in memory, so every access is a full 128-byte line.
          const bool row_valid = r < p.M;
          const bool row_pad = !row_valid && r < p.M_pad;
          if (sg.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) v[j] += __ldg(bias_p + cbase + j);
          }
          const size_t trow0 = (size_t)outer * p.trans_rows + (size_t)colg0;
          if (row_valid) {
            if (sg.addend != nullptr) {
              const float* ap = sg.addend + trow0 * (size_t)sg.ld32 + r;
              float a[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) a[j] = (j < nvalid) ? ap[(size_t)j * sg.ld32] : 0.f;   // loads first:
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += a[j];   // the destination may alias the addend (residual add)
            }
            if (flags & EPI_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            } else if (flags & EPI_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
            }
            if (flags & EPI_MASK_BLOCKDIAG) {
              const int rh = r / p.head_dim;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if ((cbase + j) / p.head_dim != rh) v[j] = 0.f;
            }
            if (sg.out32 != nullptr) {
              float* op32 = sg.out32 + trow0 * (size_t)sg.ld32 + r;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) op32[(size_t)j * sg.ld32] = v[j];
            }
          }
          if (has_op && (row_valid || row_pad)) {
            const size_t ob = trow0 * (size_t)sg.op.ld + r;
            if (op_fmt == OP_F16) {
              uint16_t* oh = reinterpret_cast<uint16_t*>(sg.op.hi) + ob;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nvalid) oh[(size_t)j * sg.op.ld] = row_valid ? f32_to_f16_bits(v[j]) : (uint16_t)0;
            } else {
              uint16_t* oh = reinterpret_cast<uint16_t*>(sg.op.hi) + ob;
              uint16_t* ol = reinterpret_cast<uint16_t*>(sg.op.lo) + ob;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < nvalid) {
                  uint16_t h16 = 0, l16 = 0;
                  if (row_valid) f32_to_bf16x2_bits(v[j], h16, l16);
                  oh[(size_t)j * sg.op.ld] = h16;
                  ol[(size_t)j * sg.op.ld] = l16;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_remote(smem_u32(&tempty_bar[acc]), 0u);   // the leader's MMA thread waits for both CTAs
        else mbar_arrive(smem_u32(&tempty_bar[acc]));
      }
    }
    if (ew == 0) MCM_STAMP(5);
    if (lane == 0) tma_wait_all0();   // bulk stores must have fully completed before the CTA exits
    __syncwarp();
    if (ew == 0) MCM_STAMP(6);
    if (prof && lane == 0 && p.dbg != nullptr) {
      for (int i = 0; i < 6; ++i) atomicAdd(p.dbg + i, (unsigned long long)tph[i]);
    }
#undef MCM_TICK
  }

  tc_fence_before();
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();     // no CTA may exit while peers still multicast into / arrive on its smem
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
    MCM_STAMP(7);
  }
#undef MCM_STAMP
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
int g_sm_share = 1;        // persistent kernels take 1 / g_sm_share of the SMs (two batch halves side by side: tc_set_sm_share)
std::once_flag g_init_once;
int g_init_status = 0;
std::atomic<unsigned long long> g_launches{0};
int g_debug_epi = 0;
unsigned long long* g_dbg = nullptr;
// MCM_GEMM_TRACE=<file>: per-launch time stamps of CTA 0 (development aid for the launch-bound small-batch regime); the file is
// written at process exit, one line per launch slot (a CUDA-graph replay overwrites the slots of its nodes).
constexpr int TRACE_SLOTS = 4096, TRACE_WORDS = 64;
unsigned long long* g_trace = nullptr;
std::atomic<unsigned long long> g_trace_n{0};
std::vector<std::string>* g_trace_meta = nullptr;
const char* g_trace_path = nullptr;
void trace_dump() {
  if (g_trace == nullptr || g_trace_path == nullptr) return;
  if (cudaDeviceSynchronize() != cudaSuccess) return;
  std::vector<unsigned long long> h((size_t)TRACE_SLOTS * TRACE_WORDS);
  if (cudaMemcpy(h.data(), g_trace, h.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  FILE* f = fopen(g_trace_path, "w");
  if (!f) return;
  const size_t n = std::min<size_t>(g_trace_meta->size(), TRACE_SLOTS);
  for (size_t i = 0; i < n; ++i) {
    fprintf(f, "%zu %s |", i, (*g_trace_meta)[i].c_str());
    for (int j = 0; j < TRACE_WORDS; ++j) fprintf(f, " %llu", h[i * TRACE_WORDS + j]);
    fprintf(f, "\n");
  }
  fclose(f);
}
int g_pair = 0;                     // MCM_PAIR=1 enables cta_group::2 pair tiles (256 x block_n per CTA pair; validated,
                                    // -5..10 % mainloop time but a slower epilogue overlap: net neutral this round)
int g_prefetch = 0;                 // MCM_PREFETCH=1: producer issues L2 prefetches one work item ahead (measured: no gain)
int g_stage_cap = MAX_STAGES;      // MCM_STAGES: cap on the operand ring depth (experiments)
int g_max_cs = 1;                 // MCM_MAX_CLUSTER: largest cluster size the launcher may choose (1, 2 or 4).
                                  // Multicast of the B tile is implemented and tested, but measured neutral on B200 for
                                  // clusters <= 4 (the mainloop is bound by L2->SM ingest per SM, which multicast does not cut)
int g_max_pairs = 0;              // co-resident CTA pairs of the cta_group::2 instantiation
int g_max_clusters[5] = {0, 0, 0, 0, 0};   // co-resident clusters per cluster size (cudaOccupancyMaxActiveClusters)
bool g_force_generic = false;   // MCM_GENERIC_EPILOGUE=1: disable the TMA epilogue (A/B testing, debugging)

int do_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  MCM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  MCM_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  MCM_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MCM_CUDA(cudaGetDeviceProperties(&prop, dev));
  MCM_CHECK(prop.major == 10, "motioncraft_b200 needs an sm_100a (B200) device; there is no fallback path");
  g_num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("MCM_GENERIC_EPILOGUE")) g_force_generic = (e[0] == '1');
  if (const char* e = getenv("MCM_DEBUG_EPI")) g_debug_epi = atoi(e);
  if (const char* e = getenv("MCM_PREFETCH")) g_prefetch = atoi(e);
  if (const char* e = getenv("MCM_PAIR")) g_pair = atoi(e);
  if (const char* e = getenv("MCM_STAGES")) g_stage_cap = std::max(1, std::min(MAX_STAGES, atoi(e)));
  if (const char* e = getenv("MCM_MAX_CLUSTER")) g_max_cs = std::max(1, std::min(4, atoi(e)));
  MCM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  MCM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  MCM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  MCM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  for (int cs = 1; cs <= 4; cs *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g_num_sms / cs * cs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<false, false>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    g_max_clusters[cs] = n;
    if (cs == 2) {
      int np = 0;
      if (cudaOccupancyMaxActiveClusters(&np, gemm_tc_kernel<false, true>, &cfg) != cudaSuccess) { cudaGetLastError(); np = 0; }
      g_max_pairs = np;
    }
  }
  MCM_CHECK(g_max_clusters[1] > 0, "gemm_tc_kernel does not fit on this device");
  if (g_debug_epi == 3) {
    MCM_CUDA(cudaMalloc(&g_dbg, 16 * sizeof(unsigned long long)));
    MCM_CUDA(cudaMemset(g_dbg, 0, 16 * sizeof(unsigned long long)));
  }
  return 0;
}

int make_map(CUtensorMap* m, const void* ptr, int fmt, int k_dim, int rows, int batches, int ld, int box_rows) {
  MCM_CHECK(ptr != nullptr, "null operand pointer");
  MCM_CHECK((ld % 8) == 0, "operand pitch must be a multiple of 8 elements (TMA 16-byte rule)");
  MCM_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "operand base must be 16-byte aligned");
  cuuint64_t gdim[3] = {(cuuint64_t)k_dim, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, fmt == OP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                        const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (k=" + std::to_string(k_dim) +
              " rows=" + std::to_string(rows) + " batches=" + std::to_string(batches) + " ld=" + std::to_string(ld) +
              " box_rows=" + std::to_string(box_rows) + ")");
    return 1;
  }
  return 0;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// 3-D tensor map of an epilogue destination / addend: dims {d0, d1, d2} (d0 contiguous), 32 x 32 x 1 boxes
int make_epi_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int esize, long long d0, long long d1,
                 long long d2, long long stride1_elems, long long stride2_elems, CUtensorMapSwizzle sw) {
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)stride1_elems * esize, (cuuint64_t)stride2_elems * esize};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, dt, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (epilogue map) failed with CUresult " + std::to_string((int)r));
    return 1;
  }
  return 0;
}
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Decide whether segment `sg` can use the TMA epilogue and build its maps.  Falls back (mode 0) whenever an
// alignment rule of bulk tensor copies does not hold (e.g. the [B*T, 322] eps output: 1288-byte rows).
int setup_epi_maps(const GemmProblem& q, const EpiSeg& sg, int M_pad, int* mode, CUtensorMap* maps) {
  *mode = 0;
  const bool transposed = (sg.flags & EPI_TRANSPOSED) != 0;
  const bool bcast = (sg.flags & EPI_ADDEND_BCAST) != 0;
  const long long outer = q.out_batched ? q.batches : q.batches / q.inner;
  if (sg.out32 == nullptr && sg.op.hi == nullptr) return 0;
  int md = TM_TMA;
  bool ok = true;
  if (sg.out32) ok = ok && al16(sg.out32) && (sg.ld32 % 4) == 0;
  if (sg.op.hi) ok = ok && al16(sg.op.hi) && (sg.op.ld % 8) == 0 && (sg.op_fmt == OP_F16 || al16(sg.op.lo));
  if (sg.addend) {
    if (sg.addend == sg.out32 && sg.op.hi == nullptr && !bcast) md |= TM_RED;
    else if (!transposed) { md |= TM_LDADD | (bcast ? TM_BCAST : 0); ok = ok && al16(sg.addend) && (sg.ld32 % 4) == 0; }
    else ok = false;
  }
  if (!ok) return 0;
  const long long col_end = (long long)sg.col0 + (long long)(q.inner - 1) * q.out_col_inner;
  const CUtensorMapDataType opdt = sg.op_fmt == OP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if (!transposed) {
    const long long rpo = q.out_rows_per_outer;
    if (sg.out32) {
      if (col_end + sg.n > sg.ld32) return 0;
      MCM_TRY(make_epi_map(&maps[0], sg.out32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, col_end + sg.n, q.M, outer, sg.ld32,
                           rpo * sg.ld32, CU_TENSOR_MAP_SWIZZLE_128B));
    }
    if (md & TM_LDADD) {
      MCM_TRY(make_epi_map(&maps[1], sg.addend, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, col_end + sg.n, q.M, bcast ? 1 : outer,
                           sg.ld32, (bcast ? (long long)q.M : rpo) * sg.ld32, CU_TENSOR_MAP_SWIZZLE_128B));
    }
    if (sg.op.hi) {
      const long long ext = std::min<long long>(col_end + sg.n_pad, sg.op.ld);
      MCM_TRY(make_epi_map(&maps[2], sg.op.hi, opdt, 2, ext, q.M, outer, sg.op.ld, rpo * sg.op.ld, CU_TENSOR_MAP_SWIZZLE_64B));
      if (sg.op_fmt != OP_F16)
        MCM_TRY(make_epi_map(&maps[3], sg.op.lo, opdt, 2, ext, q.M, outer, sg.op.ld, rpo * sg.op.ld, CU_TENSOR_MAP_SWIZZLE_64B));
    }
  } else {
    // destination [(outer * trans_rows + c) * ld + r]: dims {r, c, outer}
    if (sg.out32) {
      if (q.M > sg.ld32) return 0;
      MCM_TRY(make_epi_map(&maps[0], sg.out32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, q.M, col_end + sg.n, outer, sg.ld32,
                           (long long)q.trans_rows * sg.ld32, CU_TENSOR_MAP_SWIZZLE_NONE));
    }
    if (sg.op.hi) {
      const long long rext = std::min<long long>(M_pad, sg.op.ld);
      MCM_TRY(make_epi_map(&maps[2], sg.op.hi, opdt, 2, rext, col_end + sg.n, outer, sg.op.ld,
                           (long long)q.trans_rows * sg.op.ld, CU_TENSOR_MAP_SWIZZLE_NONE));
      if (sg.op_fmt != OP_F16)
        MCM_TRY(make_epi_map(&maps[3], sg.op.lo, opdt, 2, rext, col_end + sg.n, outer, sg.op.ld,
                             (long long)q.trans_rows * sg.op.ld, CU_TENSOR_MAP_SWIZZLE_NONE));
    }
  }
  *mode = md;
  return 0;
}
}  // namespace

std::string g_init_error;

int gemm_tc_init() {
  std::call_once(g_init_once, [] {
    g_init_status = do_init();
    if (g_init_status != 0) g_init_error = mcm_last_error_string();
  });
  if (g_init_status != 0) set_error("gemm_tc_init failed: " + g_init_error);
  return g_init_status;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MCM_PDL"); return e == nullptr || e[0] != '0'; }();
  return on;
}

unsigned long long gemm_tc_launch_count() { return g_launches.load(); }
void gemm_tc_count_replayed(unsigned long long n) { g_launches.fetch_add(n); }

int gemm_tc_debug_read(unsigned long long* out, int reset) {
  for (int i = 0; i < 16; ++i) out[i] = 0;
  if (g_dbg == nullptr) return 0;
  MCM_CUDA(cudaDeviceSynchronize());
  MCM_CUDA(cudaMemcpy(out, g_dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) MCM_CUDA(cudaMemset(g_dbg, 0, 16 * sizeof(unsigned long long)));
  return 0;
}

// exported map builders (the fused block kernel builds its own tensor maps with the same rules)
int tc_make_operand_map(CUtensorMap* m, const void* ptr, int fmt, int k_dim, int rows, int batches, int ld, int box_rows) {
  MCM_TRY(gemm_tc_init());
  return make_map(m, ptr, fmt, k_dim, rows, batches, ld, box_rows);
}
int tc_make_tile_map(CUtensorMap* m, const void* ptr, int dtype, long long d0, long long d1, long long d2,
                     long long stride1_elems, long long stride2_elems, int swizzle_bytes) {
  MCM_TRY(gemm_tc_init());
  const CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  return make_epi_map(m, ptr, dt, dtype == 0 ? 4 : 2, d0, d1, d2, stride1_elems, stride2_elems, sw);
}
int tc_make_box_map(CUtensorMap* m, const void* ptr, int dtype, long long d0, long long d1, long long d2,
                    long long stride1_elems, long long stride2_elems, int box0, int box1) {
  MCM_TRY(gemm_tc_init());
  const int esize = dtype == 0 ? 4 : 2;
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)stride1_elems * esize, (cuuint64_t)stride2_elems * esize};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                        const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (box map) failed with CUresult " + std::to_string((int)r));
    return 1;
  }
  return 0;
}
int tc_num_sms() { return gemm_tc_init() == 0 ? g_num_sms : 0; }
void tc_set_sm_share(int share) { g_sm_share = share >= 1 ? share : 1; }
int tc_sm_share() { return g_sm_share; }

int gemm_tc_launch(const GemmProblem& q, cudaStream_t stream) {
  MCM_TRY(gemm_tc_init());
  MCM_CHECK(q.nseg >= 1 && q.nseg <= 3, "1..3 output segments");
  MCM_CHECK(q.M > 0 && q.K > 0 && q.batches > 0 && q.inner > 0 && q.batches % q.inner == 0, "bad GEMM iteration space");
  MCM_CHECK(q.a.hi && q.b.hi, "missing operand");
  const int split = q.fmt == OP_BF16X2 ? 1 : 0;
  if (split) MCM_CHECK(q.a.lo && q.b.lo, "bf16x2 operands need their lo halves");

  KParams p{};
  p.M = q.M;
  p.M_pad = round_up(q.M, 8);
  p.K = q.K;
  p.batches = q.batches;
  p.inner = q.inner;
  p.a_k_inner = q.a_k_inner;
  p.b_k_inner = q.b_k_inner;
  p.a_batched = q.a_batched;
  p.out_batched = q.out_batched;
  p.bias_inner = q.bias_inner;
  MCM_CHECK(q.a_k_inner % 8 == 0 && q.b_k_inner % 8 == 0, "per-inner K offsets must be multiples of 8 elements (TMA start alignment)");
  p.b_batched = q.b_batched;
  p.out_col_inner = q.out_col_inner;
  p.out_rows_per_outer = q.out_rows_per_outer;
  p.trans_rows = q.trans_rows;
  p.head_dim = q.head_dim > 0 ? q.head_dim : 1;
  p.split = split;
  p.nseg = q.nseg;
  p.debug = g_debug_epi;
  p.prefetch = g_prefetch;
  p.dbg = g_dbg;
  p.trace = nullptr;
  static const bool trace_on = [] {
    g_trace_path = getenv("MCM_GEMM_TRACE");
    if (g_trace_path == nullptr) return false;
    if (cudaMalloc(&g_trace, (size_t)TRACE_SLOTS * TRACE_WORDS * 8) != cudaSuccess) return false;
    cudaMemset(g_trace, 0, (size_t)TRACE_SLOTS * TRACE_WORDS * 8);
    g_trace_meta = new std::vector<std::string>();
    atexit(trace_dump);
    return true;
  }();

  int nmax = 0;
  for (int s = 0; s < q.nseg; ++s) {
    p.seg[s] = q.seg[s];
    EpiSeg& sg = p.seg[s];
    MCM_CHECK(sg.n > 0, "empty segment");
    const bool transposed = (sg.flags & EPI_TRANSPOSED) != 0;
    sg.n_pad = (sg.op.hi != nullptr && !transposed) ? round_up(sg.n, 8) : sg.n;
    if (sg.op.hi != nullptr) {
      MCM_CHECK(sg.op_fmt == OP_F16 || sg.op.lo != nullptr, "bf16x2 operand output needs a lo buffer");
      if (!transposed) MCM_CHECK(sg.op.ld >= sg.col0 + (q.inner - 1) * q.out_col_inner + sg.n_pad, "operand output pitch too small");
    }
    sg.vec32 = ((sg.ld32 % 4) == 0 && ((sg.col0 + 0) % 4) == 0 && (q.out_col_inner % 4) == 0 &&
                (sg.out32 == nullptr || (reinterpret_cast<uintptr_t>(sg.out32) & 15) == 0) &&
                (sg.addend == nullptr || (reinterpret_cast<uintptr_t>(sg.addend) & 15) == 0))
                   ? 4 : 1;
    nmax = std::max(nmax, sg.n_pad);
  }
  p.m_tiles = (q.M + BLOCK_M - 1) / BLOCK_M;
  // cluster size: CTAs of a cluster take consecutive M tiles of ONE (batch, N tile) and share its B tile by
  // TMA multicast, cutting the L2 -> SM operand traffic that bounds these small-K GEMMs
  int cs = 1;
  for (int c = g_max_cs; c >= 2; c /= 2) {
    const bool aligned = (q.batches == 1) ? (p.m_tiles >= c) : (p.m_tiles % c == 0);
    if (aligned && g_max_clusters[c] > 0) { cs = c; break; }
  }
  // CTA pairs (cta_group::2): the two CTAs of a cluster compute one 256 x block_n tile, each holding its own 128 A rows
  // and HALF of the B rows -- a third fewer operand bytes per MMA through the L2 -> SM path that bounds the mainloop.
  p.pair = 0;
  if (g_pair && g_max_pairs > 0 && ((q.batches == 1) ? (p.m_tiles >= 2) : (p.m_tiles % 2 == 0))) {
    p.pair = 1;
    cs = 2;
  }
  p.cs = cs;
  p.m_supers = (p.m_tiles + cs - 1) / cs;
  int bn_cap = split ? 128 : 256;
  {
    // Latency regime (small batches: a GEMM of a few tiles leaves most SMs idle and its duration is the epilogue of ONE
    // CTA -- 16 us per launch at B = 1): narrower N tiles spread the same work over more CTAs, each with a shorter epilogue.
    static const int spread = [] { const char* e = getenv("MCM_GEMM_SPREAD"); return e ? atoi(e) : 1; }();
    static const int min_bn = [] { const char* e = getenv("MCM_GEMM_MIN_BN"); return e ? atoi(e) : 64; }();
    auto items = [&](int cap) {
      long long t = 0;
      for (int s2 = 0; s2 < q.nseg; ++s2) t += (p.seg[s2].n_pad + cap - 1) / cap;
      return t * ((p.m_tiles + cs - 1) / cs) * q.batches;
    };
    while (spread && bn_cap > min_bn && items(bn_cap) * 2 <= g_num_sms) bn_cap /= 2;
  }
  const int tiles_for_max = (nmax + bn_cap - 1) / bn_cap;
  // a segment split over several N tiles needs tiles of whole 32-column chunks: the TMA epilogue stores
  // 32-wide boxes and must not spill into the neighbouring tile's columns
  // ... and a tile is fetched as cs slices of whole 8-row swizzle atoms
  const int gran = std::max(tiles_for_max > 1 ? 32 : 16, 8 * cs);
  p.block_n = std::min(bn_cap, round_up((nmax + tiles_for_max - 1) / tiles_for_max, gran));
  int tile0 = 0;
  for (int s = 0; s < q.nseg; ++s) {
    p.seg[s].tile0 = tile0;
    p.seg[s].n_tiles = (p.seg[s].n_pad + p.block_n - 1) / p.block_n;
    tile0 += p.seg[s].n_tiles;
  }
  p.n_tiles = tile0;
  EpiMaps em;
  std::memset(&em, 0, sizeof(em));
  for (int s = 0; s < q.nseg; ++s) {
    MCM_TRY(setup_epi_maps(q, p.seg[s], p.M_pad, &p.tma_mode[s], em.m[s]));
    if (g_force_generic) p.tma_mode[s] = 0;
  }
  p.num_kb = (q.K + BLOCK_K - 1) / BLOCK_K;
  p.total_tiles = p.n_tiles * p.m_supers * q.batches;   // cluster work items
  p.a_bytes = BLOCK_M * BLOCK_K * 2;
  p.b_bytes = (uint32_t)(p.pair ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;   // per CTA
  p.stage_bytes = (split ? 2u : 1u) * (p.a_bytes + p.b_bytes);
  bool big_stg = false;
  for (int s = 0; s < q.nseg; ++s)
    big_stg = big_stg || !(p.tma_mode[s] & TM_TMA) || (p.seg[s].out32 != nullptr && p.seg[s].op.hi != nullptr);
  p.stg_bytes = big_stg ? 8192 : 4096;
  const int ring_budget = SMEM_TOTAL - 1024 - NUM_EPI_WARPS * p.stg_bytes;
  p.stages = std::min(g_stage_cap, (int)(ring_budget / p.stage_bytes));
  MCM_CHECK(p.stages >= 1, "tile does not fit in shared memory");
  // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32, a/b format, K-major both, N>>3, M>>4
  const uint32_t ab = split ? 1u : 0u;   // 0 = F16, 1 = BF16
  p.idesc = (1u << 4) | (ab << 7) | (ab << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
            ((uint32_t)((p.pair ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);

  CUtensorMap tmA, tmAlo, tmB, tmBlo;
  MCM_TRY(make_map(&tmA, q.a.hi, q.fmt, q.a_k, q.a_rows, q.a_batches, q.a.ld, BLOCK_M));
  MCM_TRY(make_map(&tmB, q.b.hi, q.fmt, q.b_k, q.b_rows, q.b_batches, q.b.ld, p.block_n / cs));
  if (split) {
    MCM_TRY(make_map(&tmAlo, q.a.lo, q.fmt, q.a_k, q.a_rows, q.a_batches, q.a.ld, BLOCK_M));
    MCM_TRY(make_map(&tmBlo, q.b.lo, q.fmt, q.b_k, q.b_rows, q.b_batches, q.b.ld, p.block_n / cs));
  } else {
    tmAlo = tmA;
    tmBlo = tmB;
  }
  const int n_clusters = std::min(p.total_tiles, std::max(1, (p.pair ? g_max_pairs : g_max_clusters[cs]) / g_sm_share));
  const int grid = n_clusters * cs;
  if (trace_on) {
    const unsigned long long ti = g_trace_n.fetch_add(1);
    if (ti < (unsigned long long)TRACE_SLOTS) {
      p.trace = g_trace + ti * TRACE_WORDS;
      char buf[160];
      snprintf(buf, sizeof(buf), "M=%d K=%d batches=%d nseg=%d n0=%d bn=%d tiles=%d grid=%d cs=%d stages=%d", q.M, q.K, q.batches,
               q.nseg, q.seg[0].n, p.block_n, p.total_tiles, grid, p.cs, p.stages);
      g_trace_meta->push_back(buf);
    }
  }
  double flops = q.algo_flops;
  if (flops <= 0.0) {
    long long ncols = 0;
    for (int s = 0; s < q.nseg; ++s) ncols += q.seg[s].n;
    flops = 2.0 * (double)q.M * (double)ncols * (double)q.K * (double)q.batches;
  }
  {
    LaunchTimer lt(LK_GEMM, stream, flops);
    bool need_generic = false;
    for (int s = 0; s < q.nseg; ++s) need_generic = need_generic || !(p.tma_mode[s] & TM_TMA);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t le;
    if (p.pair) le = need_generic ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, true>, tmA, tmAlo, tmB, tmBlo, em, p)
                                  : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, true>, tmA, tmAlo, tmB, tmBlo, em, p);
    else le = need_generic ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, false>, tmA, tmAlo, tmB, tmBlo, em, p)
                           : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<false, false>, tmA, tmAlo, tmB, tmBlo, em, p);
    MCM_CUDA(le);
  }
  MCM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  if (g_debug_epi == 3 && getenv("MCM_DEBUG_PRINT")) {
    unsigned long long d[16];
    gemm_tc_debug_read(d, 1);
    const double ch = d[5] ? (double)d[5] : 1.0;
    fprintf(stderr, "gemm M=%d K=%d batches=%d bn=%d tiles=%d nseg=%d n0=%d flags0=%d tma=%d/%d/%d pair=%d | per chunk: acc %.0f stg %.0f ld %.0f math %.0f issue %.0f | chunks %llu"
            " || producer: wait-empty %.0f%% of %.0f cyc/kblock | mma: wait-acc %.0f%% wait-full %.0f%% of its time, %.0f cyc/kblock\n",
            q.M, q.K, q.batches, p.block_n, p.total_tiles, q.nseg, q.seg[0].n, q.seg[0].flags, p.tma_mode[0], p.tma_mode[1],
            p.tma_mode[2], p.pair, d[0] / ch, d[1] / ch, d[2] / ch, d[3] / ch, d[4] / ch, d[5],
            100.0 * d[8] / (d[9] ? (double)d[9] : 1.0), (double)d[9] / (d[10] ? (double)d[10] : 1.0),
            100.0 * d[11] / (d[13] ? (double)d[13] : 1.0), 100.0 * d[12] / (d[13] ? (double)d[13] : 1.0),
            (double)d[13] / (d[10] ? (double)d[10] : 1.0));
  }
  return 0;
}

}  // namespace mcm
