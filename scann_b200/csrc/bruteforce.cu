// bruteforce.cu -- bf16 brute-force MIPS scoring as a tcgen05 / TMEM tensor-core GEMM fed by TMA,
// with the top-k pre-filter fused into the TMEM epilogue.
//
// Replaces (a12) of SURVEY.md section 8:
//   Bfloat16BruteForceSearcher::FindNeighborsImpl   brute_force/bfloat16_brute_force.cc:101-152
//   OneToManyBf16FloatImpl                          distance_measures/one_to_many/one_to_many_asymmetric_impl.inc
//   FastTopNeighbors::PushBlock                      utils/fast_top_neighbors.h:394-440
// The reference streams the whole bf16 database once PER QUERY (f32 query x bf16 row, f32
// accumulate).  Here a batch of queries is one GEMM  S[q][i] = sum_d q[d] * x[i][d]:
//   * the f32 query is split into two bf16 terms (hi + lo, 16 mantissa bits) stacked as two
//     A operands; both MMAs accumulate into the same fp32 TMEM tile, so the approximate score
//     is within ~2^-16 relative of the f32-query score;
//   * 128 queries x 256 database rows per CTA, K walked in 64-element (128-byte, SWIZZLE_128B)
//     slices through a 3-stage TMA -> mbarrier -> tcgen05.mma pipeline (one elected thread
//     issues the MMAs, tcgen05.commit releases the stage);
//   * the epilogue warps read the accumulator with tcgen05.ld (one TMEM lane = one query) and
//     push only scores that beat the query's current threshold key into the same per-query
//     candidate buffers the LUT16 scan uses; the database is walked in geometrically growing
//     rounds with a compaction between rounds, so the inflow stays ~4k' per round;
//   * the final k' candidates are re-scored exactly (f32 query x bf16 row, f32 FMA chain, the
//     oracle's arithmetic) and the top-k of those is returned, so ids and distances are the
//     reference's as long as the true top-k is inside the over-retrieved k' = 2k + 64 set.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"

namespace sb {

namespace bf {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kSplits = 2;
constexpr int kABytes = BM * BK * 2;                       // 16 KB per split
constexpr int kBBytes = BN * BK * 2;                       // 32 KB
constexpr int kThreads = 384;                              // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-11 epilogue
constexpr int kEpiWarps = 8;                               // two per TMEM lane quadrant, 128 columns each
constexpr int kTmemCols = 512;                             // two 256-column fp32 accumulators

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmArgs {
  uint32_t nq, m_pad;            // valid A rows, padded A rows per split
  uint32_t mt, nt;               // tiles along M (A rows) and N (B rows of this launch)
  uint32_t row0, row1;           // B rows [row0, row1) of this launch
  uint32_t num_kb;               // ceil(K / 64)
  // kEpiFilter: candidate buffers
  uint64_t* buf; uint32_t* cnt; const uint64_t* tau; uint32_t* ovf; uint32_t cap;
  // kEpiStore: raw accumulators, out[a_row * ld + b_row]; optionally the maximum of every 32-column chunk,
  // cmax[a_row * ld_c + b_row / 32] (the chunk pre-selection of the tokenizer, prep.cu)
  float* out; uint32_t ld;
  float* cmax; uint32_t ld_c;
  const float* cbias;  // with cmax: per-column bias b; the chunk statistic becomes -min_j(b[j] - 2 acc[j]) (squared L2)
  int prof;   // debug: CTA 0 prints where its MMA thread waited (SCANN_B200_GEMM_PROFILE=1)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Threshold filter of one accumulator tile for one query (= one TMEM lane), 256 columns.
// Pass 1 builds a per-chunk bit mask with ONE float compare per score against thr = -distance(tau)
// (a superset of "key < tau": score ties with tau are included); one reservation per (query, tile);
// pass 2 re-reads only chunks with survivors, applies the exact 64-bit key test and pads the few
// reserved-but-rejected slots with kKeyMax, which every later compaction sorts past the end.
__device__ __forceinline__ void filter_tile(const GemmArgs& a, uint32_t tbase, uint32_t q, bool qvalid, uint32_t n0,
                                            uint32_t* masks /* [BN/32] stride 128 */, int c_begin, int c_end) {
  const uint64_t tau = qvalid ? a.tau[q] : 0ull;
  float thr = __int_as_float(0x7F800000);                         // +inf: nothing passes
  if (qvalid) thr = tau == kKeyMax ? __int_as_float(0xFF800000)   // -inf: everything passes
                                   : -ord2f((uint32_t)(tau >> 32));
  uint32_t total = 0;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t v[32];
    tmem_ld32(tbase + (uint32_t)(c * 32), v);
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (__uint_as_float(v[j]) >= thr) m0 |= 1u << j;
      if (__uint_as_float(v[j + 1]) >= thr) m1 |= 2u << j;
      if (__uint_as_float(v[j + 2]) >= thr) m2 |= 4u << j;
      if (__uint_as_float(v[j + 3]) >= thr) m3 |= 8u << j;
    }
    uint32_t m = (m0 | m1) | (m2 | m3);
    const uint32_t col = n0 + c * 32;  // columns past the end of this round's rows do not exist
    if (col + 32 > a.row1) m &= col < a.row1 ? (0xFFFFFFFFu >> (32 - (a.row1 - col))) : 0u;
    if (!qvalid) m = 0u;  // padding rows of the A operand: an inf/NaN accumulator must not reserve slots
    masks[c * 128] = m;
    total += __popc(m);
  }
  uint32_t pos = total ? atomicAdd(&a.cnt[q], total) : 0u;
  if (__any_sync(0xFFFFFFFFu, total != 0)) {
    const uint32_t end = pos + total;
    uint64_t* dst = a.buf + (size_t)q * a.cap;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      const uint32_t m = masks[c * 128];
      if (__any_sync(0xFFFFFFFFu, m != 0)) {
        uint32_t v[32];
        tmem_ld32(tbase + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if ((m >> j) & 1u) {
            const uint64_t key = ((uint64_t)f2ord(-__uint_as_float(v[j])) << 32) | (n0 + c * 32 + j);
            if (key < tau) {
              if (pos < a.cap) dst[pos] = key;
              ++pos;
            }
          }
        }
      }
    }
    for (; pos < end; ++pos)
      if (pos < a.cap) dst[pos] = kKeyMax;
    if (end > a.cap) a.ovf[q] = 1u;
  }
}

constexpr int kEpiFilter = 0, kEpiStore = 1;

// The store epilogue transposes its 32 x 32 blocks through shared memory (below): one stage less, 8 x 4.5 KB of blocks.
constexpr int kTrRowBytes = 144;                 // 32 floats + 16 bytes of padding: 16-byte aligned rows, conflict-free
constexpr int kTrBlockBytes = 32 * kTrRowBytes;  // per epilogue warp
template <int kSplitsT, int kEpi = 0>
struct StageLayout {
  static constexpr int kBytes = kSplitsT * kABytes + kBBytes;
  static constexpr int kStagesT = kSplitsT == 1 ? (kEpi == 1 ? 3 : 4) : 3;
  static constexpr int kTrBytes = kEpi == 1 ? kEpiWarps * kTrBlockBytes : 0;
  static constexpr size_t kSmem = (size_t)kStagesT * kBytes + kTrBytes + 1024;
};

// Persistent GEMM: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... with the A
// (query) tile index fastest, so the CTAs resident at any moment share a couple of B tiles through
// L2 while the whole A operand stays L2-resident.  The fp32 accumulator is double-buffered in TMEM
// (2 x 256 columns): the epilogue warps drain tile i while the MMA warp already accumulates tile
// i + 1.
template <int kSplitsT, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs a) {
  using SL = StageLayout<kSplitsT, kEpi>;
  constexpr int kSt = SL::kStagesT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kSt], empty_bar[kSt], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_masks[kEpi == kEpiFilter ? BN / 32 : 1][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t total_tiles = a.mt * a.nt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer ----
      uint32_t it = 0;  // running k-block counter across tiles
      for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const uint32_t m0 = (t % a.mt) * BM, n0 = a.row0 + (t / a.mt) * BN;
        for (uint32_t kb = 0; kb < a.num_kb; ++kb, ++it) {
          const int st = it % kSt;
          const uint32_t ph = (it / kSt) & 1;
          mbar_wait(&empty_bar[st], ph ^ 1);
          uint8_t* base = smem + (size_t)st * SL::kBytes;
          mbar_expect_tx(&full_bar[st], SL::kBytes);
          for (int s = 0; s < kSplitsT; ++s)
            tma_load_2d(base + s * kABytes, &tmA, &full_bar[st], (int)(kb * BK), (int)(s * a.m_pad + m0));
          tma_load_2d(base + kSplitsT * kABytes, &tmB, &full_bar[st], (int)(kb * BK), (int)n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---- MMA issuer ----
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      uint32_t it = 0, lt = 0;
      for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aph ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem + as * BN;
        for (uint32_t kb = 0; kb < a.num_kb; ++kb, ++it) {
          const int st = it % kSt;
          const uint32_t ph = (it / kSt) & 1;
          mbar_wait(&full_bar[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sbase = smem_u32(smem + (size_t)st * SL::kBytes);
          const uint32_t sB = sbase + kSplitsT * kABytes;
#pragma unroll
          for (int s = 0; s < kSplitsT; ++s) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(sbase + s * kABytes + k * 32);
              const uint64_t bd = umma_desc_sw128(sB + k * 32);
              umma_bf16(tacc, ad, bd, idesc, (kb | (uint32_t)s | (uint32_t)k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[st]);  // frees the smem stage once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[as]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {  // ---- epilogue: one TMEM lane = one A row (query) ----
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;  // which 128 columns of the tile
    uint32_t lt = 0;
    for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      const uint32_t m0 = (t % a.mt) * BM, n0 = a.row0 + (t / a.mt) * BN;
      const uint32_t q = m0 + quad * 32 + lane;
      const bool qvalid = q < a.nq;
      const uint32_t tbase = tmem + as * BN + ((uint32_t)(quad * 32) << 16);
      if (kEpi == kEpiFilter) {
        mbar_wait(&tmem_full_bar[as], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        filter_tile(a, tbase, q, qvalid, n0, &s_masks[0][quad * 32 + lane], half * (BN / 64), (half + 1) * (BN / 64));
      } else {
        mbar_wait(&tmem_full_bar[as], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* dst = a.out + (size_t)q * a.ld;
#pragma unroll
        for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
          uint32_t v[32];
          tmem_ld32(tbase + (uint32_t)(c * 32), v);
          const uint32_t col = n0 + c * 32;
          if (qvalid && a.cmax && col < a.row1) {
            float mx = __int_as_float(0xFF800000);
            if (!a.cbias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < a.row1) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else if (col + 32 <= a.row1) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a.cbias + col + j));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j])), b.x));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j + 1])), b.y));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j + 2])), b.z));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j + 3])), b.w));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < a.row1) mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j])), __ldg(a.cbias + col + j)));
            }
            a.cmax[(size_t)q * a.ld_c + (col >> 5)] = mx;
          }
          if (a.out && col + 32 <= a.row1 && (a.ld & 3u) == 0) {
            // A TMEM lane is an output ROW: stored straight from the registers, every instruction would write 16 bytes
            // into each of 32 rows (32 half-used sectors, the LSU queue was the limiter of the 40k-centre tokenization).
            // The warp's 32 x 32 block goes through shared memory instead and leaves as whole 128-byte row segments,
            // four rows per instruction.
            const uint32_t blk = smem_u32(smem + (size_t)kSt * SL::kBytes) + (uint32_t)(warp - 4) * kTrBlockBytes;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(blk + (uint32_t)lane * kTrRowBytes + (uint32_t)j * 4u),
                           "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3]) : "memory");
            __syncwarp();
            const uint32_t r0 = m0 + (uint32_t)quad * 32u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t rr = (uint32_t)i * 4u + ((uint32_t)lane >> 3);
              uint4 x;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                           : "r"(blk + rr * kTrRowBytes + ((uint32_t)lane & 7u) * 16u));
              if (r0 + rr < a.nq)
                *reinterpret_cast<uint4*>(a.out + (size_t)(r0 + rr) * a.ld + col + ((uint32_t)lane & 7u) * 4u) = x;
            }
          } else if (qvalid && a.out) {
            if (col + 32 <= a.row1 && (a.ld & 3u) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(dst + col + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < a.row1) dst[col + j] = __uint_as_float(v[j]);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

// ---- CTA-pair variant (cta_group::2) ---------------------------------------------------------
// Two CTAs of a cluster (one TPC) work on one 256 x 256 tile: each owns 128 A rows (its own
// accumulator half in its own TMEM) and loads only HALF of the B tile (128 rows); the leader CTA's
// single MMA thread issues tcgen05.mma.cta_group::2 (M = 256), which reads A and the B halves from both
// CTAs' shared memory.  Per SM this halves the B bytes fetched from L2 and written to / read from
// shared memory, which is what bounds the one-CTA kernel (60 B/clk/SM of TMA traffic against an L2
// that sustains ~42 B/clk/SM, plus operand reads on the same shared-memory port).
//   full_bar      leader's: armed by the leader's producer for BOTH CTAs' bytes; the peer's TMA
//                 completes its bytes on the leader's barrier (address with the peer bit cleared)
//   empty_bar     per CTA: tcgen05.commit multicast to both CTAs
//   tmem_full     per CTA: tcgen05.commit multicast to both CTAs
//   tmem_empty    leader's: 8 arrivals (4 epilogue warps x 2 CTAs), the peer arrives remotely
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same variable in CTA 0 of the pair
constexpr int kBHalfBytes = (BN / 2) * BK * 2;  // 16 KB

template <int kSplitsT>
struct StageLayout2 {
  static constexpr int kBytes = kSplitsT * kABytes + kBHalfBytes;
  static constexpr int kStagesT = kSplitsT == 1 ? 6 : 4;
  static constexpr size_t kSmem = (size_t)kStagesT * kBytes + 1024;
};

__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int kSplitsT, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, GemmArgs a) {
  using SL = StageLayout2<kSplitsT>;
  constexpr int kSt = SL::kStagesT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kSt], empty_bar[kSt], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_masks[kEpi == kEpiFilter ? BN / 32 : 1][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const uint32_t mt2 = (a.mt + 1) >> 1;  // A tiles are taken two at a time
  const uint32_t total_tiles = mt2 * a.nt;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBh)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer (both CTAs) ----
      uint32_t it = 0;
      for (uint32_t t = pair; t < total_tiles; t += npairs) {
        const uint32_t m0 = ((t % mt2) * 2 + rank) * BM, n0 = a.row0 + (t / mt2) * BN + rank * (BN / 2);
        for (uint32_t kb = 0; kb < a.num_kb; ++kb, ++it) {
          const int st = it % kSt;
          const uint32_t ph = (it / kSt) & 1;
          mbar_wait(&empty_bar[st], ph ^ 1);
          uint8_t* base = smem + (size_t)st * SL::kBytes;
          if (rank == 0) mbar_expect_tx(&full_bar[st], 2 * SL::kBytes);
          for (int s = 0; s < kSplitsT; ++s)
            tma_load_2d_pair(base + s * kABytes, &tmA, &full_bar[st], (int)(kb * BK), (int)(s * a.m_pad + m0));
          tma_load_2d_pair(base + kSplitsT * kABytes, &tmBh, &full_bar[st], (int)(kb * BK), (int)n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // ---- MMA issuer (leader CTA only) ----
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      uint32_t it = 0, lt = 0;
      long long w_empty = 0, w_full = 0;
      const long long t_begin = clock64();
      for (uint32_t t = pair; t < total_tiles; t += npairs, ++lt) {
        const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
        long long c0 = clock64();
        mbar_wait(&tmem_empty_bar[as], aph ^ 1);
        w_empty += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem + as * BN;
        for (uint32_t kb = 0; kb < a.num_kb; ++kb, ++it) {
          const int st = it % kSt;
          const uint32_t ph = (it / kSt) & 1;
          c0 = clock64();
          mbar_wait(&full_bar[st], ph);
          w_full += clock64() - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sbase = smem_u32(smem + (size_t)st * SL::kBytes);
          const uint32_t sB = sbase + kSplitsT * kABytes;
#pragma unroll
          for (int s = 0; s < kSplitsT; ++s) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(sbase + s * kABytes + k * 32);
              const uint64_t bd = umma_desc_sw128(sB + k * 32);
              umma_bf16_pair(tacc, ad, bd, idesc, (kb | (uint32_t)s | (uint32_t)k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_pair(&empty_bar[st]);
        }
        umma_commit_pair(&tmem_full_bar[as]);
      }
      if (a.prof && blockIdx.x == 0)
        printf("[gemm_pair] tiles %u kb %u: total %lld clk, wait tmem_empty %lld, wait full %lld\n", lt, a.num_kb,
               clock64() - t_begin, w_empty, w_full);
    }
  } else if (warp >= 4) {  // ---- epilogue (both CTAs, each its own 128 A rows) ----
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    uint32_t lt = 0;
    for (uint32_t t = pair; t < total_tiles; t += npairs, ++lt) {
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      const uint32_t m0 = ((t % mt2) * 2 + rank) * BM, n0 = a.row0 + (t / mt2) * BN;
      const uint32_t q = m0 + quad * 32 + lane;
      const bool qvalid = q < a.nq;
      const uint32_t tbase = tmem + as * BN + ((uint32_t)(quad * 32) << 16);
      if (kEpi == kEpiFilter) {
        mbar_wait(&tmem_full_bar[as], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        filter_tile(a, tbase, q, qvalid, n0, &s_masks[0][quad * 32 + lane], half * (BN / 64), (half + 1) * (BN / 64));
      } else {
        mbar_wait(&tmem_full_bar[as], aph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* dst = a.out + (size_t)q * a.ld;
#pragma unroll
        for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
          uint32_t v[32];
          tmem_ld32(tbase + (uint32_t)(c * 32), v);
          const uint32_t col = n0 + c * 32;
          if (qvalid && a.cmax && col < a.row1) {
            float mx = __int_as_float(0xFF800000);
            if (!a.cbias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < a.row1) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else if (col + 32 <= a.row1) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a.cbias + col + j));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j])), b.x));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j + 1])), b.y));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j + 2])), b.z));
                mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j + 3])), b.w));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < a.row1) mx = fmaxf(mx, __fsub_rn(__fmul_rn(2.f, __uint_as_float(v[j])), __ldg(a.cbias + col + j)));
            }
            a.cmax[(size_t)q * a.ld_c + (col >> 5)] = mx;
          }
          if (a.out && col + 32 <= a.row1 && (a.ld & 3u) == 0) {
            // A TMEM lane is an output ROW: stored straight from the registers, every instruction would write 16 bytes
            // into each of 32 rows (32 half-used sectors, the LSU queue was the limiter of the 40k-centre tokenization).
            // The warp's 32 x 32 block goes through shared memory instead and leaves as whole 128-byte row segments,
            // four rows per instruction.
            const uint32_t blk = smem_u32(smem + (size_t)kSt * SL::kBytes) + (uint32_t)(warp - 4) * kTrBlockBytes;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(blk + (uint32_t)lane * kTrRowBytes + (uint32_t)j * 4u),
                           "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3]) : "memory");
            __syncwarp();
            const uint32_t r0 = m0 + (uint32_t)quad * 32u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t rr = (uint32_t)i * 4u + ((uint32_t)lane >> 3);
              uint4 x;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                           : "r"(blk + rr * kTrRowBytes + ((uint32_t)lane & 7u) * 16u));
              if (r0 + rr < a.nq)
                *reinterpret_cast<uint4*>(a.out + (size_t)(r0 + rr) * a.ld + col + ((uint32_t)lane & 7u) * 4u) = x;
            }
          } else if (qvalid && a.out) {
            if (col + 32 <= a.row1 && (a.ld & 3u) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(dst + col + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < a.row1) dst[col + j] = __uint_as_float(v[j]);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[as]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // the peer's shared memory and barriers stay valid until both CTAs are done
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
  }
}

// f32 queries -> two stacked bf16 operands: hi = bf16(q), lo = bf16(q - hi); padded rows are zero.
__global__ void split_queries_kernel(const float* __restrict__ q, uint32_t nq, uint32_t d, uint32_t dp,
                                     uint32_t m_pad, __nv_bfloat16* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)m_pad * dp;
  if (i >= total) return;
  const uint32_t r = (uint32_t)(i / dp), c = (uint32_t)(i % dp);
  float v = (r < nq && c < d) ? q[(size_t)r * d + c] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  out[i] = hi;
  out[total + i] = lo;
}

__global__ void init_topk_state_kernel(uint32_t nq, uint32_t* cnt, uint64_t* tau, uint32_t* ovf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) { cnt[i] = 0; tau[i] = kKeyMax; ovf[i] = 0; }
}

// Is the candidate window wide enough?  The k' kept rows are the k' smallest APPROXIMATE keys (split-bf16 tensor-core
// scores), so every row outside has approximate distance >= a_last, the largest kept one, and exact distance
// >= a_last - eps with eps = eps_rel * ||q|| * max_i ||x_i|| bounding |approximate - exact| (dropped split terms
// <= 2^-15, fp32 accumulation of K products <= K * 2^-21 including the exact chain's own rounding, all relative to
// sum |q_d x_d| <= ||q|| ||x||; the same bound as the tokenization pre-filter).  If the k-th
// exact distance is strictly below that, no outside row can enter or tie into the top k.  Otherwise the query is
// flagged and the host widens the window (or finishes the query with the exact all-rows kernel).  Called by all
// threads after the final sort; sq = the query in shared memory.
struct SafetyArgs {
  float eps_rel, max_row_norm;
  uint32_t* unsafe;    // [nq] flag per query
  uint32_t* n_unsafe;  // counter
  // squared L2 (float brute force): the approximate keys live in the AUGMENTED dot-product space x' = [x, -||x||^2 / 2],
  // q' = [q, 1] (t = ||x||^2 / 2 - <q, x> = (L2 - ||q||^2) / 2), the exact keys are squared-L2 chain values
  int l2;
};
// ||q||^2 as SquaredL2Norm computes it for the many-to-many kernel (many_to_many_impl.inc:417-426): the four strided
// double accumulators of DenseSingleAccumulate, narrowed once (exact_math.cuh)
__device__ __forceinline__ float query_sqnorm_ref(const float* sq, uint32_t d) { return squared_l2_norm_strided(sq, d); }
__device__ __forceinline__ void check_window(const SafetyArgs& sa, uint32_t qi, const float* sq, uint32_t d, uint32_t m,
                                             uint32_t kprime, uint32_t kk, const uint64_t* src, const uint64_t* ka) {
  if (!sa.unsafe) return;
  __shared__ float s_part[4];
  float a = 0.f;
  for (uint32_t i = threadIdx.x; i < d; i += 128) a = fmaf(sq[i], sq[i], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float qn = sqrtf((s_part[0] + s_part[1]) + (s_part[2] + s_part[3])) * 1.0001f;
    bool bad = false;
    if (m == kprime && kk > 0 && qn > 0.f) {  // fewer than k' candidates = every row is one; zero query: all keys exact
      const float a_last = ord2f((uint32_t)(src[m - 1] >> 32));  // the buffer is sorted by the last compaction
      float e_k = ord2f((uint32_t)(ka[kk - 1] >> 32));
      float eps = sa.eps_rel * qn * sa.max_row_norm;
      if (sa.l2) {
        // augmented query norm sqrt(||q||^2 + 1); the exact key back in the augmented scale; plus the rounding of the
        // squared-L2 chain itself, (d + 4) ulps of (||q|| + ||x||)^2 (||x|| <= ||x'|| = max_row_norm)
        const float qa = sqrtf(qn * qn + 1.0f) * 1.0001f;
        eps = sa.eps_rel * qa * sa.max_row_norm + (float)(d + 4) * 1.2e-7f * (qn + sa.max_row_norm) * (qn + sa.max_row_norm);
        e_k = 0.5f * (e_k - qn * qn * 0.9998f);
      }
      bad = !(e_k < a_last - eps);
    }
    sa.unsafe[qi] = bad ? 1u : 0u;
    if (bad) atomicAdd(sa.n_unsafe, 1u);
  }
}

// Exact re-scoring of the k' candidates (f32 query x bf16 row, the oracle's 8-lane FMA order),
// final top-k by (distance, id), output.
__global__ void __launch_bounds__(128)
rescore_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ db, uint32_t d, uint32_t dpitch,
               const uint64_t* __restrict__ buf, const uint32_t* __restrict__ cnt, uint32_t cap, uint32_t kprime,
               uint32_t k, uint32_t out_k, uint32_t id_base, uint32_t* __restrict__ out_idx, float* __restrict__ out_dist,
               int np2, SafetyArgs sa) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* ka = reinterpret_cast<uint64_t*>(smem);
  float* sq = reinterpret_cast<float*>(ka + np2);
  const int tid = threadIdx.x;
  const uint32_t qi = blockIdx.x;
  const uint32_t m = min(cnt[qi], kprime);
  const uint64_t* src = buf + (size_t)qi * cap;
  for (uint32_t i = tid; i < d; i += 128) sq[i] = q[(size_t)qi * d + i];
  for (int i = tid; i < np2; i += 128) ka[i] = kKeyMax;
  __syncthreads();
  const int l = tid & 7, grp = tid >> 3;
  for (uint32_t c0 = 0; c0 < m; c0 += 16) {
    const uint32_t c = c0 + grp;
    const bool valid = c < m;
    const uint32_t dp = (uint32_t)src[valid ? c : 0];
    const __nv_bfloat16* x = db + (size_t)dp * dpitch;
    float acc = 0.f;
    uint32_t j = 0;
    for (; j + 8 <= d; j += 8) acc = __fmaf_rn(-sq[j + l], __bfloat162float(x[j + l]), acc);
    // the reference's asymmetric kernel (exact_math.cuh neg_dot_asym_order): 4-wide step into lanes 0..3 before the sum
    if (j + 4 <= d) { if (l < 4) acc = __fmaf_rn(-sq[j + l], __bfloat162float(x[j + l]), acc); j += 4; }
    const float b = __fadd_rn(acc, __shfl_down_sync(0xFFFFFFFFu, acc, 4, 8));
    const float t2 = __fadd_rn(b, __shfl_down_sync(0xFFFFFFFFu, b, 2, 8));
    float r = __fadd_rn(t2, __shfl_down_sync(0xFFFFFFFFu, t2, 1, 8));
    if (l == 0) for (; j < d; ++j) r = __fmaf_rn(-sq[j], __bfloat162float(x[j]), r);
    if (valid && l == 0) ka[c] = make_key(r, dp);
  }
  __syncthreads();
  block_bitonic_sort(ka, np2);
  const uint32_t kk = min(k, m);
  check_window(sa, qi, sq, d, m, kprime, kk, src, ka);
  for (uint32_t i = tid; i < out_k; i += 128) {
    uint32_t id = 0;
    float dist = __uint_as_float(0x7FC00000u);
    if (i < kk) { id = (uint32_t)ka[i] + id_base; dist = -ord2f((uint32_t)(ka[i] >> 32)); }
    out_idx[(size_t)qi * out_k + i] = id;
    out_dist[(size_t)qi * out_k + i] = dist;
  }
}

// Float brute force (BruteForceSearcher<float>::FinishBatchedSearchSimple, brute_force/brute_force.cc:376-393 ->
// DenseDistanceManyToManyTopK, many_to_many_impl.inc:522-567): exact re-scoring of the k' candidates with the
// reference's arithmetic, acc = 0; acc = fnmadd(q[d], x[d], acc) sequentially in d.  One thread per candidate.
// xnorm != NULL: squared L2 (brute_force.cc:376-393 with SquaredL2Distance -> many_to_many_impl.inc:236-257,522-567):
// acc = ||x||^2 + ||q||^2, then acc = fnmadd(q[d], 2 x[d], acc) sequentially in d; the distance is acc itself.
__global__ void __launch_bounds__(128)
rescore_f32_kernel(const float* __restrict__ q, const float* __restrict__ db, uint32_t d, const uint64_t* __restrict__ buf,
                   const uint32_t* __restrict__ cnt, uint32_t cap, uint32_t kprime, uint32_t k, uint32_t out_k,
                   uint32_t id_base, uint32_t* __restrict__ out_idx, float* __restrict__ out_dist, int np2, SafetyArgs sa,
                   const float* __restrict__ xnorm) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* ka = reinterpret_cast<uint64_t*>(smem);
  float* sq = reinterpret_cast<float*>(ka + np2);
  __shared__ float s_qn;
  const int tid = threadIdx.x;
  const uint32_t qi = blockIdx.x;
  const uint32_t m = min(cnt[qi], kprime);
  const uint64_t* src = buf + (size_t)qi * cap;
  for (uint32_t i = tid; i < d; i += 128) sq[i] = q[(size_t)qi * d + i];
  for (int i = tid; i < np2; i += 128) ka[i] = kKeyMax;
  __syncthreads();
  if (xnorm && tid == 0) s_qn = query_sqnorm_ref(sq, d);
  __syncthreads();
  const float two = xnorm ? 2.0f : 1.0f;  // 2 x is exact
  for (uint32_t c = tid; c < m; c += 128) {
    const uint32_t dp = (uint32_t)src[c];
    const float* x = db + (size_t)dp * d;
    float acc = xnorm ? __fadd_rn(xnorm[dp], s_qn) : 0.f;
    if ((d & 3u) == 0) {
      for (uint32_t j = 0; j < d; j += 16) {  // 16 dims in flight before the first FMA of the chunk
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(x + min(j + 4 * u, d - 4)));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t jj = j + 4 * u;
          if (jj < d) {
            acc = __fmaf_rn(-sq[jj], __fmul_rn(v[u].x, two), acc);
            acc = __fmaf_rn(-sq[jj + 1], __fmul_rn(v[u].y, two), acc);
            acc = __fmaf_rn(-sq[jj + 2], __fmul_rn(v[u].z, two), acc);
            acc = __fmaf_rn(-sq[jj + 3], __fmul_rn(v[u].w, two), acc);
          }
        }
      }
    } else {
      for (uint32_t j = 0; j < d; ++j) acc = __fmaf_rn(-sq[j], __fmul_rn(__ldg(x + j), two), acc);
    }
    ka[c] = make_key(acc, dp);
  }
  __syncthreads();
  block_bitonic_sort(ka, np2);
  const uint32_t kk = min(k, m);
  check_window(sa, qi, sq, d, m, kprime, kk, src, ka);
  for (uint32_t i = tid; i < out_k; i += 128) {
    uint32_t id = 0;
    float dist = __uint_as_float(0x7FC00000u);
    if (i < kk) { id = (uint32_t)ka[i] + id_base; dist = xnorm ? ord2f((uint32_t)(ka[i] >> 32)) : -ord2f((uint32_t)(ka[i] >> 32)); }
    out_idx[(size_t)qi * out_k + i] = id;
    out_dist[(size_t)qi * out_k + i] = dist;
  }
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
// max_i ||x_i||^2 over the database rows (bf16 or f32), as the bits of a non-negative float (atomicMax on u32).
template <typename T>
__global__ void max_row_norm_kernel(const T* __restrict__ db, uint32_t n, uint32_t d, uint32_t pitch, uint32_t* out_bits) {
  const uint32_t row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float a = 0.f;
  if (row < n) {
    const T* x = db + (size_t)row * pitch;
    for (uint32_t j = lane; j < d; j += 32) { const float v = to_float(x[j]); a = fmaf(v, v, a); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
  if (lane == 0 && row < n) atomicMax(out_bits, __float_as_uint(a));
}

// Exact all-rows fallback for the queries whose candidate window could not be proven wide enough even at the largest
// k' (thousands of near-ties at the k-th score).  One round = rows [row0, row1) against the flagged queries: eight
// lanes per row evaluate the reference's exact arithmetic, keys below the query's tau go to its candidate buffer; the
// host compacts (keep k) between rounds and sizes the rounds so that the buffer cannot overflow.
template <bool kF32>
__global__ void __launch_bounds__(128)
exact_round_kernel(const float* __restrict__ q, const void* __restrict__ dbv, uint32_t d, uint32_t dpitch,
                   const uint32_t* __restrict__ flagged, uint32_t row0, uint32_t row1, uint64_t* __restrict__ buf,
                   uint32_t* __restrict__ cnt, const uint64_t* __restrict__ tau, uint32_t* __restrict__ ovf, uint32_t cap,
                   const float* __restrict__ xnorm) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* sq = reinterpret_cast<float*>(smem);
  __shared__ float s_qn;
  const int tid = threadIdx.x;
  const uint32_t qi = flagged[blockIdx.y];
  for (uint32_t i = tid; i < d; i += 128) sq[i] = q[(size_t)qi * d + i];
  __syncthreads();
  if (kF32 && xnorm && tid == 0) s_qn = query_sqnorm_ref(sq, d);
  __syncthreads();
  const uint64_t t = tau[qi];
  const int l = tid & 7, grp = tid >> 3;
  const uint32_t per_cta = (row1 - row0 + gridDim.x - 1) / gridDim.x;
  const uint32_t r_begin = row0 + blockIdx.x * per_cta, r_end = min(row1, r_begin + per_cta);
  for (uint32_t r0 = r_begin; r0 < r_end; r0 += 16) {
    const uint32_t row = r0 + grp;
    const bool valid = row < r_end;
    float r = 0.f;
    if (kF32) {
      // BruteForceSearcher<float>: acc = fnmadd(q[d], x[d], acc) sequentially in d -- one lane walks the row
      if (valid && l == 0) {
        const float* x = static_cast<const float*>(dbv) + (size_t)row * d;
        if (xnorm) {  // squared L2: ||x||^2 + ||q||^2, then fnmadd(q, 2 x)
          r = __fadd_rn(xnorm[row], s_qn);
          for (uint32_t j = 0; j < d; ++j) r = __fmaf_rn(-sq[j], __fmul_rn(__ldg(x + j), 2.0f), r);
        } else {
          for (uint32_t j = 0; j < d; ++j) r = __fmaf_rn(-sq[j], __ldg(x + j), r);
        }
      }
    } else {
      const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(dbv) + (size_t)(valid ? row : row0) * dpitch;
      float acc = 0.f;
      uint32_t j = 0;
      for (; j + 8 <= d; j += 8) acc = __fmaf_rn(-sq[j + l], __bfloat162float(x[j + l]), acc);
      if (j + 4 <= d) { if (l < 4) acc = __fmaf_rn(-sq[j + l], __bfloat162float(x[j + l]), acc); j += 4; }
      const float b = __fadd_rn(acc, __shfl_down_sync(0xFFFFFFFFu, acc, 4, 8));
      const float t2 = __fadd_rn(b, __shfl_down_sync(0xFFFFFFFFu, b, 2, 8));
      r = __fadd_rn(t2, __shfl_down_sync(0xFFFFFFFFu, t2, 1, 8));
      if (l == 0) for (; j < d; ++j) r = __fmaf_rn(-sq[j], __bfloat162float(x[j]), r);
    }
    if (valid && l == 0) {
      const uint64_t key = make_key(r, row);
      if (key < t) {
        const uint32_t pos = atomicAdd(&cnt[qi], 1u);
        if (pos < cap) buf[(size_t)qi * cap + pos] = key;
        else ovf[qi] = 1u;
      }
    }
  }
}

// flagged queries: reset the candidate state; all others: park (cnt = 0, so the compactions skip them)
__global__ void exact_prepare_kernel(uint32_t nq, const uint32_t* __restrict__ unsafe, uint32_t* cnt, uint64_t* tau,
                                     uint32_t* ovf, uint32_t* flagged, uint32_t* n_flagged) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  cnt[i] = 0; ovf[i] = 0; tau[i] = kKeyMax;
  if (unsafe[i]) flagged[atomicAdd(n_flagged, 1u)] = i;
}

__global__ void exact_emit_kernel(const uint32_t* __restrict__ flagged, const uint64_t* __restrict__ buf,
                                  const uint32_t* __restrict__ cnt, uint32_t cap, uint32_t k, uint32_t out_k,
                                  uint32_t id_base, uint32_t* __restrict__ out_idx, float* __restrict__ out_dist, int l2) {
  const uint32_t qi = flagged[blockIdx.x];
  const uint32_t kk = min(k, cnt[qi]);
  for (uint32_t i = threadIdx.x; i < out_k; i += blockDim.x) {
    uint32_t id = 0;
    float dist = __uint_as_float(0x7FC00000u);
    if (i < kk) {
      const uint64_t key = buf[(size_t)qi * cap + i];  // sorted by the last compaction
      id = (uint32_t)key + id_base; dist = l2 ? ord2f((uint32_t)(key >> 32)) : -ord2f((uint32_t)(key >> 32));
    }
    out_idx[(size_t)qi * out_k + i] = id;
    out_dist[(size_t)qi * out_k + i] = dist;
  }
}

// ---- squared-L2 float brute force: row norms and the augmented rows of the pre-filter ----
// ||x||^2 as the many-to-many transposer computes it (AugmentWithL2Norms, many_to_many_impl.inc:236-257):
// norm = fnmadd(x, x, norm) in dimension order, times -1.
__global__ void row_sqnorm_ref_kernel(const float* __restrict__ x, uint32_t n, uint32_t d, float* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* p = x + (size_t)r * d;
  float a = 0.f;
  for (uint32_t j = 0; j < d; ++j) { const float v = __ldg(p + j); a = __fmaf_rn(-v, v, a); }
  out[r] = __fmul_rn(a, -1.0f);
}
// out [n][d + 1]: the row, then `last` (queries: 1) or -norm[r] / 2 (database rows)
__global__ void augment_rows_kernel(const float* __restrict__ x, const float* __restrict__ norm, float last, uint32_t n,
                                    uint32_t d, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)n * (d + 1);
  if (i >= total) return;
  const size_t r = i / (d + 1);
  const uint32_t c = (uint32_t)(i - r * (d + 1));
  out[i] = c < d ? x[r * d + c] : (norm ? -0.5f * norm[r] : last);
}

}  // namespace bf

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return (EncodeTiledFn) nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

static cudaError_t make_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                             uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)bf::BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// A rows are padded to whole CTA pairs (2 x 128 rows)
static uint32_t bf_m_pad(uint32_t nq) { return (nq + 2 * bf::BM - 1) / (2 * bf::BM) * (2 * bf::BM); }

size_t bf_query_operand_bytes(uint32_t nq, uint32_t dpitch) {
  const uint32_t m_pad = bf_m_pad(nq);
  return (size_t)bf::kSplits * m_pad * dpitch * 2;
}

cudaError_t bf_split_queries(const float* q, uint32_t nq, uint32_t d, uint32_t dpitch, void* a_operand, cudaStream_t s) {
  const uint32_t m_pad = bf_m_pad(nq);
  const size_t total = (size_t)m_pad * dpitch;
  bf::split_queries_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(q, nq, d, dpitch, m_pad,
                                                                           reinterpret_cast<__nv_bfloat16*>(a_operand));
  return cudaGetLastError();
}

cudaError_t bf_init_state(uint32_t nq, uint32_t* cnt, uint64_t* tau, uint32_t* ovf, cudaStream_t s) {
  bf::init_topk_state_kernel<<<(nq + 255) / 256, 256, 0, s>>>(nq, cnt, tau, ovf);
  return cudaGetLastError();
}

static int sm_count() {
  static int sms = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  return sms;
}

template <int kSplitsT, int kEpi>
static cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const bf::GemmArgs& a, cudaStream_t s) {
  using SL = bf::StageLayout<kSplitsT, kEpi>;
  cudaError_t e = cudaFuncSetAttribute(bf::gemm_kernel<kSplitsT, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SL::kSmem);
  if (e != cudaSuccess) return e;
  const uint32_t tiles = a.mt * a.nt;
  if (tiles == 0) return cudaSuccess;
  const uint32_t grid = tiles < (uint32_t)sm_count() ? tiles : (uint32_t)sm_count();  // persistent, one CTA per SM
  bf::gemm_kernel<kSplitsT, kEpi><<<grid, bf::kThreads, SL::kSmem, s>>>(tmA, tmB, a);
  return cudaGetLastError();
}

template <int kSplitsT, int kEpi>
static cudaError_t launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmBh, const bf::GemmArgs& a, cudaStream_t s) {
  using SL = bf::StageLayout2<kSplitsT>;
  cudaError_t e = cudaFuncSetAttribute(bf::gemm_pair_kernel<kSplitsT, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)SL::kSmem);
  if (e != cudaSuccess) return e;
  const uint32_t tiles = ((a.mt + 1) / 2) * a.nt;
  if (tiles == 0) return cudaSuccess;
  const uint32_t pairs = (uint32_t)sm_count() / 2;
  const uint32_t grid = 2 * (tiles < pairs ? tiles : pairs);  // persistent, one CTA pair per TPC
  bf::gemm_pair_kernel<kSplitsT, kEpi><<<grid, bf::kThreads, SL::kSmem, s>>>(tmA, tmBh, a);
  return cudaGetLastError();
}

static bool use_pair_kernel() {
  const char* e = getenv("SCANN_B200_GEMM");
  return !(e && !strcmp(e, "1cta"));
}

// One round: database rows [row0, row1) against all queries.
cudaError_t bf_gemm_round(const void* a_operand, const void* db, uint32_t nq, uint32_t n_total, uint32_t dpitch,
                          uint32_t row0, uint32_t row1, const ScanWork& w, int splits, cudaStream_t s) {
  const uint32_t m_pad = bf_m_pad(nq);
  const bool pair = use_pair_kernel();
  CUtensorMap tmA, tmB;
  cudaError_t e = make_tmap(&tmA, a_operand, (uint64_t)splits * m_pad, dpitch, dpitch, bf::BM);
  if (e != cudaSuccess) return e;
  e = make_tmap(&tmB, db, n_total, dpitch, dpitch, pair ? bf::BN / 2 : bf::BN);
  if (e != cudaSuccess) return e;
  bf::GemmArgs a{};
  a.nq = nq; a.m_pad = m_pad; a.row0 = row0; a.row1 = row1;
  a.mt = m_pad / bf::BM; a.nt = (row1 - row0 + bf::BN - 1) / bf::BN;
  a.num_kb = (dpitch + bf::BK - 1) / bf::BK;
  a.buf = w.buf; a.cnt = w.cnt; a.tau = w.tau; a.ovf = w.ovf; a.cap = w.cap;
  { const char* pe = getenv("SCANN_B200_GEMM_PROFILE"); a.prof = pe && pe[0] == '1'; }
  if (splits == 1) {  // operands with the hi/lo terms concatenated along K (float brute force)
    if (pair) return launch_gemm_pair<1, bf::kEpiFilter>(tmA, tmB, a, s);
    return launch_gemm<1, bf::kEpiFilter>(tmA, tmB, a, s);
  }
  if (pair) return launch_gemm_pair<bf::kSplits, bf::kEpiFilter>(tmA, tmB, a, s);
  return launch_gemm<bf::kSplits, bf::kEpiFilter>(tmA, tmB, a, s);
}

uint32_t bf_query_rows_pad(uint32_t nq) { return bf_m_pad(nq); }

// Plain C[a_row][b_row] = sum_k A[a_row][k] * B[b_row][k] (bf16 operands, fp32 accumulate and output): the
// tokenization pre-filter's GEMM (prep.cu), K = the concatenated hi/lo terms.
cudaError_t gemm_bf16_nt(const void* a_operand, uint32_t a_rows, uint32_t a_rows_pad, const void* b_operand,
                         uint32_t b_rows, uint32_t kpitch, float* out, uint32_t ld, cudaStream_t s, float* cmax,
                         uint32_t ld_c, const float* cbias) {
  // SCANN_B200_TOK_GEMM=pair: the CTA-pair kernel (each CTA loads half of the B tile) for this GEMM too; at K = 320
  // the one-CTA kernel is bound by its operand traffic from L2 (tokenization)
  bool pair = false;
  { const char* pe = getenv("SCANN_B200_TOK_GEMM"); pair = pe && !strcmp(pe, "pair"); }
  CUtensorMap tmA, tmB;
  cudaError_t e = make_tmap(&tmA, a_operand, a_rows_pad, kpitch, kpitch, bf::BM);
  if (e != cudaSuccess) return e;
  e = make_tmap(&tmB, b_operand, b_rows, kpitch, kpitch, pair ? bf::BN / 2 : bf::BN);
  if (e != cudaSuccess) return e;
  bf::GemmArgs a{};
  a.nq = a_rows; a.m_pad = a_rows_pad; a.row0 = 0; a.row1 = b_rows;
  a.mt = a_rows_pad / bf::BM; a.nt = (b_rows + bf::BN - 1) / bf::BN;
  a.num_kb = (kpitch + bf::BK - 1) / bf::BK;
  a.out = out; a.ld = ld;
  a.cmax = cmax; a.ld_c = ld_c; a.cbias = cbias;
  if (pair) return launch_gemm_pair<1, bf::kEpiStore>(tmA, tmB, a, s);
  return launch_gemm<1, bf::kEpiStore>(tmA, tmB, a, s);
}

cudaError_t bf_max_row_norm(const void* db, bool f32, uint32_t n, uint32_t d, uint32_t pitch, float* out, cudaStream_t s) {
  uint32_t* bits = nullptr;
  cudaError_t e = cudaMalloc(&bits, 4);
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(bits, 0, 4, s);
  if (n) {
    if (f32) bf::max_row_norm_kernel<float><<<(n + 7) / 8, 256, 0, s>>>(static_cast<const float*>(db), n, d, pitch, bits);
    else bf::max_row_norm_kernel<__nv_bfloat16><<<(n + 7) / 8, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(db), n, d, pitch, bits);
  }
  float sq = 0.f;
  e = cudaMemcpyAsync(&sq, bits, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(bits);
  if (e != cudaSuccess) return e;
  *out = sqrtf(sq) * 1.0001f;
  return cudaGetLastError();
}

cudaError_t bf_exact_prepare(uint32_t nq, const uint32_t* unsafe, const ScanWork& w, uint32_t* flagged, uint32_t* n_flagged,
                             cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(n_flagged, 0, 4, s);
  if (e != cudaSuccess) return e;
  bf::exact_prepare_kernel<<<(nq + 255) / 256, 256, 0, s>>>(nq, unsafe, w.cnt, w.tau, w.ovf, flagged, n_flagged);
  return cudaGetLastError();
}

cudaError_t bf_row_sqnorms(const float* x, uint32_t n, uint32_t d, float* out, cudaStream_t s) {
  if (!n) return cudaSuccess;
  bf::row_sqnorm_ref_kernel<<<(n + 127) / 128, 128, 0, s>>>(x, n, d, out);
  return cudaGetLastError();
}
cudaError_t bf_augment_rows(const float* x, const float* norm, float last, uint32_t n, uint32_t d, float* out, cudaStream_t s) {
  if (!n) return cudaSuccess;
  const size_t total = (size_t)n * (d + 1);
  bf::augment_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, norm, last, n, d, out);
  return cudaGetLastError();
}

cudaError_t bf_exact_round(const float* q, const void* db, bool f32, uint32_t d, uint32_t dpitch, const uint32_t* flagged,
                           uint32_t n_flagged, uint32_t row0, uint32_t row1, const ScanWork& w, cudaStream_t s,
                           const float* xnorm) {
  if (!n_flagged || row1 <= row0) return cudaSuccess;
  const size_t smem = (((size_t)d + 3) & ~(size_t)3) * 4;
  const uint32_t ctas = std::min<uint32_t>(std::max<uint32_t>(1u, (row1 - row0 + 1023) / 1024), 1024u);
  dim3 grid(ctas, n_flagged);
  if (f32) {
    cudaError_t e = cudaFuncSetAttribute(bf::exact_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bf::exact_round_kernel<true><<<grid, 128, smem, s>>>(q, db, d, dpitch, flagged, row0, row1, w.buf, w.cnt, w.tau, w.ovf, w.cap, xnorm);
  } else {
    cudaError_t e = cudaFuncSetAttribute(bf::exact_round_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bf::exact_round_kernel<false><<<grid, 128, smem, s>>>(q, db, d, dpitch, flagged, row0, row1, w.buf, w.cnt, w.tau, w.ovf, w.cap, nullptr);
  }
  return cudaGetLastError();
}

cudaError_t bf_exact_emit(const uint32_t* flagged, uint32_t n_flagged, const ScanWork& w, uint32_t k, uint32_t out_k,
                          uint32_t id_base, uint32_t* out_idx, float* out_dist, cudaStream_t s, bool l2) {
  if (!n_flagged) return cudaSuccess;
  bf::exact_emit_kernel<<<n_flagged, 128, 0, s>>>(flagged, w.buf, w.cnt, w.cap, k, out_k, id_base, out_idx, out_dist, l2 ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t bf_rescore(const float* q, const void* db, uint32_t nq, uint32_t d, uint32_t dpitch, const ScanWork& w,
                       uint32_t kprime, uint32_t k, uint32_t out_k, uint32_t id_base, uint32_t* out_idx, float* out_dist,
                       cudaStream_t s, const BfSafety* safety) {
  bf::SafetyArgs sa{};
  if (safety) sa = bf::SafetyArgs{safety->eps_rel, safety->max_row_norm, safety->unsafe, safety->n_unsafe, 0};
  int np2 = 2;
  while ((uint32_t)np2 < kprime) np2 <<= 1;
  const size_t smem = (size_t)np2 * 8 + (((size_t)d + 3) & ~(size_t)3) * 4;
  cudaError_t e = cudaFuncSetAttribute(bf::rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  bf::rescore_kernel<<<nq, 128, smem, s>>>(q, reinterpret_cast<const __nv_bfloat16*>(db), d, dpitch, w.buf, w.cnt, w.cap,
                                           kprime, k, out_k, id_base, out_idx, out_dist, np2, sa);
  return cudaGetLastError();
}

cudaError_t bf_rescore_f32(const float* q, const float* db, uint32_t nq, uint32_t d, const ScanWork& w, uint32_t kprime,
                           uint32_t k, uint32_t out_k, uint32_t id_base, uint32_t* out_idx, float* out_dist, cudaStream_t s,
                           const BfSafety* safety, const float* xnorm) {
  bf::SafetyArgs sa{};
  if (safety) sa = bf::SafetyArgs{safety->eps_rel, safety->max_row_norm, safety->unsafe, safety->n_unsafe, xnorm ? 1 : 0};
  int np2 = 2;
  while ((uint32_t)np2 < kprime) np2 <<= 1;
  const size_t smem = (size_t)np2 * 8 + (((size_t)d + 3) & ~(size_t)3) * 4;
  cudaError_t e = cudaFuncSetAttribute(bf::rescore_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  bf::rescore_f32_kernel<<<nq, 128, smem, s>>>(q, db, d, w.buf, w.cnt, w.cap, kprime, k, out_k, id_base, out_idx, out_dist, np2, sa, xnorm);
  return cudaGetLastError();
}

}  // namespace sb
