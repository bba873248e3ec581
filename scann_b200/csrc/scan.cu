// scan.cu -- the LUT16 scan: 4-bit codes x uint8 LUT, int16 accumulate, fused threshold top-N.
//
// Replaces (a7/a8/a9) of SURVEY.md section 8:
//   TreeAHHybridResidual::FindNeighborsBatchedImpl     tree_x_hybrid/tree_ah_hybrid_residual.cc:631-786
//   LUT16Avx2<>::GetTopFloatDistances / BottomLoop     hashes/internal/lut16_avx2.inc:55-124,404-527
//   FastTopNeighbors<float>                            utils/fast_top_neighbors.h:43-299
//
// B200 formulation.  The CPU kernel looks one query's 16-entry uint8 table up with vpshufb
// for 32 datapoints at a time.  Here the unit of work of the main scan is an OCT of eight queries
// that probe the same leaf: their eight uint8 tables are interleaved in shared memory into one
// table of 64-bit entries  T[b][c] = {lut0 | lut1 << 8 | lut2 << 16 | lut3 << 24, lut4 | ... | lut7 << 24},
// so ONE conflict-free LDS.64 (16 distinct entries = 16 distinct bank pairs, equal entries
// broadcast) serves eight (query, datapoint, block) lookups; masks and byte permutes split the
// two words into four registers of two u16 lanes and four IMADs accumulate them (B <= 256 =>
// sum <= 65280, no carry between the halves).  Each thread owns one datapoint of a 32-slot
// group: its B nibbles live in W = ceil(B/8) registers, loaded once per work item with coalesced
// 128-bit loads and reused for every oct of the item.  (The pilot and the debug hook score one
// query at a time through the 32-bit "quad" table of score_quad.)
//
// Top-N.  Every query owns a candidate buffer buf[q][cap] in HBM, a count and a threshold
// key tau[q] (score, global slot).  A PILOT kernel (one CTA per query) scores whole leaves,
// nearest first, until it has seen >= 4 N slots and sets tau to (a bound of) the N-th best key
// of that sample + 1; it publishes no candidates.  The MAIN kernel scans every probed (query, leaf)
// pair, grouped by leaf into octs, with an integer pre-filter (sum <= thr, conservative) and an
// exact 64-bit key comparison (key < tau); survivors are appended with one warp-aggregated atomic.
// COMPACT keeps the N smallest keys of each buffer (sort, or bound select + sort for long buffers).
// The result is the exact top-N under (float score, leaf, slot) -- the contract of SURVEY.md
// section 7 hard-part 1 -- regardless of scheduling and of how tight tau was.  If a buffer
// overflowed, the host re-scans only the affected queries with the tightened tau (duplicates are
// removed by key), so no candidate can be lost.
#include <utility>

#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"
#include "lut.cuh"
#include "scan_common.cuh"

// How the main scan accumulates the four u16-lane registers of an oct lookup:
//   0 = plain adds (ptxas merges pairs into IADD3 on the ALU pipe), 1 = two on ALU + two IMAD on the
//   FMA pipe, 2 = all four as IMAD, 3 = one mask per word + a 64-bit IMAD.WIDE sum of the whole word (fewer
//   instructions, but IMAD.WIDE issues far below the IMAD rate: 1.94 ms vs 1.45 ms on C2).
//   Measured on B200 (C2): see DESIGN.md.
#ifndef SB_SCAN_ACC
#define SB_SCAN_ACC 2
#endif
// Accumulation of a wide-quad lookup (two registers of two u16 lanes): 0 = plain adds, 1 = one add + one IMAD,
// 2 = two IMADs.
// resident CTAs per SM the main scan is compiled for (5 x 128 threads => at most 96 registers)
#ifndef SB_SCAN_MIN_CTAS
#define SB_SCAN_MIN_CTAS 5
#endif
#ifndef SB_SCAN_WACC
#define SB_SCAN_WACC 1
#endif

namespace sb {

constexpr uint32_t kInvalidQuery = 0xFFFFFFFFu;

// Work tiling of a leaf of `ng` 32-slot groups: tiles of at most `max_gpt` groups, evenly sized.  max_gpt is a per-batch
// choice of the host (launch heuristics in index.cu): 32 keeps items small where there are few of them, larger tiles
// amortise the per-item set-up (oct tables, thresholds) where a batch has many more items than resident CTAs.
__device__ __forceinline__ void leaf_tiling(uint32_t ng, uint32_t max_gpt, uint32_t* ntiles, uint32_t* gpt) {
  const uint32_t nt = ng ? (ng + max_gpt - 1) / max_gpt : 0u;
  *ntiles = nt;
  *gpt = nt ? (ng + nt - 1) / nt : 0u;
}
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kMaxQPI = 32;  // queries per work item (4 octs of 8 queries)

// Four queries' u16 sums for one datapoint: a01 = s0 | s1 << 16, a23 = s2 | s3 << 16.
// NL = number of real blocks in the last word: 1..8 compile-time (the padded lookups vanish from
// the instruction stream), 0 = use the runtime `nlast` (predicated lookups).
template <int W, int NL = 0>
__device__ __forceinline__ void score_quad(const uint32_t (&w)[W], const unsigned char* tbl,
                                           int nlast, uint32_t& a01, uint32_t& a23) {
  uint32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  const int nl = NL ? NL : nlast;
#pragma unroll
  for (int j = 0; j < W; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (j == W - 1 && k >= nl) continue;
      const uint32_t off = (k == 0) ? ((w[j] << 2) & 0x3Cu) : ((w[j] >> (4 * k - 2)) & 0x3Cu);
      const uint32_t v = *reinterpret_cast<const uint32_t*>(tbl + (8 * j + k) * 64 + off);
      const uint32_t e = v & 0x00FF00FFu, o = __byte_perm(v, 0u, 0x4341);  // (q0,q2) and (q1,q3) as u16 lanes
      if (k & 1) { x1 += e; y1 += o; } else { x0 += e; y0 += o; }
    }
  }
  const uint32_t ae = x0 + x1, ao = y0 + y1;  // ae = s0 | s2 << 16, ao = s1 | s3 << 16
  a01 = __byte_perm(ae, ao, 0x5410);
  a23 = __byte_perm(ae, ao, 0x7632);
}

// Pilot scoring: ONE query, its u8 LUT as it is ([8W][16] bytes).  One LDS.U8 and one add per lookup (the 16 bytes of a
// block row lie in four adjacent banks: no conflicts) instead of the quad table's load + mask + permute + two adds.
template <int W, int NL = 0>
__device__ __forceinline__ int score_one(const uint32_t (&w)[W], const unsigned char* tbl, int nlast) {
  uint32_t s0 = 0, s1 = 0;
  const int nl = NL ? NL : nlast;
#pragma unroll
  for (int j = 0; j < W; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (j == W - 1 && k >= nl) continue;
      const uint32_t nib = (w[j] >> (4 * k)) & 0xFu;
      const uint32_t v = tbl[(8 * j + k) * 16 + nib];
      if (k & 1) s1 += v; else s0 += v;
    }
  }
  return (int)(s0 + s1);
}

// Runtime-W pilot scoring (generic kernels): the code words come straight from global memory, one at a time.
__device__ __forceinline__ int score_one_rt(const uint32_t* __restrict__ gbase, int lane, int W, const unsigned char* tbl, int nlast) {
  uint32_t s0 = 0, s1 = 0;
#pragma unroll 1
  for (int j = 0; j < W; ++j) {
    const uint32_t word = load_code_word(gbase, lane, W, j);
    const int nk = j == W - 1 ? nlast : 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < nk) {
        const uint32_t v = tbl[(8 * j + k) * 16 + ((word >> (4 * k)) & 0xFu)];
        if (k & 1) s1 += v; else s0 += v;
      }
    }
  }
  return (int)(s0 + s1);
}

// Main-scan scoring.  An OCT of eight queries shares one table of 64-bit entries
//   T[b][c] = { lut0 | lut1 << 8 | lut2 << 16 | lut3 << 24,  lut4 | lut5 << 8 | lut6 << 16 | lut7 << 24 },
// so one LDS.64 (the LSU issues about one warp-wide shared load per two cycles whatever its width)
// serves EIGHT (query, datapoint, block) lookups.  Two masks and two byte permutes (ALU pipe) split
// the pair of words into four registers of two u16 lanes, four IMADs (FMA pipe, `one` is a runtime 1
// so ptxas cannot fold them back into IADD3s) accumulate them: LDS : ALU : FMA = 1 : 4 : 4 per eight
// lookups.  ad[i] is the shared address of block row i's entry for this slot in oct table 0; oct QD's
// table is a compile-time displacement away.
template <int... Is, class F>
__device__ __forceinline__ void static_for(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}

// One oct (eight queries) against one 32-slot group.  ad[i] = table base | nibble offset of lookup i (the
// table is 128-byte aligned, so the OR is the add); the row of lookup i (i * 128 bytes) and the oct's table
// (QD * W * 1024 bytes) are immediate displacements of the load.
template <int W, int NL, int QD>
__device__ __forceinline__ void score_oct_addr(const uint32_t (&ad)[8 * W], int nlast, uint32_t one,
                                               uint32_t (&acc)[4]) {
  uint32_t e0 = 0, o0 = 0, e1 = 0, o1 = 0;
  const int nl = NL ? NL : nlast;
#if SB_SCAN_ACC == 3
  unsigned long long sx = 0, sy = 0;
#endif
  static_for(std::make_integer_sequence<int, 8 * W>{}, [&](auto ic) {
    constexpr int i = decltype(ic)::value;
    if (i >= 8 * (W - 1) + nl) return;
    uint32_t vx, vy;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(vx), "=r"(vy) : "r"(ad[i]), "n"(QD * W * 128 * 8 + i * 128));
#if SB_SCAN_ACC == 3
    // One mask per word instead of mask + permute: the even bytes are accumulated as two u16 lanes (IMAD),
    // the WHOLE word as a 64-bit integer (IMAD.WIDE); sum(word) = E_lo + 2^8 O_lo + 2^16 E_hi + 2^24 O_hi, so
    // the odd lanes are (S - E) >> 8 at the end.
    const uint32_t xe = vx & 0x00FF00FFu, ye = vy & 0x00FF00FFu;
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e0) : "r"(xe), "r"(one));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e1) : "r"(ye), "r"(one));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(sx) : "r"(vx), "r"(one));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(sy) : "r"(vy), "r"(one));
#else
    const uint32_t xe = vx & 0x00FF00FFu, xo = __byte_perm(vx, 0u, 0x4341);
    const uint32_t ye = vy & 0x00FF00FFu, yo = __byte_perm(vy, 0u, 0x4341);
#if SB_SCAN_ACC == 0
    (void)one;
    e0 += xe; o0 += xo; e1 += ye; o1 += yo;
#elif SB_SCAN_ACC == 1
    e0 += xe; o0 += xo;
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e1) : "r"(ye), "r"(one));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(o1) : "r"(yo), "r"(one));
#else
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e0) : "r"(xe), "r"(one));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(o0) : "r"(xo), "r"(one));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e1) : "r"(ye), "r"(one));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(o1) : "r"(yo), "r"(one));
#endif
#endif
  });
#if SB_SCAN_ACC == 3
  acc[0] = e0; acc[1] = (uint32_t)((sx - e0) >> 8); acc[2] = e1; acc[3] = (uint32_t)((sy - e1) >> 8);
#else
  acc[0] = e0; acc[1] = o0; acc[2] = e1; acc[3] = o1;  // (s0,s2) (s1,s3) (s4,s6) (s5,s7) as u16 lanes
#endif
}

// Interleave up to eight uint8 LUTs (8W*16 bytes each, NULL = all zero) into an oct table.
__device__ __forceinline__ void build_oct_table(uint2* __restrict__ tbl, const uint8_t* const (&l)[8],
                                                int n_entries, int tid, int nthreads) {
  for (int t = tid; t < n_entries / 4; t += nthreads) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = l[i] ? reinterpret_cast<const uint32_t*>(l[i])[t] : 0u;
    uint32_t x[4], y[4];
    {
      const uint32_t ab01 = __byte_perm(v[0], v[1], 0x5140), cd01 = __byte_perm(v[2], v[3], 0x5140);
      const uint32_t ab23 = __byte_perm(v[0], v[1], 0x7362), cd23 = __byte_perm(v[2], v[3], 0x7362);
      x[0] = __byte_perm(ab01, cd01, 0x5410); x[1] = __byte_perm(ab01, cd01, 0x7632);
      x[2] = __byte_perm(ab23, cd23, 0x5410); x[3] = __byte_perm(ab23, cd23, 0x7632);
    }
    {
      const uint32_t ab01 = __byte_perm(v[4], v[5], 0x5140), cd01 = __byte_perm(v[6], v[7], 0x5140);
      const uint32_t ab23 = __byte_perm(v[4], v[5], 0x7362), cd23 = __byte_perm(v[6], v[7], 0x7362);
      y[0] = __byte_perm(ab01, cd01, 0x5410); y[1] = __byte_perm(ab01, cd01, 0x7632);
      y[2] = __byte_perm(ab23, cd23, 0x5410); y[3] = __byte_perm(ab23, cd23, 0x7632);
    }
    reinterpret_cast<uint4*>(tbl)[2 * t] = make_uint4(x[0], y[0], x[1], y[1]);
    reinterpret_cast<uint4*>(tbl)[2 * t + 1] = make_uint4(x[2], y[2], x[3], y[3]);
  }
}

// ---- wide quads (sparse batches) -----------------------------------------------------------
// Where a leaf is probed by only a few queries of the batch (C5 shape: 10k queries over 40k leaves, ~6 per leaf and
// fewer per scan phase) most lanes of an oct are empty and its mask / permute / accumulate work is wasted.  A WIDE QUAD
// serves four queries from a table of 64-bit entries that already hold u16 lanes,
//   T[b][c] = { lut0 | lut1 << 16,  lut2 | lut3 << 16 },
// so a lookup is one LDS.64 and two adds (3 issue slots per four lookups instead of 9 per eight) and the unit of wasted
// work is four queries instead of eight.  It moves twice the shared-memory bytes per lookup (64 lookups / clk / SM at
// the LDS limit instead of 128), so dense batches keep the octs; launch_scan picks per launch.
template <int W, int NL, int QD>
__device__ __forceinline__ void score_wquad_addr(const uint32_t (&ad)[8 * W], int nlast, uint32_t one,
                                                 uint32_t (&acc)[2]) {
  uint32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  const int nl = NL ? NL : nlast;
  static_for(std::make_integer_sequence<int, 8 * W>{}, [&](auto ic) {
    constexpr int i = decltype(ic)::value;
    if (i >= 8 * (W - 1) + nl) return;
    uint32_t vx, vy;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(vx), "=r"(vy) : "r"(ad[i]), "n"(QD * W * 128 * 8 + i * 128));
#if SB_SCAN_WACC == 0
    (void)one;
    if (i & 1) { x1 += vx; y1 += vy; } else { x0 += vx; y0 += vy; }
#elif SB_SCAN_WACC == 1
    if (i & 1) { x1 += vx; asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y1) : "r"(vy), "r"(one)); }
    else { x0 += vx; asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y0) : "r"(vy), "r"(one)); }
#else
    if (i & 1) {
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(vx), "r"(one));
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y1) : "r"(vy), "r"(one));
    } else {
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(vx), "r"(one));
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y0) : "r"(vy), "r"(one));
    }
#endif
  });
  acc[0] = x0 + x1;  // s0 | s1 << 16
  acc[1] = y0 + y1;  // s2 | s3 << 16
}

// Widen up to four uint8 LUTs (8W*16 bytes each, NULL = all zero) into a wide-quad table.
__device__ __forceinline__ void build_wquad_table(uint2* __restrict__ tbl, const uint8_t* const (&l)[4],
                                                  int n_entries, int tid, int nthreads) {
  for (int t = tid; t < n_entries / 4; t += nthreads) {
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = l[i] ? reinterpret_cast<const uint32_t*>(l[i])[t] : 0u;
    // entry e: x = byte e of v0 | byte e of v1 << 16, y likewise for v2, v3
    uint2 o[4];
    o[0] = make_uint2(__byte_perm(v[0], v[1], 0x4440) & 0x00FF00FFu, __byte_perm(v[2], v[3], 0x4440) & 0x00FF00FFu);
    o[1] = make_uint2(__byte_perm(v[0], v[1], 0x5551) & 0x00FF00FFu, __byte_perm(v[2], v[3], 0x5551) & 0x00FF00FFu);
    o[2] = make_uint2(__byte_perm(v[0], v[1], 0x6662) & 0x00FF00FFu, __byte_perm(v[2], v[3], 0x6662) & 0x00FF00FFu);
    o[3] = make_uint2(__byte_perm(v[0], v[1], 0x7773) & 0x00FF00FFu, __byte_perm(v[2], v[3], 0x7773) & 0x00FF00FFu);
    reinterpret_cast<uint4*>(tbl)[2 * t] = make_uint4(o[0].x, o[0].y, o[1].x, o[1].y);
    reinterpret_cast<uint4*>(tbl)[2 * t + 1] = make_uint4(o[2].x, o[2].y, o[3].x, o[3].y);
  }
}

// Interleave up to four uint8 LUTs (8W*16 bytes each, NULL = all zero) into a quad table.
__device__ __forceinline__ void build_quad_table(uint32_t* __restrict__ tbl, const uint8_t* l0,
                                                 const uint8_t* l1, const uint8_t* l2,
                                                 const uint8_t* l3, int n_entries, int tid,
                                                 int nthreads) {
  for (int t = tid; t < n_entries / 4; t += nthreads) {
    const uint32_t a = l0 ? reinterpret_cast<const uint32_t*>(l0)[t] : 0u;
    const uint32_t b = l1 ? reinterpret_cast<const uint32_t*>(l1)[t] : 0u;
    const uint32_t c = l2 ? reinterpret_cast<const uint32_t*>(l2)[t] : 0u;
    const uint32_t d = l3 ? reinterpret_cast<const uint32_t*>(l3)[t] : 0u;
    // 4x4 byte transpose: entry i = a.byte[i] | b.byte[i] << 8 | c.byte[i] << 16 | d.byte[i] << 24
    const uint32_t ab01 = __byte_perm(a, b, 0x5140), cd01 = __byte_perm(c, d, 0x5140);
    const uint32_t ab23 = __byte_perm(a, b, 0x7362), cd23 = __byte_perm(c, d, 0x7362);
    uint4 o;
    o.x = __byte_perm(ab01, cd01, 0x5410);
    o.y = __byte_perm(ab01, cd01, 0x7632);
    o.z = __byte_perm(ab23, cd23, 0x5410);
    o.w = __byte_perm(ab23, cd23, 0x7632);
    reinterpret_cast<uint4*>(tbl)[t] = o;
  }
}

// K-th smallest (1-based) of n UNIQUE u64 keys in shared memory: 8-pass MSB radix select.
// All threads of the block call it; `hist` (256 words) and `sh` (2 words + 1 u64) are shared scratch.
struct SelectScratch { uint32_t hist[256]; unsigned long long prefix; uint32_t need; };
__device__ __forceinline__ uint64_t block_radix_select(const uint64_t* s, uint32_t n, uint32_t K, SelectScratch* sc) {
  const int tid = threadIdx.x, lane = tid & 31;
  uint64_t prefix = 0, mask = 0;
  uint32_t need = K;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) sc->hist[i] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += blockDim.x) {
      const uint64_t k = s[i];
      if ((k & mask) == prefix) atomicAdd(&sc->hist[(uint32_t)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {  // warp 0: 8 bins per lane, warp scan, pick the bin where the count crosses `need`
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { c[k] = sc->hist[lane * 8 + k]; tot += c[k]; }
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t hit = __ballot_sync(kFull, incl >= need);
      const int tl = hit ? (__ffs(hit) - 1) : 31;
      if (lane == tl) {
        uint32_t cum = incl - tot, digit = (uint32_t)lane * 8 + 7;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (cum + c[k] >= need) { digit = (uint32_t)lane * 8 + k; break; }
          cum += c[k];
        }
        sc->prefix = prefix | ((uint64_t)digit << shift);
        sc->need = need - cum;
      }
    }
    __syncthreads();
    prefix = sc->prefix;
    need = sc->need;
    mask |= 0xFFull << shift;
  }
  return prefix;
}

// Upper bound of the K-th smallest (1-based) of n u64 keys in shared memory, n >= K: returns an actual
// key T >= the K-th smallest and, in *n_le, how many keys are <= T.  One histogram pass over a linear,
// monotone 256-bin image of the keys' high words (the float score), a second one inside the pivot's bin
// when that bin is crowded, then a max-reduction: 3-4 block barriers per level instead of the 24 of the
// exact 8-pass select.  A pruning threshold only has to be an upper bound, it need not be tight.
__device__ __forceinline__ uint64_t block_select_bound(const uint64_t* s, uint32_t n, uint32_t K, SelectScratch* sc,
                                                       uint32_t* n_le) {
  const int tid = threadIdx.x, lane = tid & 31;
  __shared__ uint32_t sb_lo[4], sb_hi[4];
  __shared__ unsigned long long sb_max[4];
  uint32_t lo = 0, hi = 0xFFFFFFFFu, need = K, below = 0;  // current high-word range [lo, hi] holding the K-th key
  uint64_t T = 0;
  for (int level = 0; level < 2; ++level) {
    // range of the high words inside [lo, hi]
    uint32_t mn = 0xFFFFFFFFu, mx = 0;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
      const uint32_t o = (uint32_t)(s[i] >> 32);
      if (o >= lo && o <= hi) { mn = min(mn, o); mx = max(mx, o); }
    }
    mn = __reduce_min_sync(kFull, mn);
    mx = __reduce_max_sync(kFull, mx);
    for (int i = tid; i < 256; i += blockDim.x) sc->hist[i] = 0;
    if (lane == 0) { sb_lo[tid >> 5] = mn; sb_hi[tid >> 5] = mx; }
    __syncthreads();
    mn = min(min(sb_lo[0], sb_lo[1]), min(sb_lo[2], sb_lo[3]));
    mx = max(max(sb_hi[0], sb_hi[1]), max(sb_hi[2], sb_hi[3]));
    const float scale = mx > mn ? 255.9f / (float)(mx - mn) : 0.f;
    auto bin = [&](uint32_t o) -> uint32_t { return min((uint32_t)((float)(o - mn) * scale), 255u); };
    for (uint32_t i = tid; i < n; i += blockDim.x) {
      const uint32_t o = (uint32_t)(s[i] >> 32);
      if (o >= lo && o <= hi) atomicAdd(&sc->hist[bin(o)], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { c[k] = sc->hist[lane * 8 + k]; tot += c[k]; }
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t hit = __ballot_sync(kFull, incl >= need);
      const int tl = hit ? (__ffs(hit) - 1) : 31;
      if (lane == tl) {
        uint32_t cum = incl - tot, digit = (uint32_t)lane * 8 + 7, bucket = c[7];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (cum + c[k] >= need) { digit = (uint32_t)lane * 8 + k; bucket = c[k]; break; }
          cum += c[k];
        }
        sc->prefix = ((unsigned long long)digit << 32) | bucket;
        sc->need = cum;
      }
    }
    __syncthreads();
    const uint32_t digit = (uint32_t)(sc->prefix >> 32), bucket = (uint32_t)sc->prefix, cum = sc->need;
    // largest key of the pivot's bin and the new range
    unsigned long long tmax = 0;
    uint32_t nlo = 0xFFFFFFFFu, nhi = 0;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
      const uint64_t k = s[i];
      const uint32_t o = (uint32_t)(k >> 32);
      if (o >= lo && o <= hi && bin(o) == digit) {
        tmax = max(tmax, (unsigned long long)k);
        nlo = min(nlo, o);
        nhi = max(nhi, o);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(kFull, tmax, o);
      tmax = max(tmax, t);
    }
    nlo = __reduce_min_sync(kFull, nlo);
    nhi = __reduce_max_sync(kFull, nhi);
    __syncthreads();  // sb_* are re-used
    if (lane == 0) { sb_max[tid >> 5] = tmax; sb_lo[tid >> 5] = nlo; sb_hi[tid >> 5] = nhi; }
    __syncthreads();
    T = max(max(sb_max[0], sb_max[1]), max(sb_max[2], sb_max[3]));
    nlo = min(min(sb_lo[0], sb_lo[1]), min(sb_lo[2], sb_lo[3]));
    nhi = max(max(sb_hi[0], sb_hi[1]), max(sb_hi[2], sb_hi[3]));
    __syncthreads();
    *n_le = below + cum + bucket;
    if (bucket <= 16 || nlo == nhi) break;  // tight enough, or the bin cannot be split by the high word
    lo = nlo; hi = nhi;
    below += cum;
    need -= cum;
  }
  return T;
}

// ---------------------------------------------------------------------------------------
// Pilot: one CTA per query, nearest leaves first, exact top-N threshold in shared memory.
// ---------------------------------------------------------------------------------------
// W = 0: the generic instantiation for 16 < ix.W <= 32 (128 < B <= 256 blocks, asymmetric_hashing_impl.cc:656-688:
// still the int16 accumulator): the word count is a runtime value and the code words are loaded one at a time.
template <int WT>
__global__ void __launch_bounds__(kScanThreads)
pilot_kernel(DevIndex ix, ScanWork w, int capl) {
  const int W = WT ? WT : (int)ix.W;
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* tbl = reinterpret_cast<uint32_t*>(smem);                  // [W*128]
  uint64_t* scand = reinterpret_cast<uint64_t*>(smem + W * 128 * 4);  // [capl]
  __shared__ uint64_t s_tau;
  __shared__ int s_thr;
  __shared__ uint32_t s_cnt;
  __shared__ SelectScratch s_sel;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t q = blockIdx.x;
  const int nlast = (int)ix.B - 8 * (W - 1);
  const int off128 = 128 * (int)ix.B;
  const uint32_t nover = w.nover;
  uint64_t* grow = w.buf + (size_t)q * w.cap;  // this query's (still empty) buffer row is the compaction scratch
  float mult, inv;
  if (w.q_for_lut) {
    // Fused LUT build (lut_kernel's arithmetic, lut.cuh): raw table in the still unused candidate buffer, block
    // maximum, multiplier, then the u8 table goes to shared memory for this kernel and to global memory for the
    // main scan -- one launch and one re-read of the table less per batch.
    __shared__ float s_red[kScanThreads / 32];
    __shared__ float s_mi[2];
    float* sq = reinterpret_cast<float*>(scand);
    float* raw = sq + ((ix.d + 3) & ~3u);
    for (uint32_t k = tid; k < ix.d; k += kScanThreads) sq[k] = w.q_for_lut[(size_t)q * ix.d + k];
    __syncthreads();
    const uint32_t ne = ix.B * 16;
    float mx = 0.f;
    for (uint32_t e = tid; e < ne; e += kScanThreads) {
      const float r = lut_raw_entry(ix, sq, e);
      raw[e] = r;
      mx = fmaxf(mx, fabsf(r));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
      float m = 0.f;
      for (int i = 0; i < kScanThreads / 32; ++i) m = fmaxf(m, s_red[i]);
      const float mu = lut_multiplier(m);
      const float iv = lut_inverse_multiplier(ix, mu);
      s_mi[0] = mu; s_mi[1] = iv;
      const_cast<float*>(w.mult)[q] = mu;
      const_cast<float*>(w.inv_mult)[q] = iv;
    }
    __syncthreads();
    mult = s_mi[0]; inv = s_mi[1];
    uint32_t* gl = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(w.lut) + (size_t)q * W * 128);
    for (int t = tid; t < W * 32; t += kScanThreads) {
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t e = 4 * t + j;
        if (e < ne) word |= lut_quantize(raw[e], mult) << (8 * j);
      }
      tbl[t] = word;
      gl[t] = word;
    }
    __syncthreads();  // raw / sq alias the candidate buffer
  } else {  // the query's own u8 LUT, [8W][16] bytes
    const uint32_t* src = reinterpret_cast<const uint32_t*>(w.lut + (size_t)q * W * 128);
    for (int t = tid; t < W * 32; t += kScanThreads) tbl[t] = src[t];
    mult = w.mult[q]; inv = w.inv_mult[q];
  }
  if (tid == 0) { s_tau = kKeyMax; s_cnt = 0; }
  // Sample whole leaves, nearest first, until at least 4 N slots have been scored: the N-th best of
  // that sample is the pruning threshold of the main scan.  Nothing is published -- the main scan
  // covers every probed leaf, these too (1-2 % more scan work, and no hand-over of candidates).
  // Leaf-sharded search: only the owner of the query's nearest leaf samples (its own leaves, nearest first).
  uint32_t seen = 0;
  const int leaf0 = w.leaves[(size_t)q * w.P];
  const bool sampler = w.pilot_world <= 1 || (leaf0 >= 0 && (uint32_t)leaf0 % w.pilot_world == w.pilot_rank);
  for (uint32_t r = 0; sampler && r < w.P; ++r) {
    const int leaf = w.leaves[(size_t)q * w.P + r];
    if (leaf < 0) break;
    const float bias = ix.key_by_dp ? 0.f : w.bias[(size_t)q * w.P + r];
    const uint32_t n = ix.leaf_size[leaf];
    const uint32_t gbeg = ix.leaf_goff[leaf], ng = ix.leaf_goff[leaf + 1] - gbeg;
    __syncthreads();
    if (tid == 0) s_thr = acc_threshold(s_tau, mult, inv, bias) + off128;
    __syncthreads();
    for (uint32_t g0 = 0; g0 < ng; g0 += kScanWarps) {
      if (w.pilot_partial && seen + g0 * 32 >= w.pilot_target) break;  // enough slots sampled (slot order = id order)
      const uint32_t g = g0 + warp;
      if (g < ng) {
        int s0;
        if constexpr (WT != 0) {
          uint32_t cw[WT ? WT : 1];
          load_codes<(WT ? WT : 1)>(ix.codes + (size_t)(gbeg + g) * W * 32, lane, cw);
          s0 = score_one<(WT ? WT : 1)>(cw, reinterpret_cast<const unsigned char*>(tbl), nlast);
        } else {
          s0 = score_one_rt(ix.codes + (size_t)(gbeg + g) * W * 32, lane, W, reinterpret_cast<const unsigned char*>(tbl), nlast);
        }
        bool p = (g * 32 + lane < n) && s0 <= s_thr;
        uint64_t key = 0;
        if (p) {
          const uint32_t gslot = (gbeg + g) * 32 + lane;
          key = make_key(ah_float_score(s0 - off128, inv, bias), ix.key_by_dp ? ix.slot_dp[gslot] : gslot);
          p = key < s_tau;
        }
        const uint32_t m = __ballot_sync(kFull, p);
        if (m) {
          const int leader = __ffs(m) - 1;
          uint32_t base = 0;
          if (lane == leader) base = atomicAdd(&s_cnt, (uint32_t)__popc(m));
          base = __shfl_sync(kFull, base, leader);
          if (p) scand[base + __popc(m & ((1u << lane) - 1u))] = key;
        }
      }
      __syncthreads();
      const uint32_t c = s_cnt;
      __syncthreads();  // everyone has read s_cnt before the next round may bump it
      if (c > (uint32_t)(capl - kScanThreads)) {
        // buffer full: keep the keys up to a bound T of the N-th smallest (compacted through the global row)
        uint32_t n_le = 0;
        uint64_t T = block_select_bound(scand, c, nover, &s_sel, &n_le);
        if (n_le > (uint32_t)capl / 2) T = block_radix_select(scand, c, nover, &s_sel);  // crowded bin: exact select
        const uint32_t nkeep = n_le > (uint32_t)capl / 2 ? nover : n_le;
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        for (uint32_t i = tid; i < c; i += kScanThreads) {
          const uint64_t k = scand[i];
          if (k <= T) grow[atomicAdd(&s_cnt, 1u)] = k;
        }
        __syncthreads();
        for (uint32_t i = tid; i < nkeep; i += kScanThreads) scand[i] = grow[i];
        if (tid == 0) {
          s_tau = T;
          s_thr = acc_threshold(T, mult, inv, bias) + off128;
        }
        __syncthreads();
      }
    }
    seen += n;
    if (seen >= w.pilot_target) break;
  }
  __syncthreads();
  // tau: an upper bound of the N-th smallest sampled key, plus one so that the bounding element itself
  // passes the main scan's strict `key < tau` test.
  const uint32_t c = s_cnt;
  uint64_t tau = kKeyMax;
  if (c >= nover) {
    uint32_t n_le = 0;
    tau = block_select_bound(scand, c, nover, &s_sel, &n_le);
    if (n_le > 2 * nover + 64) tau = block_radix_select(scand, c, nover, &s_sel);  // crowded bin: exact N-th key
    if (tau != kKeyMax) tau += 1;
  }
  if (tid == 0) {
    w.cnt[q] = 0;
    w.tau[q] = tau;
    w.pilot_end[q] = 0;
    w.ovf[q] = 0;
  }
  // Work-list counts of the leaves left to the main scan, and the traffic statistics of all probed
  // leaves (worklist_count_kernel's job, done here while the query's leaf list is hot).
  unsigned long long bytes = 0, pairs = 0;
  for (uint32_t rr = tid; rr < w.P; rr += kScanThreads) {
    const int leaf = w.leaves[(size_t)q * w.P + rr];
    if (leaf >= 0) {
      bytes += (unsigned long long)(ix.leaf_goff[leaf + 1] - ix.leaf_goff[leaf]) * 16ull * ix.B;
      pairs += 1;
      // the count IS the pair's position in its leaf's entry list: the scatter pass needs no second atomic
      if (rr >= w.rank_lo && rr < w.rank_hi) w.pair_pos[(size_t)q * w.P + rr] = atomicAdd(&w.leaf_cnt[leaf], 1u);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bytes += __shfl_xor_sync(kFull, bytes, o);
    pairs += __shfl_xor_sync(kFull, pairs, o);
  }
  if (lane == 0 && pairs) {
    atomicAdd(&w.stats[0], bytes);
    atomicAdd(&w.stats[1], pairs);
  }
}

// ---------------------------------------------------------------------------------------
// Work list: (query, rank >= pilot_end) pairs bucketed by leaf, then cut into items of
// (leaf tile, <= quads_per_item quads).
// ---------------------------------------------------------------------------------------
__global__ void worklist_count_kernel(DevIndex ix, ScanWork w, int only_ovf, int count_stats) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)w.nq * w.P;
  unsigned long long bytes = 0, pairs = 0;
  if (i < total) {
    const uint32_t q = (uint32_t)(i / w.P), r = (uint32_t)(i % w.P);
    const int leaf = w.leaves[i];
    if (leaf >= 0) {
      if (count_stats) {
        const uint32_t ng = ix.leaf_goff[leaf + 1] - ix.leaf_goff[leaf];
        bytes = (unsigned long long)ng * 16ull * ix.B;
        pairs = 1;
      }
      if (r >= w.rank_lo && r < w.rank_hi && (!only_ovf || w.ovf[q])) w.pair_pos[i] = atomicAdd(&w.leaf_cnt[leaf], 1u);
    }
  }
  if (count_stats) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bytes += __shfl_xor_sync(kFull, bytes, o);
      pairs += __shfl_xor_sync(kFull, pairs, o);
    }
    if ((threadIdx.x & 31) == 0 && pairs) {
      atomicAdd(&w.stats[0], bytes);
      atomicAdd(&w.stats[1], pairs);
    }
  }
}

__global__ void __launch_bounds__(1024) worklist_scan_kernel(DevIndex ix, ScanWork w) {
  __shared__ uint32_t wsum_e[32], wsum_i[32];
  __shared__ uint32_t carry_e, carry_i;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_e = 0; carry_i = 0; }
  __syncthreads();
  for (uint32_t base = 0; base < ix.L; base += 1024) {
    const uint32_t l = base + tid;
    uint32_t ce = 0, ci = 0;
    if (l < ix.L) {
      ce = w.leaf_cnt[l];
      const uint32_t qpi = w.quads_per_item * kQueriesPerQuad;
      uint32_t nt, gp;
      leaf_tiling(ix.leaf_goff[l + 1] - ix.leaf_goff[l], w.max_gpt, &nt, &gp);
      ci = ((ce + qpi - 1) / qpi) * nt;
    }
    uint32_t se = ce, si = ci;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t te = __shfl_up_sync(kFull, se, o), ti = __shfl_up_sync(kFull, si, o);
      if (lane >= o) { se += te; si += ti; }
    }
    if (lane == 31) { wsum_e[warp] = se; wsum_i[warp] = si; }
    __syncthreads();
    if (warp == 0) {
      uint32_t ve = wsum_e[lane], vi = wsum_i[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t te = __shfl_up_sync(kFull, ve, o), ti = __shfl_up_sync(kFull, vi, o);
        if (lane >= o) { ve += te; vi += ti; }
      }
      wsum_e[lane] = ve; wsum_i[lane] = vi;
    }
    __syncthreads();
    const uint32_t pe = carry_e + (warp ? wsum_e[warp - 1] : 0u) + se - ce;
    const uint32_t pi = carry_i + (warp ? wsum_i[warp - 1] : 0u) + si - ci;
    if (l < ix.L) {
      w.leaf_eoff[l] = pe; w.item_off[l] = pi; w.leaf_cur[l] = 0;
      // item -> leaf table (the scan kernel's alternative to a binary search over item_off per item)
      if (w.item_leaf)
        for (uint32_t j = 0; j < ci && pi + j < w.item_leaf_cap; ++j) w.item_leaf[pi + j] = l;
    }
    __syncthreads();
    if (tid == 0) { carry_e += wsum_e[31]; carry_i += wsum_i[31]; }
    __syncthreads();
  }
  if (tid == 0) {
    w.leaf_eoff[ix.L] = carry_e;
    w.item_off[ix.L] = carry_i;
    w.counters[0] = 0;
    w.counters[1] = carry_i;
    w.counters[3] = carry_e;
  }
}

__global__ void worklist_scatter_kernel(DevIndex ix, ScanWork w, int only_ovf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)w.nq * w.P;
  if (i >= total) return;
  const uint32_t q = (uint32_t)(i / w.P), r = (uint32_t)(i % w.P);
  const int leaf = w.leaves[i];
  if (leaf < 0) return;
  if (r < w.rank_lo || r >= w.rank_hi || (only_ovf && !w.ovf[q])) return;
  const uint32_t pos = w.leaf_eoff[leaf] + w.pair_pos[i];
  w.entry_q[pos] = q;
  w.entry_bias[pos] = w.bias[i];
}

// The first work list of a batch is counted by the pilot kernel (launch_pilot zeroes leaf_cnt); later
// ones (second scan phase, re-scans of overflowed queries) are counted here.
void launch_worklist(const DevIndex& ix, const ScanWork& w, bool only_overflowed, bool counted, cudaStream_t s,
                     int* launches) {
  const size_t total = (size_t)w.nq * w.P;
  const int blocks = (int)((total + 255) / 256);
  if (!counted) {
    cudaMemsetAsync(w.leaf_cnt, 0, sizeof(uint32_t) * (ix.L + 1), s);
    worklist_count_kernel<<<blocks, 256, 0, s>>>(ix, w, only_overflowed ? 1 : 0, 0);
    if (launches) *launches += 1;
  }
  worklist_scan_kernel<<<1, 1024, 0, s>>>(ix, w);
  worklist_scatter_kernel<<<blocks, 256, 0, s>>>(ix, w, only_overflowed ? 1 : 0);
  if (launches) *launches += 2;
}

// ---------------------------------------------------------------------------------------
// Main scan: persistent CTAs pull (leaf tile, query chunk) items from an atomic counter.
// ---------------------------------------------------------------------------------------
struct ItemMeta {
  uint64_t tau[kMaxQPI];
  uint32_t q[kMaxQPI];
  int thr[kMaxQPI];
  float inv[kMaxQPI], bias[kMaxQPI];
  // thresholds of an oct packed like its accumulators, (t0,t2) (t1,t3) (t4,t6) (t5,t7) as u16 halves, each
  // biased by 32768: bit 15 of a half of (thrp - acc) is set iff sum <= thr (sums < 2^15 because B <= 128)
  uint4 thrp[kMaxQPI / 8];
  // the same for a wide quad: (t0,t1) (t2,t3)
  uint2 thrpw[kMaxQPI / 4];
  // per-item staging of the candidates of each query: the scan appends here with shared-memory atomics and the item's
  // epilogue publishes a query's keys with ONE global atomicAdd (one lane per staged key).  Where a few slots of almost
  // every 32-slot group pass the threshold (C5-size leaves: ~20 candidates per (query, item)) the global atomic and its
  // round trip leave the scoring loop; a full stage spills to the direct path below.
  uint32_t scnt[kMaxQPI];
  uint64_t sbuf[kMaxQPI * 32];  // [queries per item][stage entries]: 32 per query at 32 queries per item, 64 at 16
};
constexpr uint32_t kStageKeys = kMaxQPI * 32;

// Push path of the main scan: exact key test and warp-aggregated append for one oct.  All 32 lanes call.  `hit` holds
// the lane's packed threshold test (bit 15 / 31 of word j set iff the sum in that half passed); one warp-wide OR finds
// the queries of the oct with any passing slot and only those are visited.
// WIDE = 1: a wide quad -- a0 = s0 | s1 << 16, a1 = s2 | s3 << 16 (a2, a3, h2, h3 unused).
template <int WIDE>
__device__ __noinline__ void push_candidates(const DevIndex& ix, const ScanWork& w, ItemMeta* meta, uint32_t qd,
                                             uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t h0, uint32_t h1,
                                             uint32_t h2, uint32_t h3, bool valid, uint32_t gslot, int off128) {
  const int lane = threadIdx.x & 31;
  const uint32_t kstage = kStageKeys / (w.quads_per_item * kQueriesPerQuad);
  // query i of the oct: sums (s0,s2) (s1,s3) (s4,s6) (s5,s7) sit in a0..a3 as u16 halves
  uint32_t m8 = WIDE ? (((h0 >> 15) & 1u) | ((h0 >> 30) & 2u) | ((h1 >> 13) & 4u) | ((h1 >> 28) & 8u))
                     : (((h0 >> 15) & 1u) | ((h1 >> 14) & 2u) | ((h0 >> 29) & 4u) | ((h1 >> 28) & 8u) |
                        ((h2 >> 11) & 16u) | ((h3 >> 10) & 32u) | ((h2 >> 25) & 64u) | ((h3 >> 24) & 128u));
  if (!valid) m8 = 0u;
  uint32_t any8 = __reduce_or_sync(kFull, m8);
  while (any8) {
    const int i = __ffs(any8) - 1;
    any8 &= any8 - 1;
    const int qi = qd * (WIDE ? 4 : 8) + i;
    bool p = (m8 >> i) & 1u;
    uint64_t key = 0;
    if (p) {
      const uint32_t word = WIDE ? ((i & 2) ? a1 : a0) : ((i & 4) ? ((i & 1) ? a3 : a2) : ((i & 1) ? a1 : a0));
      const int sv = (int)(((WIDE ? (i & 1) : (i & 2)) != 0) ? (word >> 16) : (word & 0xFFFFu));
      key = make_key(ah_float_score(sv - off128, meta->inv[qi], meta->bias[qi]),
                     ix.key_by_dp ? ix.slot_dp[gslot] : gslot);
      p = key < meta->tau[qi];
    }
    uint32_t m = __ballot_sync(kFull, p);
    if (m) {
      // stage in shared memory; the part of the group that does not fit goes to the query's buffer directly
      int leader = __ffs(m) - 1;
      uint32_t sbase = 0;
      if (w.stage) {
        if (lane == leader) sbase = atomicAdd(&meta->scnt[qi], (uint32_t)__popc(m));
        sbase = __shfl_sync(kFull, sbase, leader);
        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
        if (p && sbase + rank < kstage) { meta->sbuf[qi * kstage + sbase + rank] = key; p = false; }
        m = __ballot_sync(kFull, p);
      }
      if (m) {
        const uint32_t qq = meta->q[qi];
        leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&w.cnt[qq], (uint32_t)__popc(m));
        base = __shfl_sync(kFull, base, leader);
        if (p) {
          const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
          if (pos < w.cap) w.buf[(size_t)qq * w.cap + pos] = key;
          else w.ovf[qq] = 1u;
        }
      }
    }
  }
}

// Push path of a wide quad with per-item staging: every lane appends its own passing (query, slot) pairs with a
// shared-memory atomic -- no warp-wide loop over the queries, no ballots: with a handful of passing pairs per warp,
// mostly on different lanes, the divergent loop runs once or twice.  A full stage spills to the query's buffer.
__device__ __forceinline__ void push_lane_staged(const DevIndex& ix, const ScanWork& w, ItemMeta* meta, uint32_t qd,
                                                 uint32_t a0, uint32_t a1, uint32_t h0, uint32_t h1, uint32_t gslot,
                                                 int off128, uint32_t kstage) {
  uint32_t m = ((h0 >> 15) & 1u) | ((h0 >> 30) & 2u) | ((h1 >> 13) & 4u) | ((h1 >> 28) & 8u);
  const uint32_t tie = ix.key_by_dp ? ix.slot_dp[gslot] : gslot;
  while (m) {
    const int i = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t qi = qd * 4 + i;
    const uint32_t word = (i & 2) ? a1 : a0;
    const int sv = (int)((i & 1) ? (word >> 16) : (word & 0xFFFFu));
    const uint64_t key = make_key(ah_float_score(sv - off128, meta->inv[qi], meta->bias[qi]), tie);
    if (key < meta->tau[qi]) {
      const uint32_t pos = atomicAdd(&meta->scnt[qi], 1u);
      if (pos < kstage) {
        meta->sbuf[qi * kstage + pos] = key;
      } else {
        const uint32_t qq = meta->q[qi];
        const uint32_t gpos = atomicAdd(&w.cnt[qq], 1u);
        if (gpos < w.cap) w.buf[(size_t)qq * w.cap + gpos] = key;
        else w.ovf[qq] = 1u;
      }
    }
  }
}

template <int W, int NL, int WIDE>
__global__ void __launch_bounds__(kScanThreads, SB_SCAN_MIN_CTAS)
scan_main_kernel(DevIndex ix, ScanWork w) {
  extern __shared__ __align__(16) unsigned char smem[];
  // [octs per item][W*128] 64-bit entries, 128-byte aligned: the lookup addresses are formed with OR
  uint2* tables = reinterpret_cast<uint2*>((reinterpret_cast<uintptr_t>(smem) + 127) & ~(uintptr_t)127);
  __shared__ ItemMeta meta;
  uint32_t (&s_q)[kMaxQPI] = meta.q;
  int (&s_thr)[kMaxQPI] = meta.thr;
  uint64_t (&s_tau)[kMaxQPI] = meta.tau;
  float (&s_inv)[kMaxQPI] = meta.inv;
  float (&s_bias)[kMaxQPI] = meta.bias;
  __shared__ uint32_t s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Wr = W ? W : (int)ix.W;  // W = 0: the generic instantiation (16 < ix.W <= 32), wide quads only
  const int nlast = (int)ix.B - 8 * (Wr - 1);
  const int off128 = 128 * (int)ix.B;
  const uint32_t n_items = w.counters[1];
  const uint32_t qpi = w.quads_per_item * kQueriesPerQuad;
  const int kTblEntries = Wr * 128;
  // Items are claimed one ahead: the atomic for item i + 1 is issued when item i starts and its result is only needed
  // when item i is done, so its round trip (and the item -> leaf lookup) overlaps the scoring.
  if (tid == 0) s_item = atomicAdd(&w.counters[0], 1u);
  for (;;) {
    __syncthreads();  // previous item's tables are no longer read; s_item is visible
    const uint32_t item = s_item;
    if (item >= n_items) break;
    uint32_t next_item = 0;
    if (tid == 0) next_item = atomicAdd(&w.counters[0], 1u);
    uint32_t leaf;
    if (w.item_leaf && item < w.item_leaf_cap) {
      leaf = w.item_leaf[item];
    } else {  // leaf = upper_bound(item_off, item) - 1
      uint32_t lo = 0, hi = ix.L;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (w.item_off[mid + 1] <= item) lo = mid + 1; else hi = mid;
      }
      leaf = lo;
    }
    const uint32_t local = item - w.item_off[leaf];
    const uint32_t gbeg_t = ix.leaf_goff[leaf], ng_t = ix.leaf_goff[leaf + 1] - gbeg_t;
    uint32_t ntiles, gpt;
    leaf_tiling(ng_t, w.max_gpt, &ntiles, &gpt);
    const uint32_t chunk = local / ntiles, tile = local - chunk * ntiles;
    const uint32_t gbeg = ix.leaf_goff[leaf], ng = ix.leaf_goff[leaf + 1] - gbeg;
    const uint32_t g0 = tile * gpt, g1 = min(g0 + gpt, ng);
    const uint32_t nleaf = ix.leaf_size[leaf];
    const uint32_t ebase = w.leaf_eoff[leaf] + chunk * qpi;
    const uint32_t ecount = min(qpi, w.leaf_eoff[leaf + 1] - ebase);
    constexpr uint32_t kQPT = WIDE ? 4u : (uint32_t)kQueriesPerQuad;  // queries per table
    const uint32_t nquads = (ecount + kQPT - 1) / kQPT;  // octs (wide quads) in this item
    if (tid < (int)qpi) {
      uint32_t qq = kInvalidQuery;
      int thr = -1;
      if ((uint32_t)tid < ecount) {
        qq = w.entry_q[ebase + tid];
        const float bias = ix.key_by_dp ? 0.f : w.entry_bias[ebase + tid];
        const uint64_t tau = w.tau[qq];
        const float inv = w.inv_mult[qq];
        thr = acc_threshold(tau, w.mult[qq], inv, bias) + off128;
        if (thr < -1) thr = -1;
        s_tau[tid] = tau; s_inv[tid] = inv; s_bias[tid] = bias;
      }
      s_q[tid] = qq;
      s_thr[tid] = thr;
      meta.scnt[tid] = 0;
    }
    __syncthreads();
    if (tid < (int)(qpi / kQPT)) {
      auto bt = [&](int i) -> uint32_t { return (uint32_t)(min(max(s_thr[tid * kQPT + i], -1), 32767) + 32768); };
      if constexpr (WIDE) meta.thrpw[tid] = make_uint2(bt(0) | (bt(1) << 16), bt(2) | (bt(3) << 16));
      else meta.thrp[tid] = make_uint4(bt(0) | (bt(2) << 16), bt(1) | (bt(3) << 16), bt(4) | (bt(6) << 16), bt(5) | (bt(7) << 16));
    }
    for (uint32_t qd = 0; qd < nquads; ++qd) {
      const uint8_t* lp[kQPT];
#pragma unroll
      for (int i = 0; i < (int)kQPT; ++i) {
        const uint32_t qq = s_q[qd * kQPT + i];
        lp[i] = (qq == kInvalidQuery) ? nullptr : w.lut + (size_t)qq * kTblEntries;
      }
      if constexpr (WIDE) build_wquad_table(tables + (size_t)qd * kTblEntries, lp, kTblEntries, tid, kScanThreads);
      else build_oct_table(tables + (size_t)qd * kTblEntries, lp, kTblEntries, tid, kScanThreads);
    }
    __syncthreads();
    if constexpr (W == 0) {
      // Generic path (128 < B <= 256): one code word at a time, the nibbles of a word against up to four wide quads;
      // the u16 lanes hold sums up to 256 * 255 = 65280, so the threshold test unpacks them (the packed test of the
      // specialised kernels needs sums below 2^15).
      static_assert(WIDE == 1, "the generic scan kernel uses wide quads");
      const uint32_t tb32 = (uint32_t)__cvta_generic_to_shared(tables);
      const uint32_t kstage = kStageKeys / qpi;
      for (uint32_t g = g0 + warp; g < g1; g += kScanWarps) {
        const bool valid = g * 32 + lane < nleaf;
        const uint32_t gslot = (gbeg + g) * 32 + lane;
        const uint32_t* gbase = ix.codes + (size_t)(gbeg + g) * Wr * 32;
        uint32_t ax[4] = {0u, 0u, 0u, 0u}, ay[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int j = 0; j < Wr; ++j) {
          const uint32_t word = load_code_word(gbase, lane, Wr, j);
          const int nk = j == Wr - 1 ? nlast : 8;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < nk) {
              const uint32_t a = tb32 + (uint32_t)((8 * j + k) * 128) + (((word >> (4 * k)) & 0xFu) << 3);
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                if ((uint32_t)qd < nquads) {
                  uint32_t vx, vy;
                  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vx), "=r"(vy) : "r"(a + (uint32_t)qd * (uint32_t)kTblEntries * 8u));
                  ax[qd] += vx; ay[qd] += vy;
                }
              }
            }
          }
        }
        if (valid) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            if ((uint32_t)qd >= nquads) break;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t wv = (i & 2) ? ay[qd] : ax[qd];
              const int sv = (int)((i & 1) ? (wv >> 16) : (wv & 0xFFFFu));
              const uint32_t qi = (uint32_t)qd * 4u + (uint32_t)i;
              if (sv <= s_thr[qi]) {
                const uint64_t key = make_key(ah_float_score(sv - off128, s_inv[qi], s_bias[qi]), ix.key_by_dp ? ix.slot_dp[gslot] : gslot);
                if (key < s_tau[qi]) {
                  uint32_t pos = w.stage ? atomicAdd(&meta.scnt[qi], 1u) : kstage;
                  if (pos < kstage) {
                    meta.sbuf[qi * kstage + pos] = key;
                  } else {
                    const uint32_t qq = s_q[qi];
                    const uint32_t gpos = atomicAdd(&w.cnt[qq], 1u);
                    if (gpos < w.cap) w.buf[(size_t)qq * w.cap + gpos] = key;
                    else w.ovf[qq] = 1u;
                  }
                }
              }
            }
          }
        }
      }
    } else {
    // The code words of the warp's next group are requested as soon as the current group's lookup addresses have been
    // formed from them (the registers are free again), so their DRAM latency overlaps the scoring of the current group.
    uint32_t cw[W ? W : 1];
    if (g0 + warp < g1) load_codes<(W ? W : 1)>(ix.codes + (size_t)(gbeg + g0 + warp) * W * 32, lane, cw);
    for (uint32_t g = g0 + warp; g < g1; g += kScanWarps) {
      const bool valid = g * 32 + lane < nleaf;
      const uint32_t gslot = (gbeg + g) * 32 + lane;
      uint32_t ad[8 * W];
      {
        const uint32_t tb32 = (uint32_t)__cvta_generic_to_shared(tables);
#pragma unroll
        for (int j = 0; j < W; ++j)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            ad[8 * j + k] = tb32 | ((k == 0) ? ((cw[j] << 3) & 0x78u) : ((cw[j] >> (4 * k - 3)) & 0x78u));
      }
      if (g + kScanWarps < g1) load_codes<W>(ix.codes + (size_t)(gbeg + g + kScanWarps) * W * 32, lane, cw);
      // fast path inline (8 compares per oct), candidate push out of line: the scoring loops of the
      // octs are unrolled copies and must stay inside the instruction cache
      if constexpr (WIDE) {
        auto filter = [&](const uint32_t qd, const uint32_t (&acc)[2]) {
          const uint2 t = meta.thrpw[qd];
          const uint32_t h0 = t.x - acc[0], h1 = t.y - acc[1];
          const uint32_t hit = (h0 | h1) & 0x80008000u;
          if (w.stage) {
            if (valid && hit != 0) push_lane_staged(ix, w, &meta, qd, acc[0], acc[1], h0, h1, gslot, off128, kStageKeys / qpi);
          } else if (__any_sync(kFull, valid && hit != 0)) {
            push_candidates<1>(ix, w, &meta, qd, acc[0], acc[1], 0u, 0u, h0, h1, 0u, 0u, valid, gslot, off128);
          }
        };
#define SB_DO_WQUAD(QD)                                      \
  if (QD < nquads) {                                         \
    uint32_t acc[2];                                         \
    score_wquad_addr<W, NL, QD>(ad, nlast, w.one, acc);      \
    filter(QD, acc);                                         \
  }
        SB_DO_WQUAD(0) SB_DO_WQUAD(1) SB_DO_WQUAD(2) SB_DO_WQUAD(3)
#undef SB_DO_WQUAD
      } else {
        auto filter = [&](const uint32_t qd, const uint32_t (&acc)[4]) {
          const uint4 t = meta.thrp[qd];
          const uint32_t h0 = t.x - acc[0], h1 = t.y - acc[1], h2 = t.z - acc[2], h3 = t.w - acc[3];
          const uint32_t hit = (h0 | h1 | h2 | h3) & 0x80008000u;
          if (__any_sync(kFull, valid && hit != 0))
            push_candidates<0>(ix, w, &meta, qd, acc[0], acc[1], acc[2], acc[3], h0, h1, h2, h3, valid, gslot, off128);
        };
#define SB_DO_QUAD(QD)                                       \
  if (QD < nquads) {                                         \
    uint32_t acc[4];                                         \
    score_oct_addr<W, NL, QD>(ad, nlast, w.one, acc);        \
    filter(QD, acc);                                         \
  }
        SB_DO_QUAD(0) SB_DO_QUAD(1)
#undef SB_DO_QUAD
      }
    }
    }  // W != 0
    // item epilogue: publish the staged candidates, one warp per query, one global atomicAdd per (query, item)
    if (w.stage) {
      __syncthreads();
      const uint32_t kstage = kStageKeys / qpi;
      for (uint32_t qi = warp; qi < ecount; qi += kScanWarps) {
        const uint32_t c = min(meta.scnt[qi], kstage);
        if (c == 0) continue;
        const uint32_t qq = s_q[qi];
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&w.cnt[qq], c);
        base = __shfl_sync(kFull, base, 0);
        for (uint32_t j = lane; j < c; j += 32) {
          const uint32_t pos = base + j;
          if (pos < w.cap) w.buf[(size_t)qq * w.cap + pos] = meta.sbuf[qi * kstage + j];
          else w.ovf[qq] = 1u;
        }
      }
    }
    // hand the pre-claimed item to the next iteration (every thread read s_item before the first barrier of the body)
    if (tid == 0) s_item = next_item;
  }
}

// ---------------------------------------------------------------------------------------
// Compact: sort one query's buffer, drop duplicate keys (re-scan mode), keep the N smallest,
// publish the new tau.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void compact_one(const ScanWork& w, uint32_t q, int dedup, uint64_t* s,
                                            uint32_t* s_dups) {
  const int tid = threadIdx.x;
  const uint32_t nraw = w.cnt[q];
  const uint32_t n = min(nraw, w.cap);
  const bool over = nraw > w.cap;
  if (n == 0 || (dedup && w.ovf[q] == 0)) {  // re-scan passes only touch flagged queries
    __syncthreads();
    if (tid == 0) w.ovf[q] = 0;
    return;
  }
  int np2 = 2;
  while ((uint32_t)np2 < n) np2 <<= 1;
  for (int i = tid; i < np2; i += kScanThreads) s[i] = (uint32_t)i < n ? w.buf[(size_t)q * w.cap + i] : kKeyMax;
  if (tid == 0) *s_dups = 0;
  __syncthreads();
  uint32_t nuniq = n;
  if (!dedup && n > max(2u * w.nover, 128u)) {
    // Heavy tail: do not sort thousands of keys to keep N of them.  A bound select finds a key T >= the
    // N-th smallest with only a few more than N keys <= T (exact 8-pass radix select if its bin is
    // crowded); those keys are compacted through the query's own buffer row and only they are sorted.
    __shared__ SelectScratch sc;
    __shared__ uint32_t sel_cnt;
    uint32_t n_le = 0;
    uint64_t prefix = block_select_bound(s, n, w.nover, &sc, &n_le);
    if (n_le > 1024u) prefix = block_radix_select(s, n, w.nover, &sc);  // crowded bin: exact N-th key
    if (tid == 0) sel_cnt = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kScanThreads) {
      const uint64_t k = s[i];
      if (k <= prefix) w.buf[(size_t)q * w.cap + atomicAdd(&sel_cnt, 1u)] = k;
    }
    __syncthreads();
    nuniq = sel_cnt;  // nover (exact select) or a few more (bound select)
    np2 = 2;
    while ((uint32_t)np2 < nuniq) np2 <<= 1;
    for (int i = tid; i < np2; i += kScanThreads) s[i] = (uint32_t)i < nuniq ? w.buf[(size_t)q * w.cap + i] : kKeyMax;
    __syncthreads();
  }
  block_bitonic_sort(s, np2);
  if (dedup) {
    // keys are unique per (leaf, slot); equal neighbours are re-pushes of the same candidate
    uint32_t mine = 0;
    for (int i0 = 0; i0 < np2; i0 += kScanThreads) {
      const int i = i0 + tid;
      const bool dup = i > 0 && (uint32_t)i < n && s[i] == s[i - 1];
      __syncthreads();
      if (dup) { s[i] = kKeyMax; ++mine; }
      __syncthreads();
    }
    if (mine) atomicAdd(s_dups, mine);
    __syncthreads();
    nuniq = n - *s_dups;
    if (*s_dups) block_bitonic_sort(s, np2);
  }
  const uint32_t keep = min(nuniq, w.nover);
  for (uint32_t i = tid; i < keep; i += kScanThreads) w.buf[(size_t)q * w.cap + i] = s[i];
  if (tid == 0) {
    w.cnt[q] = keep;
    // fewer than N keys passed the current tau: it stays (it is an upper bound of the N-th best key that came from
    // elsewhere -- another rank's pilot in the sharded search; on one GPU it can only be kKeyMax here)
    if (keep >= w.nover) w.tau[q] = s[w.nover - 1];
    // sticky across the scan phases of a batch; the re-scan (dedup) passes clear it
    if (over || dedup) w.ovf[q] = over ? 1u : 0u;
    if (over) atomicAdd(&w.counters[2], 1u);
  }
}

// Common case: one CTA per query, shared memory for `hi` keys only (high occupancy).  Queries
// with more buffered candidates are appended to a list (w.entry_q is free after the scan) and
// handled by compact_big_kernel with a handful of large-shared-memory CTAs.
__global__ void __launch_bounds__(kScanThreads)
compact_small_kernel(ScanWork w, int dedup, uint32_t hi, uint32_t mid) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t s_dups;
  const uint32_t q = blockIdx.x;
  const uint32_t nraw = w.cnt[q];
  if (threadIdx.x == 0 && !dedup) {  // candidate-inflow statistics of the main pass
    atomicAdd(&w.stats[2], (unsigned long long)nraw);
    atomicMax(&w.stats[3], (unsigned long long)nraw);
  }
  if (min(nraw, w.cap) > hi) {
    // two size classes behind the common case: up to `mid` keys (32 KB of shared memory, several CTAs per SM: the
    // typical C5-shape query buffers 1-3k keys) and the heavy tail (one CTA per SM).  The lists share entry_q: the
    // big one grows from the front, the medium one from the back (together at most nq entries).
    if (threadIdx.x == 0) {
      if (min(nraw, w.cap) > mid) w.entry_q[atomicAdd(&w.counters[4], 1u)] = q;
      else w.entry_q[(size_t)w.nq * (w.P ? w.P : 1u) - 1 - atomicAdd(&w.counters[6], 1u)] = q;  // P = 0: brute force
    }
    return;
  }
  compact_one(w, q, dedup, reinterpret_cast<uint64_t*>(smem), &s_dups);
}

__global__ void __launch_bounds__(kScanThreads)
compact_big_kernel(ScanWork w, int dedup, int medium) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t s_dups;
  const uint32_t nlist = w.counters[medium ? 6 : 4];
  const size_t last = (size_t)w.nq * (w.P ? w.P : 1u) - 1;
  for (uint32_t i = blockIdx.x; i < nlist; i += gridDim.x) {
    compact_one(w, w.entry_q[medium ? last - i : i], dedup, reinterpret_cast<uint64_t*>(smem), &s_dups);
    __syncthreads();
  }
}

// ---- debug: int16 scores of one leaf under one uint8 LUT ---------------------------------
// Runs the MAIN scan's scoring path (oct table, score_oct_addr<W, NL, 0>, the same (W, NL) instantiation launch_scan
// picks): the LUT sits in oct lane `lane_q` (the other seven tables are zero), so calls with different lanes cover all
// eight u16 accumulator positions.
// generic instantiation (16 < W <= 32): the pilot's runtime-W scoring over the plain u8 LUT
__global__ void __launch_bounds__(kScanThreads)
leaf_scores_generic_kernel(DevIndex ix, const uint8_t* __restrict__ lut, uint32_t leaf, int16_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = (int)ix.W;
  for (int t = tid; t < W * 32; t += kScanThreads) reinterpret_cast<uint32_t*>(smem)[t] = reinterpret_cast<const uint32_t*>(lut)[t];
  __syncthreads();
  const int nlast = (int)ix.B - 8 * (W - 1);
  const uint32_t gbeg = ix.leaf_goff[leaf], ng = ix.leaf_goff[leaf + 1] - gbeg;
  const uint32_t n = ix.leaf_size[leaf];
  for (uint32_t g = blockIdx.x * kScanWarps + warp; g < ng; g += gridDim.x * kScanWarps) {
    const int sum = score_one_rt(ix.codes + (size_t)(gbeg + g) * W * 32, lane, W, smem, nlast);
    const uint32_t slot = g * 32 + lane;
    if (slot < n) out[slot] = (int16_t)(sum - 128 * (int)ix.B);
  }
}

template <int W, int NL>
__global__ void __launch_bounds__(kScanThreads)
leaf_scores_kernel(DevIndex ix, const uint8_t* __restrict__ lut, uint32_t leaf, int lane_q, uint32_t one,
                   int16_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint2* tables = reinterpret_cast<uint2*>((reinterpret_cast<uintptr_t>(smem) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t* lp[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) lp[i] = (i == lane_q) ? lut : nullptr;
  build_oct_table(tables, lp, W * 128, tid, kScanThreads);
  __syncthreads();
  const int nlast = (int)ix.B - 8 * (W - 1);
  const uint32_t gbeg = ix.leaf_goff[leaf], ng = ix.leaf_goff[leaf + 1] - gbeg;
  const uint32_t n = ix.leaf_size[leaf];
  for (uint32_t g = blockIdx.x * kScanWarps + warp; g < ng; g += gridDim.x * kScanWarps) {
    uint32_t cw[W];
    load_codes<W>(ix.codes + (size_t)(gbeg + g) * W * 32, lane, cw);
    uint32_t ad[8 * W];
    const uint32_t tb32 = (uint32_t)__cvta_generic_to_shared(tables);
#pragma unroll
    for (int j = 0; j < W; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        ad[8 * j + k] = tb32 | ((k == 0) ? ((cw[j] << 3) & 0x78u) : ((cw[j] >> (4 * k - 3)) & 0x78u));
    uint32_t acc[4];
    score_oct_addr<W, NL, 0>(ad, nlast, one, acc);
    // (s0,s2) (s1,s3) (s4,s6) (s5,s7) as u16 lanes
    const uint32_t word = acc[((lane_q >> 2) << 1) | (lane_q & 1)];
    const uint32_t sum = ((lane_q >> 1) & 1) ? (word >> 16) : (word & 0xFFFFu);
    const uint32_t slot = g * 32 + lane;
    if (slot < n) out[slot] = (int16_t)((int)sum - 128 * (int)ix.B);
  }
}

// ---- launchers ---------------------------------------------------------------------------
#ifdef SB_DEV_W  // development builds (make DEVW=1): only the block counts of the bench shapes, minutes less of ptxas
#define SB_DISPATCH_W(Wv, ...)                                                         \
  switch (Wv) {                                                                          \
    case 6: { constexpr int W = 6; __VA_ARGS__; } break;   case 7: { constexpr int W = 7; __VA_ARGS__; } break;   \
    default: return cudaErrorInvalidValue;                                               \
  }
#else
#define SB_DISPATCH_W(Wv, ...)                                                         \
  switch (Wv) {                                                                          \
    case 1: { constexpr int W = 1; __VA_ARGS__; } break;   case 2: { constexpr int W = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int W = 3; __VA_ARGS__; } break;   case 4: { constexpr int W = 4; __VA_ARGS__; } break;   \
    case 5: { constexpr int W = 5; __VA_ARGS__; } break;   case 6: { constexpr int W = 6; __VA_ARGS__; } break;   \
    case 7: { constexpr int W = 7; __VA_ARGS__; } break;   case 8: { constexpr int W = 8; __VA_ARGS__; } break;   \
    case 9: { constexpr int W = 9; __VA_ARGS__; } break;   case 10: { constexpr int W = 10; __VA_ARGS__; } break; \
    case 11: { constexpr int W = 11; __VA_ARGS__; } break; case 12: { constexpr int W = 12; __VA_ARGS__; } break; \
    case 13: { constexpr int W = 13; __VA_ARGS__; } break; case 14: { constexpr int W = 14; __VA_ARGS__; } break; \
    case 15: { constexpr int W = 15; __VA_ARGS__; } break; case 16: { constexpr int W = 16; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;                                               \
  }
#endif

static int pilot_capl(const ScanWork& w) {
  int capl = 1024;  // keys buffered between selections; a typical C2 leaf fits
  while ((uint32_t)capl < 2 * w.nover + kScanThreads) capl <<= 1;
  // large leaves: hold a whole leaf, so the pilot selects once at the end instead of every ~900 keys
  while ((uint32_t)capl < w.pilot_cap && capl < 8192) capl <<= 1;
  return capl;
}
size_t pilot_smem_bytes(const DevIndex& ix, const ScanWork& w) {
  return (size_t)ix.W * 128 * 4 + (size_t)pilot_capl(w) * 8;
}
// the fused LUT build keeps the query and the raw table in the pilot's candidate buffer
bool pilot_can_build_lut(const DevIndex& ix, const ScanWork& w) {
  return ((size_t)((ix.d + 3) & ~3u) + (size_t)ix.B * 16) * 4 <= (size_t)pilot_capl(w) * 8;
}
size_t scan_smem_bytes(const DevIndex& ix, uint32_t quads_per_item) {
  return (size_t)quads_per_item * ix.W * 128 * 8 + 128;  // + alignment slack
}

cudaError_t launch_pilot(const DevIndex& ix, const ScanWork& w, cudaStream_t s) {
  const int capl = pilot_capl(w);
  cudaError_t em = cudaMemsetAsync(w.leaf_cnt, 0, sizeof(uint32_t) * (ix.L + 1), s);
  if (em != cudaSuccess) return em;
  const size_t smem = pilot_smem_bytes(ix, w);
  if (ix.W > 16 && ix.W <= 32) {  // generic instantiation
    cudaError_t e = cudaFuncSetAttribute(pilot_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pilot_kernel<0><<<w.nq, kScanThreads, smem, s>>>(ix, w, capl);
    return cudaGetLastError();
  }
  SB_DISPATCH_W(ix.W, {
    cudaError_t e = cudaFuncSetAttribute(pilot_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pilot_kernel<W><<<w.nq, kScanThreads, smem, s>>>(ix, w, capl);
  });
  return cudaGetLastError();
}

template <int W, int NL, int WIDE>
static cudaError_t launch_scan_t(const DevIndex& ix, const ScanWork& w, int grid, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(scan_main_kernel<W, NL, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (grid <= 0) {  // persistent: exactly as many CTAs as can be resident
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_main_kernel<W, NL, WIDE>, kScanThreads, smem);
    if (e != cudaSuccess) return e;
    grid = sms * (per_sm > 0 ? per_sm : 1);
  }
  scan_main_kernel<W, NL, WIDE><<<grid, kScanThreads, smem, s>>>(ix, w);
  return cudaGetLastError();
}

// The kernel of the next work list, by the expected number of queries per leaf in it (qpl_per_rank x ranks):
//  * wide quads (3 issue slots per four lookups, padding unit four queries, twice the shared-memory bytes per lookup)
//    where leaves hold few queries -- measured cross-over with the octs near 45 queries per leaf (20M x 96, B = 48);
//  * octs (9 issue slots per eight lookups, 128 lookups / clk / SM at the LDS limit) above that;
//  * the tensor-core scan (scan_tc.cu: blocks of 64 queries per leaf) where the blocks fill.
// SCANN_B200_SCAN_WIDE=0|1, SCANN_B200_SCAN_TC=0|1 force a path; *_QPL move the thresholds (tests, tuning).
void scan_prepare_phase(const DevIndex& ix, ScanWork* w) {
  const uint32_t ranks = w->rank_hi > w->rank_lo ? w->rank_hi - w->rank_lo : w->P;
  const float qpl = w->rescan ? 0.f : w->qpl_per_rank * (float)ranks;
  const char* ew = getenv("SCANN_B200_SCAN_WIDE");
  const char* et = getenv("SCANN_B200_SCAN_TC");
  const char* lw = getenv("SCANN_B200_SCAN_WIDE_QPL");
  const char* lt = getenv("SCANN_B200_SCAN_TC_QPL");
  const float wide_limit = lw ? (float)atof(lw) : 45.f;
  // the tensor-core scan is opt-in: bit-exact, but at its current 1.5 ms per C2 batch not faster than the octs (1.41 ms)
  const float tc_limit = lt ? (float)atof(lt) : 3.0e38f;
  const bool tc_ok = scan_tc_supported(ix) && w->lut_e4m3 != nullptr;
  bool tc = tc_ok && !w->rescan && qpl >= tc_limit;
  if (et && (et[0] == '0' || et[0] == '1')) tc = et[0] == '1' && tc_ok;
  bool wide = w->qpl_per_rank > 0.f && qpl <= wide_limit;
  if (ew && (ew[0] == '0' || ew[0] == '1')) wide = ew[0] == '1';
  if (ix.W > 16) { tc = false; wide = true; }  // 128 < B <= 256: the generic kernel (wide quads, unpacked thresholds)
  if (w->max_gpt_simt == 0) w->max_gpt_simt = w->max_gpt;
  if (tc) {
    // items = (leaf, block of 64 queries): eight "octs" per item, the whole leaf in one tile
    w->scan_mode = 2; w->quads_per_item = 8; w->max_gpt = 0x7FFFFFFFu;
  } else {
    // sixteen queries (two octs) per item for dense lists, eight (two wide quads) for sparse ones -- few leaves hold
    // more there, the tables take half the shared memory and the per-item candidate stage has 128 entries per query
    w->scan_mode = wide ? 1u : 0u; w->quads_per_item = wide ? 1u : 2u; w->max_gpt = w->max_gpt_simt;
  }
}

cudaError_t launch_scan(const DevIndex& ix, const ScanWork& w, int grid, cudaStream_t s) {
  if (w.scan_mode == 2) return launch_scan_tc(ix, w, s);
  const bool wide = w.scan_mode == 1;
  const size_t smem = scan_smem_bytes(ix, w.quads_per_item * (wide ? 2u : 1u));
  if (ix.W > 16 && ix.W <= 32) {  // generic instantiation, wide quads only (scan_prepare_phase)
    if (!wide) return cudaErrorInvalidValue;
    return launch_scan_t<0, 0, 1>(ix, w, grid, smem, s);
  }
  // the common block counts (B % 8 == 0, 2, 4) get a kernel without padded lookups
  const int nlast = (int)ix.B - 8 * ((int)ix.W - 1);
  if (wide) {
    SB_DISPATCH_W(ix.W, {
      if (nlast == 8) return launch_scan_t<W, 8, 1>(ix, w, grid, smem, s);
      if (nlast == 2) return launch_scan_t<W, 2, 1>(ix, w, grid, smem, s);
      if (nlast == 4) return launch_scan_t<W, 4, 1>(ix, w, grid, smem, s);
      return launch_scan_t<W, 0, 1>(ix, w, grid, smem, s);
    });
    return cudaGetLastError();
  }
  SB_DISPATCH_W(ix.W, {
    if (nlast == 8) return launch_scan_t<W, 8, 0>(ix, w, grid, smem, s);
    if (nlast == 2) return launch_scan_t<W, 2, 0>(ix, w, grid, smem, s);
    if (nlast == 4) return launch_scan_t<W, 4, 0>(ix, w, grid, smem, s);
    return launch_scan_t<W, 0, 0>(ix, w, grid, smem, s);
  });
  return cudaGetLastError();
}

cudaError_t launch_compact(const DevIndex& ix, const ScanWork& w, bool dedup, cudaStream_t s, int* n_launched) {
  (void)ix;
  int np2 = 2;
  while ((uint32_t)np2 < w.cap) np2 <<= 1;
  const uint32_t small = (uint32_t)np2 < 1024u ? (uint32_t)np2 : 1024u;
  // the medium class pays where many queries buffer 1-4k keys (large leaves, brute-force rounds); on C2-size work
  // its extra launch costs more than the heavy-tail kernel loses (0.077 -> 0.096 ms), so it is off there
  const bool use_mid = w.stage != 0 || w.P == 0;
  const uint32_t mid = !use_mid ? small : ((uint32_t)np2 < 4096u ? (uint32_t)np2 : 4096u);
  cudaError_t e = cudaMemsetAsync(w.counters + 4, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(w.counters + 6, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  compact_small_kernel<<<w.nq, kScanThreads, (size_t)small * 8, s>>>(w, dedup ? 1 : 0, small, mid);
  if ((uint32_t)np2 > small) {
    e = cudaFuncSetAttribute(compact_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)np2 * 8));
    if (e != cudaSuccess) return e;
    if (mid > small) compact_big_kernel<<<148 * 6, kScanThreads, (size_t)mid * 8, s>>>(w, dedup ? 1 : 0, 1);
    if ((uint32_t)np2 > mid) compact_big_kernel<<<296, kScanThreads, (size_t)np2 * 8, s>>>(w, dedup ? 1 : 0, 0);
    if (n_launched) *n_launched = 1 + (mid > small ? 1 : 0) + ((uint32_t)np2 > mid ? 1 : 0);
  } else if (n_launched) {
    *n_launched = 1;
  }
  return cudaGetLastError();
}

template <int W, int NL>
static cudaError_t launch_leaf_scores_t(const DevIndex& ix, const uint8_t* lut, uint32_t leaf, int lane_q, int16_t* out,
                                        cudaStream_t s) {
  const size_t smem = scan_smem_bytes(ix, 1);
  cudaError_t e = cudaFuncSetAttribute(leaf_scores_kernel<W, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  leaf_scores_kernel<W, NL><<<8, kScanThreads, smem, s>>>(ix, lut, leaf, lane_q, 1u, out);
  return cudaGetLastError();
}

cudaError_t launch_leaf_scores(const DevIndex& ix, const uint8_t* lut, uint32_t leaf, int16_t* out, cudaStream_t s) {
  if (ix.W > 16 && ix.W <= 32) {
    leaf_scores_generic_kernel<<<8, kScanThreads, (size_t)ix.W * 128, s>>>(ix, lut, leaf, out);
    return cudaGetLastError();
  }
  const int nlast = (int)ix.B - 8 * ((int)ix.W - 1);
  const int lane_q = (int)(leaf & 7u);
  SB_DISPATCH_W(ix.W, {
    if (nlast == 8) return launch_leaf_scores_t<W, 8>(ix, lut, leaf, lane_q, out, s);
    if (nlast == 2) return launch_leaf_scores_t<W, 2>(ix, lut, leaf, lane_q, out, s);
    if (nlast == 4) return launch_leaf_scores_t<W, 4>(ix, lut, leaf, lane_q, out, s);
    return launch_leaf_scores_t<W, 0>(ix, lut, leaf, lane_q, out, s);
  });
  return cudaGetLastError();
}

}  // namespace sb
