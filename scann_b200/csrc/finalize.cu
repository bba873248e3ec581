// finalize.cu -- slot -> datapoint id, SOAR de-duplication, exact reordering, final top-k,
// and the merge of sharded partial results.
//
// Replaces (a2/a10/a11) of SURVEY.md section 8:
//   global index decode                          tree_x_hybrid/tree_ah_hybrid_residual.cc:771-778
//   DeduplicateDatabaseSpilledResults            tree_x_hybrid/internal/utils.cc:135-156
//   ExactReorderingHelper::ComputeDistancesForReordering   utils/reordering_helper.cc:257-283
//     -> DenseDistanceOneToMany (gather by index) distance_measures/one_to_many/one_to_many_symmetric.h:373-503
//   SortAndDropResults                           base/single_machine_base.cc:872-901
//   ReshapeBatchedNNResult                       scann_ops/cc/scann.h:165-180
#include <math.h>

#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"

namespace sb {

constexpr int kFinThreads = 128;
constexpr uint32_t kInvalidId = 0xFFFFFFFFu;

// int8 (fixed point) reordering: FixedPointFloatDense{DotProduct,SquaredL2}ReorderingHelper
// (utils/reordering_helper.cc:430-441,610-618).  `qp` = the query scaled by the inverse multipliers
// (PrepareForAsymmetricScalarQuantizedDotProduct), val = -<qp, float(x)> in the order of
// OneToManyAsymmetricTemplate<.., int8_t> on AVX2 (one_to_many_asymmetric_impl.inc:296-353): eight fnmadd lanes over
// whole groups of 8 dims, one 4-wide step into lanes 0..3, HorizontalSum3X = ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)),
// the remaining dims one by one on the scalar.  Squared L2: (|q|^2 + dp_norm) + 2 val (SetSquaredL2DistanceFunctor).
__device__ __forceinline__ float exact_distance_i8(const DevIndex& ix, const float* __restrict__ qp, float qnorm,
                                                   uint32_t dp) {
  const uint32_t row = ix.dp_row ? ix.dp_row[dp] : dp;
  const int8_t* __restrict__ x = ix.dataset_i8 + (size_t)row * ix.d;
  const uint32_t n = ix.d;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) a[l] = __fmaf_rn(-qp[j + l], (float)__ldg(x + j + l), a[l]);
  }
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = __fmaf_rn(-qp[j + l], (float)__ldg(x + j + l), a[l]);
    j += 4;
  }
  float r = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[4]), __fadd_rn(a[2], a[6])),
                      __fadd_rn(__fadd_rn(a[1], a[5]), __fadd_rn(a[3], a[7])));
  for (; j < n; ++j) r = __fmaf_rn(-qp[j], (float)__ldg(x + j), r);
  if (ix.distance == 0) return r;
  return __fadd_rn(__fadd_rn(qnorm, ix.i8_dp_norm[dp]), __fmul_rn(2.0f, r));
}

__device__ __forceinline__ float exact_distance(const DevIndex& ix, const float* __restrict__ q,
                                                uint32_t dp) {
  const uint32_t row = ix.dp_row ? ix.dp_row[dp] : dp;
  auto lq = [&](uint32_t i) { return q[i]; };
  if (!ix.dataset) {
    // bfloat16 reordering (utils/reordering_helper.cc:745-757): f32 query x bf16 row, f32 FMA, in the order of the
    // reference's asymmetric one-to-many kernel (exact_math.cuh)
    const uint16_t* xb = ix.dataset_bf16 + (size_t)row * ix.d;
    auto lb = [&](uint32_t i) { return __uint_as_float((uint32_t)__ldg(xb + i) << 16); };
    return ix.distance == 0 ? neg_dot_asym_order(lq, lb, ix.d) : sql2_asym_order(lq, lb, ix.d);
  }
  const float* x = ix.dataset + (size_t)row * ix.d;
  auto lx = [&](uint32_t i) { return __ldg(x + i); };
  if (ix.distance == 0) return ix.d < 8 ? neg_dot_small(lq, lx, ix.d) : neg_dot_avx2_order(lq, lx, ix.d);
  return ix.d < 8 ? sql2_small(lq, lx, ix.d) : sql2_avx2_order(lq, lx, ix.d);
}

// The same arithmetic as exact_distance (dims >= 8), spread over 8 consecutive lanes: lane l
// owns AVX lane l of the reference kernel (one_to_many_symmetric.h:373-503), so a row's 32-byte
// sectors are fetched by 8 threads at once instead of one thread walking the row.  All 32 lanes
// of the warp must call this together; the result is valid in the lanes with l == 0.
__device__ __forceinline__ float exact_distance_lanes8(const DevIndex& ix, const float* __restrict__ q,
                                                       uint32_t dp, int l) {
  const uint32_t row = ix.dp_row ? ix.dp_row[dp] : dp;
  const float* __restrict__ x = ix.dataset + (size_t)row * ix.d;
  const uint32_t n = ix.d;
  const bool dot = ix.distance == 0;
  float a = 0.f;
  uint32_t j = 0;
  // four 32-byte sectors of the gathered row in flight per 8-lane group before the first FMA needs one
  for (; j + 32 <= n; j += 32) {
    float xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = __ldg(x + j + 8 * u + l);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float qv = q[j + 8 * u + l];
      if (dot) a = __fmaf_rn(-qv, xv[u], a);
      else { const float t = __fsub_rn(qv, xv[u]); a = __fmaf_rn(t, t, a); }
    }
  }
  for (; j + 8 <= n; j += 8) {
    const float xv = __ldg(x + j + l), qv = q[j + l];
    if (dot) a = __fmaf_rn(-qv, xv, a);
    else { const float t = __fsub_rn(qv, xv); a = __fmaf_rn(t, t, a); }
  }
  float b = __fadd_rn(__shfl_down_sync(0xFFFFFFFFu, a, 4, 8), a);  // lanes 0..3: a[l+4] + a[l]
  if (j + 4 <= n) {
    if (l < 4) {
      const float xv = __ldg(x + j + l), qv = q[j + l];
      if (dot) b = __fmaf_rn(-qv, xv, b);
      else { const float t = __fsub_rn(qv, xv); b = __fmaf_rn(t, t, b); }
    }
    j += 4;
  }
  if (j + 2 <= n) {
    if (l == 2 || l == 3) {
      const float xv = __ldg(x + j + (l - 2)), qv = q[j + (l - 2)];
      if (dot) b = __fmaf_rn(-qv, xv, b);
      else { const float t = __fsub_rn(qv, xv); b = __fmaf_rn(t, t, b); }
    }
    j += 2;
  }
  const float t2 = __fadd_rn(b, __shfl_down_sync(0xFFFFFFFFu, b, 2, 8));   // lanes 0,1: b0+b2, b1+b3
  float r = __fadd_rn(t2, __shfl_down_sync(0xFFFFFFFFu, t2, 1, 8));        // lane 0: (b0+b2)+(b1+b3)
  if (j < n && l == 0) {
    const float xv = __ldg(x + j), qv = q[j];
    if (dot) r = __fmaf_rn(-qv, xv, r);
    else { const float t = __fsub_rn(qv, xv); r = __fmaf_rn(t, t, r); }
  }
  return r;
}

// bf16 rows: lane l owns AVX lane l of the reference's ASYMMETRIC kernel (one_to_many_asymmetric_impl.inc:296-353, see
// neg_dot_asym_order in exact_math.cuh): the 4-wide step goes into lanes 0..3 BEFORE the lanes are summed, and every
// remaining dim is fused onto the sum.
__device__ __forceinline__ float exact_distance_lanes8_bf16(const DevIndex& ix, const float* __restrict__ q,
                                                       uint32_t dp, int l) {
  const uint32_t row = ix.dp_row ? ix.dp_row[dp] : dp;
  const uint16_t* __restrict__ x = ix.dataset_bf16 + (size_t)row * ix.d;
  const uint32_t n = ix.d;
  const bool dot = ix.distance == 0;
  float a = 0.f;
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8) {
    const float xv = __uint_as_float((uint32_t)__ldg(x + j + l) << 16), qv = q[j + l];
    if (dot) a = __fmaf_rn(-qv, xv, a);
    else { const float t = __fsub_rn(qv, xv); a = __fmaf_rn(t, t, a); }
  }
  if (j + 4 <= n) {
    if (l < 4) {
      const float xv = __uint_as_float((uint32_t)__ldg(x + j + l) << 16), qv = q[j + l];
      if (dot) a = __fmaf_rn(-qv, xv, a);
      else { const float t = __fsub_rn(qv, xv); a = __fmaf_rn(t, t, a); }
    }
    j += 4;
  }
  const float b = __fadd_rn(a, __shfl_down_sync(0xFFFFFFFFu, a, 4, 8));    // lanes 0..3: a[l] + a[l+4]
  const float t2 = __fadd_rn(b, __shfl_down_sync(0xFFFFFFFFu, b, 2, 8));   // lanes 0,1: (a0+a4)+(a2+a6), (a1+a5)+(a3+a7)
  float r = __fadd_rn(t2, __shfl_down_sync(0xFFFFFFFFu, t2, 1, 8));        // lane 0
  if (l == 0) {
    for (; j < n; ++j) {
      const float xv = __uint_as_float((uint32_t)__ldg(x + j) << 16), qv = q[j];
      if (dot) r = __fmaf_rn(-qv, xv, r);
      else { const float t = __fsub_rn(qv, xv); r = __fmaf_rn(t, t, r); }
    }
  }
  return r;
}

// Bitonic sort of (u64 key, u32 payload) pairs in shared memory.
__device__ __forceinline__ void block_bitonic_sort_kv(uint64_t* s, uint32_t* pay, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int r = l | j;
        const uint64_t a = s[l], b = s[r];
        const bool up = (l & k) == 0;
        if ((a > b) == up) {
          s[l] = b; s[r] = a;
          const uint32_t pa = pay[l];
          pay[l] = pay[r]; pay[r] = pa;
        }
      }
      __syncthreads();
    }
  }
}

// SOAR de-duplication on a sorted-by-(dp, score) array ka[0..n): writes (score', dp) keys into kb
// (kKeyMax for removed / padding entries) and returns the number of removed entries via *removed.
// kpay/kpay_out optionally carry a payload (index of the first copy).
__device__ __forceinline__ void soar_merge_sorted(const uint64_t* ka, uint64_t* kb, const uint32_t* pin,
                                                  uint32_t* pout, uint32_t n, int np2, uint32_t* s_removed) {
  uint32_t removed = 0;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    uint64_t k = kKeyMax;
    uint32_t p = 0;
    if ((uint32_t)i < n) {
      const uint64_t cur = ka[i];
      const uint32_t dp = (uint32_t)(cur >> 32);
      const bool dup_of_prev = i > 0 && (uint32_t)(ka[i - 1] >> 32) == dp;
      const bool has_next = (uint32_t)(i + 1) < n && (uint32_t)(ka[i + 1] >> 32) == dp;
      if (dup_of_prev) {
        ++removed;
      } else {
        float sc = ord2f((uint32_t)cur);
        if (has_next) {  // 0.5f * a + 0.5f * b (internal/utils.cc:146)
          const float other = ord2f((uint32_t)ka[i + 1]);
          sc = __fadd_rn(__fmul_rn(0.5f, sc), __fmul_rn(0.5f, other));
        }
        k = make_key(sc, dp);
        if (pin) p = pin[i];
      }
    }
    kb[i] = k;
    if (pout) pout[i] = p;
  }
  if (removed) atomicAdd(s_removed, removed);
}

// One CTA per query.  Input: buf[q][0..cnt) sorted ascending (score, global slot), cnt <= nover.
__global__ void __launch_bounds__(kFinThreads)
finalize_kernel(DevIndex ix, ScanWork w, FinalizeArgs a, int np2) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* ka = reinterpret_cast<uint64_t*>(smem);  // [np2]
  uint64_t* kb = ka + np2;                            // [np2]
  float* sq = reinterpret_cast<float*>(kb + np2);     // [D]
  __shared__ uint32_t s_removed;
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x;
  uint32_t n = min(w.cnt[q], w.nover);
  const uint64_t* src = w.buf + (size_t)q * w.cap;
  const bool i8 = ix.dataset_i8 != nullptr;
  const bool reorder = ix.dataset != nullptr || ix.dataset_bf16 != nullptr || i8;
  float* sqp = sq + ((ix.d + 3) & ~3u);               // [D] int8 reordering: the query scaled by the inverse multipliers
  __shared__ float s_qnorm;
  for (uint32_t i = tid; i < ix.d; i += kFinThreads) {
    const float v = a.q[(size_t)q * ix.d + i];
    sq[i] = v;
    if (i8) sqp[i] = __fmul_rn(ix.i8_inv_mult[i], v);
  }
  if (tid == 0) s_removed = 0;
  __syncthreads();
  if (i8 && ix.distance != 0 && tid == 0) {
    // float(SquaredL2Norm(query)): DenseSingleAccumulate, four strided double accumulators (utils/reduction.h:357-390)
    double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    uint32_t k = 0;
    const uint32_t D = ix.d;
    for (; k + 4 <= D; k += 4) {
      r0 = __dadd_rn(r0, __dmul_rn((double)sq[k], (double)sq[k]));
      r1 = __dadd_rn(r1, __dmul_rn((double)sq[k + 1], (double)sq[k + 1]));
      r2 = __dadd_rn(r2, __dmul_rn((double)sq[k + 2], (double)sq[k + 2]));
      r3 = __dadd_rn(r3, __dmul_rn((double)sq[k + 3], (double)sq[k + 3]));
    }
    r2 = __dadd_rn(r2, r3);
    if (k + 2 <= D) {
      r0 = __dadd_rn(r0, __dmul_rn((double)sq[k], (double)sq[k]));
      r1 = __dadd_rn(r1, __dmul_rn((double)sq[k + 1], (double)sq[k + 1]));
      k += 2;
    }
    r1 = __dadd_rn(r1, r2);
    if (k < D) r0 = __dadd_rn(r0, __dmul_rn((double)sq[k], (double)sq[k]));
    s_qnorm = (float)__dadd_rn(r0, r1);
  }
  if (i8) __syncthreads();
  const float qnorm = (i8 && ix.distance != 0) ? s_qnorm : 0.f;

  if ((a.part_ids || a.part_rec) && a.part_limit) {
    // sampled global threshold (sharded.cu): local candidates whose score exceeds it cannot be among the global N' best
    const uint32_t lim = a.part_limit[q];
    uint32_t lo = 0, hi = n;  // first index whose score word is > lim (src is sorted)
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if ((uint32_t)(src[mid] >> 32) <= lim) lo = mid + 1; else hi = mid;
    }
    n = lo;
  }
  if (a.part_ids || a.part_rec) {
    // Sharded mode: emit the raw over-retrieved candidates (before SOAR de-duplication, which is
    // only exact on the global top list) with their exact distances: four arrays, or packed 16-byte records
    // {tie-break key, id, exact distance}.
    const bool packed = a.part_ids == nullptr;
    for (uint32_t i = tid; i < a.part_cap; i += kFinThreads) {
      const size_t o = (size_t)q * a.part_cap + i;
      uint32_t id = kInvalidId;
      uint64_t tie = kKeyMax;
      float ah = INFINITY, ex = INFINITY;
      if (i < n) {
        const uint64_t s = src[i];
        const uint32_t gslot = (uint32_t)s;
        id = ix.key_by_dp ? gslot : ix.slot_dp[gslot];
        tie = ix.key_by_dp ? s : ((s & 0xFFFFFFFF00000000ull) | (ix.slot_tie ? ix.slot_tie[gslot] : gslot));
        ah = ord2f((uint32_t)(s >> 32));
        if (!reorder) ex = ah;
        else if (i8) ex = exact_distance_i8(ix, sqp, qnorm, id);
        else if (ix.d < 8) ex = exact_distance(ix, sq, id);
      }
      if (packed) {
        a.part_rec[o] = make_uint4((uint32_t)tie, (uint32_t)(tie >> 32), id, __float_as_uint(ex));
      } else {
        a.part_ids[o] = id; a.part_tie[o] = tie; a.part_ah[o] = ah; a.part_exact[o] = ex;
      }
    }
    if (reorder && ix.d >= 8 && !i8) {
      if (packed) __syncthreads();  // the exact distance is patched into the records written above
      const int l = tid & 7, grp = tid >> 3;
      for (uint32_t c0 = 0; c0 < n; c0 += kFinThreads / 8) {
        const uint32_t c = c0 + grp;
        const bool valid = c < n;
        const uint32_t gslot = (uint32_t)src[valid ? c : 0];
        const uint32_t dp = ix.key_by_dp ? gslot : ix.slot_dp[gslot];
        const float dist = ix.dataset ? exact_distance_lanes8(ix, sq, dp, l) : exact_distance_lanes8_bf16(ix, sq, dp, l);
        if (valid && l == 0) {
          if (packed) reinterpret_cast<float*>(a.part_rec + (size_t)q * a.part_cap + c)[3] = dist;
          else a.part_exact[(size_t)q * a.part_cap + c] = dist;
        }
      }
    }
    return;
  }

  uint32_t m;  // candidates that go to reordering
  if (ix.disjoint) {
    // (score, slot) order is already final; keys become (score, dp)
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t k = kKeyMax;
      if ((uint32_t)i < n) {
        const uint64_t s = src[i];
        k = ix.key_by_dp ? s : ((s & 0xFFFFFFFF00000000ull) | ix.slot_dp[(uint32_t)s]);
      }
      kb[i] = k;
    }
    m = min(n, a.npre);
    __syncthreads();
  } else {
    // SOAR de-duplication (DeduplicateDatabaseSpilledResults, tree_x_hybrid/internal/utils.cc:135-156) without sorting by
    // datapoint: the two copies of a datapoint find each other through a hash table on the id; the copy that comes
    // first in (score, slot) order survives with 0.5 a + 0.5 b, the other is dropped.  One sort by (score', id) follows.
    uint32_t* tab = reinterpret_cast<uint32_t*>(sq + 2 * ((ix.d + 3) & ~3u));  // [2 np2] slot -> list position + 1
    uint32_t* dpv = tab + 2 * np2;                                              // [np2] datapoint id
    int* partner = reinterpret_cast<int*>(dpv + np2);                           // [np2] position of the other copy or -1
    const uint32_t hmask = 2u * (uint32_t)np2 - 1u;
    for (int i = tid; i < 2 * np2; i += kFinThreads) tab[i] = 0u;
    for (int i = tid; i < np2; i += kFinThreads) {
      dpv[i] = (uint32_t)i < n ? ix.slot_dp[(uint32_t)src[i]] : kInvalidId;
      partner[i] = -1;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kFinThreads) {
      const uint32_t dp = dpv[i];
      uint32_t h = (dp * 2654435761u) & hmask;
      for (;;) {
        const uint32_t old = atomicCAS(&tab[h], 0u, i + 1u);
        if (old == 0u) break;
        if (dpv[old - 1u] == dp) { partner[i] = (int)(old - 1u); partner[old - 1u] = (int)i; break; }
        h = (h + 1u) & hmask;
      }
    }
    __syncthreads();
    uint32_t removed = 0;
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t k = kKeyMax;
      if ((uint32_t)i < n) {
        const int j = partner[i];
        if (j >= 0 && j < i) {
          ++removed;
        } else {
          float sc = ord2f((uint32_t)(src[i] >> 32));
          if (j > i) sc = __fadd_rn(__fmul_rn(0.5f, sc), __fmul_rn(0.5f, ord2f((uint32_t)(src[j] >> 32))));
          k = make_key(sc, dpv[i]);
        }
      }
      kb[i] = k;
    }
    if (removed) atomicAdd(&s_removed, removed);
    __syncthreads();
    block_bitonic_sort(kb, np2);
    m = min(n - s_removed, a.npre);
  }
  // the final sort covers the m reordered candidates only
  int np2k = 2;
  while ((uint32_t)np2k < m) np2k <<= 1;
  if (np2k > np2) np2k = np2;
  for (int i = tid; i < np2; i += kFinThreads) ka[i] = kKeyMax;
  __syncthreads();
  if (reorder && ix.d >= 8 && !i8) {
    // 8 lanes per candidate row, 16 rows per pass
    const int l = tid & 7, grp = tid >> 3;
    for (uint32_t c0 = 0; c0 < m; c0 += kFinThreads / 8) {
      const uint32_t c = c0 + grp;
      const bool valid = c < m;
      const uint32_t dp = (uint32_t)kb[valid ? c : 0];
      const float dist = ix.dataset ? exact_distance_lanes8(ix, sq, dp, l) : exact_distance_lanes8_bf16(ix, sq, dp, l);
      if (valid && l == 0) ka[c] = make_key(dist, dp);
    }
  } else {
    for (uint32_t i = tid; i < m; i += kFinThreads) {
      const uint64_t c = kb[i];
      const uint32_t dp = (uint32_t)c;
      const float dist = i8 ? exact_distance_i8(ix, sqp, qnorm, dp)
                            : (reorder ? exact_distance(ix, sq, dp) : ord2f((uint32_t)(c >> 32)));
      ka[i] = make_key(dist, dp);
    }
  }
  __syncthreads();
  block_bitonic_sort(ka, np2k);
  const uint32_t kk = min(a.k, m);
  const float mulr = ix.distance == 0 ? -1.0f : 1.0f;  // scann.cc:364-369
  for (uint32_t i = tid; i < a.out_k; i += kFinThreads) {
    uint32_t id = 0;
    float dist = __uint_as_float(0x7FC00000u);  // quiet NaN padding (scann.h:175-178)
    if (i < kk) {
      id = (uint32_t)ka[i];
      dist = mulr * ord2f((uint32_t)(ka[i] >> 32));
    }
    a.out_idx[(size_t)q * a.out_k + i] = id;
    a.out_dist[(size_t)q * a.out_k + i] = dist;
  }
}

cudaError_t launch_finalize(const DevIndex& ix, const ScanWork& w, const FinalizeArgs& a, cudaStream_t s) {
  int np2 = 2;
  while ((uint32_t)np2 < w.nover) np2 <<= 1;
  // ka, kb; the query (twice: int8 reordering keeps a scaled copy); hash table [2 np2] + ids [np2] + partners [np2]
  const size_t smem = (size_t)np2 * 16 + (((size_t)ix.d + 3) & ~(size_t)3) * 4 * 2 + (size_t)np2 * 16;
  cudaError_t e = cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  finalize_kernel<<<w.nq, kFinThreads, smem, s>>>(ix, w, a, np2);
  return cudaGetLastError();
}

// Sampled global threshold of the sharded search.  Every rank publishes, per query, the score word of every 16th key of
// its sorted local candidate list (positions 15, 31, ...).  If m of a rank's samples are <= x, that rank holds at
// least 16 m candidates with score <= x; so with `need` = ceil(N' / 16), the need-th smallest sample over all ranks is
// an upper bound of the global N'-th best score, and at most N' + 15 * world (+ ties) candidates lie below it.  Ranks
// then reorder and send only those, instead of their whole local top-N' (world x N' in total).
__global__ void sample_scores_kernel(const uint64_t* __restrict__ buf, const uint32_t* __restrict__ cnt, uint32_t cap,
                                     uint32_t nover, uint32_t nq, uint32_t S, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * S) return;
  const uint32_t q = i / S, j = i - q * S;
  const uint32_t pos = 16 * (j + 1) - 1;
  const uint32_t n = min(cnt[q], nover);
  out[i] = pos < n ? (uint32_t)(buf[(size_t)q * cap + pos] >> 32) : 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(kFinThreads)
sample_threshold_kernel(const uint32_t* __restrict__ samples, int world, uint32_t nq, uint32_t S, uint32_t need,
                        uint32_t* __restrict__ limit, int np2) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* v = reinterpret_cast<uint32_t*>(smem);
  const uint32_t q = blockIdx.x;
  const int total = world * (int)S;
  for (int i = threadIdx.x; i < np2; i += kFinThreads)
    v[i] = i < total ? samples[((size_t)(i / S) * nq + q) * S + (i % S)] : 0xFFFFFFFFu;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += kFinThreads) {
        const int l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int r = l | j;
        const uint32_t a = v[l], b = v[r];
        const bool up = (l & k) == 0;
        if ((a > b) == up) { v[l] = b; v[r] = a; }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) limit[q] = (need >= 1 && (int)need <= total) ? v[need - 1] : 0xFFFFFFFFu;
}

cudaError_t launch_sample_scores(const ScanWork& w, uint32_t S, uint32_t* out, cudaStream_t s) {
  const uint32_t total = w.nq * S;
  if (!total) return cudaSuccess;
  sample_scores_kernel<<<(total + 255) / 256, 256, 0, s>>>(w.buf, w.cnt, w.cap, w.nover, w.nq, S, out);
  return cudaGetLastError();
}

cudaError_t launch_sample_threshold(const uint32_t* samples, int world, uint32_t nq, uint32_t S, uint32_t nover,
                                    uint32_t* limit, cudaStream_t s) {
  if (!nq) return cudaSuccess;
  int np2 = 2;
  while (np2 < world * (int)S) np2 <<= 1;
  const size_t smem = (size_t)np2 * 4;
  cudaError_t e = cudaFuncSetAttribute(sample_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  sample_threshold_kernel<<<nq, kFinThreads, smem, s>>>(samples, world, nq, S, (nover + 15) / 16, limit, np2);
  return cudaGetLastError();
}

// Merge `world` partial lists of one query: global top-nover by the (AH score, unsharded slot) key, SOAR
// de-duplication, top-npre by (score, id), then top-k by (exact distance, id).  Reproduces the single-GPU result
// exactly.  Two record sources: the four all-gathered arrays [world][nq][n_cand] (kPacked =
// false, every rank merges every query) or packed records [world][nq_slice][n_cand] of the queries this rank owns.
struct MergeSrc {
  const uint32_t* ids; const uint64_t* tie; const float* exact;  // arrays
  const uint4* rec;                                              // packed
  uint32_t nq_stride;                                            // queries per rank block
  int n_cand;
};
template <bool kPacked>
__global__ void __launch_bounds__(kFinThreads)
merge_partials_kernel(int distance, int disjoint, MergeSrc src, int world,
                      uint32_t nover, uint32_t npre, uint32_t k,
                      uint32_t* out_idx, float* out_dist, uint32_t out_k, int np2) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* ka = reinterpret_cast<uint64_t*>(smem);        // [np2]
  uint64_t* kb = ka + np2;                                  // [np2]
  uint32_t* pa = reinterpret_cast<uint32_t*>(kb + np2);     // [np2] payload: record offset
  uint32_t* pb = pa + np2;                                  // [np2]
  uint64_t* ks = reinterpret_cast<uint64_t*>(pb + np2);     // [world][ml] the lists' keys (ml = min(n_cand, nover))
  __shared__ uint32_t s_removed, s_valid;
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x;
  const int n_cand = src.n_cand;
  const int ml = min(n_cand, (int)nover);  // a list's entries past its nover-th cannot be among the global nover best
  auto rec = [&](int i) { return ((size_t)(i / n_cand) * src.nq_stride + q) * n_cand + (i % n_cand); };
  auto rec_id = [&](size_t o) -> uint32_t { return kPacked ? src.rec[o].z : src.ids[o]; };
  auto rec_tie = [&](size_t o) -> uint64_t {
    if (kPacked) { const uint4 r = src.rec[o]; return ((uint64_t)r.y << 32) | r.x; }
    return src.tie[o];
  };
  auto rec_exact = [&](size_t o) -> float { return kPacked ? __uint_as_float(src.rec[o].w) : src.exact[o]; };
  if (tid == 0) { s_removed = 0; s_valid = 0; }
  for (int i = tid; i < np2; i += kFinThreads) { ka[i] = kKeyMax; pa[i] = 0; }
  __syncthreads();
  // Every list arrives sorted by key (keys are unique: (score, unsharded slot)), so the global rank of an element is
  // its position in its own list plus, for every other list, the number of keys below it (a binary search in shared
  // memory).  Elements of rank < nover drop straight into their place: a `world`-way merge without a sort.
  uint32_t nvalid = 0;
  for (int e = tid; e < world * ml; e += kFinThreads) {
    const int r = e / ml, i = e - r * ml;
    const size_t o = rec(r * n_cand + i);
    const bool ok = rec_id(o) != kInvalidId;
    ks[e] = ok ? rec_tie(o) : kKeyMax;
    nvalid += ok ? 1u : 0u;
  }
  if (nvalid) atomicAdd(&s_valid, nvalid);
  __syncthreads();
  for (int e = tid; e < world * ml; e += kFinThreads) {
    const uint64_t key = ks[e];
    if (key == kKeyMax) continue;
    const int r = e / ml;
    uint32_t rank = (uint32_t)(e - r * ml);
    for (int r2 = 0; r2 < world && rank < nover; ++r2) {
      if (r2 == r) continue;
      const uint64_t* l = ks + r2 * ml;
      int lo = 0, hi = ml;  // first index with l[idx] > key
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l[mid] < key) lo = mid + 1; else hi = mid;
      }
      rank += (uint32_t)lo;
    }
    if (rank < nover) { ka[rank] = key; pa[rank] = (uint32_t)(r * n_cand + (e - r * ml)); }
  }
  __syncthreads();
  const uint32_t n = min(s_valid, nover);  // global over-retrieved candidate list
  uint32_t m;
  if (disjoint) {
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t key = kKeyMax;
      if ((uint32_t)i < n) key = (ka[i] & 0xFFFFFFFF00000000ull) | rec_id(rec((int)pa[i]));
      kb[i] = key;
      pb[i] = pa[i];
    }
    m = min(n, npre);
    __syncthreads();
  } else {
    // SOAR de-duplication (DeduplicateDatabaseSpilledResults, tree_x_hybrid/internal/utils.cc:135-156) without sorting by
    // datapoint: the two copies of a datapoint find each other through a hash table on the id; the copy that comes
    // first in (score, slot) order survives with 0.5 a + 0.5 b, the other is dropped.  One sort by (score', id) follows.
    uint32_t* tab = reinterpret_cast<uint32_t*>(ks + (size_t)world * ml);  // [2 np2] slot -> list position + 1
    uint32_t* dpv = tab + 2 * np2;                                          // [np2] datapoint id
    int* partner = reinterpret_cast<int*>(dpv + np2);                       // [np2] position of the other copy or -1
    const uint32_t hmask = 2u * (uint32_t)np2 - 1u;
    for (int i = tid; i < 2 * np2; i += kFinThreads) tab[i] = 0u;
    for (int i = tid; i < np2; i += kFinThreads) {
      dpv[i] = (uint32_t)i < n ? rec_id(rec((int)pa[i])) : kInvalidId;
      partner[i] = -1;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kFinThreads) {
      const uint32_t dp = dpv[i];
      uint32_t h = (dp * 2654435761u) & hmask;
      for (;;) {
        const uint32_t old = atomicCAS(&tab[h], 0u, i + 1u);
        if (old == 0u) break;
        if (dpv[old - 1u] == dp) { partner[i] = (int)(old - 1u); partner[old - 1u] = (int)i; break; }
        h = (h + 1u) & hmask;
      }
    }
    __syncthreads();
    uint32_t removed = 0;
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t key = kKeyMax;
      if ((uint32_t)i < n) {
        const int j = partner[i];
        if (j >= 0 && j < i) {
          ++removed;
        } else {
          float sc = ord2f((uint32_t)(ka[i] >> 32));
          if (j > i) sc = __fadd_rn(__fmul_rn(0.5f, sc), __fmul_rn(0.5f, ord2f((uint32_t)(ka[j] >> 32))));
          key = make_key(sc, dpv[i]);
        }
      }
      kb[i] = key;
      pb[i] = pa[i];
    }
    if (removed) atomicAdd(&s_removed, removed);
    __syncthreads();
    block_bitonic_sort_kv(kb, pb, np2);
    m = min(n - s_removed, npre);
  }
  // top-k by (exact distance, id) of the m candidates that go to "reordering"; only np2k >= m keys are sorted
  int np2k = 2;
  while ((uint32_t)np2k < m) np2k <<= 1;
  if (np2k > np2) np2k = np2;
  for (int i = tid; i < np2k; i += kFinThreads) {
    uint64_t key = kKeyMax;
    if ((uint32_t)i < m) key = make_key(rec_exact(rec((int)pb[i])), (uint32_t)kb[i]);
    ka[i] = key;
  }
  __syncthreads();
  block_bitonic_sort(ka, np2k);
  const uint32_t kk = min(k, m);
  const float mulr = distance == 0 ? -1.0f : 1.0f;
  for (uint32_t i = tid; i < out_k; i += kFinThreads) {
    uint32_t id = 0;
    float dist = __uint_as_float(0x7FC00000u);
    if (i < kk) { id = (uint32_t)ka[i]; dist = mulr * ord2f((uint32_t)(ka[i] >> 32)); }
    out_idx[(size_t)q * out_k + i] = id;
    out_dist[(size_t)q * out_k + i] = dist;
  }
}

static size_t merge_smem_bytes(int world, int n_cand, uint32_t nover, int* np2_out) {
  int np2 = 2;
  while ((uint32_t)np2 < nover) np2 <<= 1;
  *np2_out = np2;
  const int m = n_cand < (int)nover ? n_cand : (int)nover;
  // ka, kb (u64) + pa, pb (u32) + the lists' keys + hash table [2 np2] + ids [np2] + partners [np2]
  return (size_t)np2 * 24 + (size_t)world * m * 8 + (size_t)np2 * 16;
}

cudaError_t launch_merge_partials(const DevIndex& ix, uint32_t nq, int world, int n_cand,
                                  const uint32_t* ids, const uint64_t* tie, const float* exact,
                                  uint32_t nover, uint32_t npre, uint32_t k, uint32_t* out_idx,
                                  float* out_dist, uint32_t out_k, cudaStream_t s) {
  int np2 = 2;
  const size_t smem = merge_smem_bytes(world, n_cand, nover, &np2);
  cudaError_t e = cudaFuncSetAttribute(merge_partials_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  MergeSrc src{ids, tie, exact, nullptr, nq, n_cand};
  merge_partials_kernel<false><<<nq, kFinThreads, smem, s>>>(ix.distance, ix.disjoint, src, world,
                                                             nover, npre, k, out_idx, out_dist, out_k, np2);
  return cudaGetLastError();
}

cudaError_t launch_merge_records(const DevIndex& ix, uint32_t nq_valid, uint32_t nq_slice, int world, int n_cand,
                                 const uint4* rec, uint32_t nover, uint32_t npre, uint32_t k, uint32_t* out_idx,
                                 float* out_dist, uint32_t out_k, cudaStream_t s) {
  if (nq_valid == 0) return cudaSuccess;
  int np2 = 2;
  const size_t smem = merge_smem_bytes(world, n_cand, nover, &np2);
  cudaError_t e = cudaFuncSetAttribute(merge_partials_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  MergeSrc src{nullptr, nullptr, nullptr, rec, nq_slice, n_cand};
  merge_partials_kernel<true><<<nq_valid, kFinThreads, smem, s>>>(ix.distance, ix.disjoint, src, world,
                                                                  nover, npre, k, out_idx, out_dist, out_k, np2);
  return cudaGetLastError();
}


// Row-sharded brute force: every rank contributes its local top-k as (global id, API-signed
// distance); the global top-k is the k smallest (internal distance, id) keys of the union, which is exactly
// the single-GPU result because the exact re-scoring of a row does not depend on which shard holds it.
__global__ void __launch_bounds__(kFinThreads)
merge_topk_kernel(int distance, uint32_t nq, int world, int k_in, const uint32_t* __restrict__ ids,
                  const float* __restrict__ dists, uint32_t k, uint32_t* out_idx, float* out_dist, uint32_t out_k,
                  int np2, int dedup_ids) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* ka = reinterpret_cast<uint64_t*>(smem);
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x;
  const int total = world * k_in;
  const float sign = distance == 0 ? -1.f : 1.f;  // API distance -> internal distance (scann.cc:364-369)
  for (int i = tid; i < np2; i += kFinThreads) {
    uint64_t key = kKeyMax;
    if (i < total) {
      const size_t o = ((size_t)(i / k_in) * nq + q) * k_in + (i % k_in);
      const float d = dists[o];
      if (d == d) key = make_key(sign * d, ids[o]);  // NaN marks the padding of a short row
    }
    ka[i] = key;
  }
  __syncthreads();
  block_bitonic_sort(ka, np2);
  if (dedup_ids) {
    // light sharded protocol: the two SOAR copies of a datapoint may reach the top-k of two ranks; both carry the
    // same exact distance, so they are neighbours after the sort -- drop the second and sort again
    bool any = false;
    for (int i0 = 0; i0 < np2; i0 += kFinThreads) {
      const int i = i0 + tid;
      const bool dup = i > 0 && i < np2 && ka[i] != kKeyMax && ka[i] == ka[i - 1];
      __syncthreads();
      if (dup) { ka[i] = kKeyMax; any = true; }
      __syncthreads();
    }
    if (__syncthreads_or(any)) block_bitonic_sort(ka, np2);
  }
  for (uint32_t i = tid; i < out_k; i += kFinThreads) {
    uint32_t id = 0;
    float dist = __uint_as_float(0x7FC00000u);
    if (i < k && (int)i < np2 && ka[i] != kKeyMax) { id = (uint32_t)ka[i]; dist = sign * ord2f((uint32_t)(ka[i] >> 32)); }
    out_idx[(size_t)q * out_k + i] = id;
    out_dist[(size_t)q * out_k + i] = dist;
  }
}

cudaError_t launch_merge_topk(int distance, uint32_t nq, int world, int k_in, const uint32_t* ids, const float* dists,
                              uint32_t k, uint32_t* out_idx, float* out_dist, uint32_t out_k, cudaStream_t s,
                              bool dedup_ids) {
  int np2 = 2;
  while (np2 < world * k_in) np2 <<= 1;
  const size_t smem = (size_t)np2 * 8;
  cudaError_t e = cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  merge_topk_kernel<<<nq, kFinThreads, smem, s>>>(distance, nq, world, k_in, ids, dists, k, out_idx, out_dist, out_k, np2,
                                                  dedup_ids ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace sb
