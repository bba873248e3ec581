// finalize.cu -- slot -> datapoint id, SOAR de-duplication, exact reordering, final top-k.
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

__device__ __forceinline__ float exact_distance(const DevIndex& ix, const float* __restrict__ q,
                                                uint32_t dp) {
  const uint32_t row = ix.dp_row ? ix.dp_row[dp] : dp;
  const float* x = ix.dataset + (size_t)row * ix.d;
  auto lq = [&](uint32_t i) { return q[i]; };
  auto lx = [&](uint32_t i) { return __ldg(x + i); };
  if (ix.distance == 0) return ix.d < 8 ? neg_dot_small(lq, lx, ix.d) : neg_dot_avx2_order(lq, lx, ix.d);
  return ix.d < 8 ? sql2_small(lq, lx, ix.d) : sql2_avx2_order(lq, lx, ix.d);
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
  const uint32_t n = min(w.cnt[q], w.nover);
  const uint64_t* src = w.buf + (size_t)q * w.cap;
  for (uint32_t i = tid; i < ix.d; i += kFinThreads) sq[i] = a.q[(size_t)q * ix.d + i];
  if (tid == 0) s_removed = 0;
  uint32_t m;  // candidates that go to reordering
  if (ix.disjoint) {
    // (score, slot) order is already final; keys become (score, dp)
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t k = kKeyMax;
      if ((uint32_t)i < n) {
        const uint64_t s = src[i];
        k = (s & 0xFFFFFFFF00000000ull) | ix.slot_dp[(uint32_t)s];
      }
      kb[i] = k;
    }
    m = min(n, a.npre);
    __syncthreads();
  } else {
    // sort by (dp, score) so the two SOAR copies of a datapoint are adjacent
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t k = kKeyMax;
      if ((uint32_t)i < n) {
        const uint64_t s = src[i];
        k = ((uint64_t)ix.slot_dp[(uint32_t)s] << 32) | (s >> 32);
      }
      ka[i] = k;
    }
    __syncthreads();
    block_bitonic_sort(ka, np2);
    uint32_t removed = 0;
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t k = kKeyMax;
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
        }
      }
      kb[i] = k;
    }
    if (removed) atomicAdd(&s_removed, removed);
    __syncthreads();
    block_bitonic_sort(kb, np2);
    m = min(n - s_removed, a.npre);
  }
  // optional partial output for the sharded path: (id, tie-break key, AH score, exact distance)
  const bool reorder = ix.dataset != nullptr;
  for (int i = tid; i < np2; i += kFinThreads) {
    uint64_t k = kKeyMax;
    if ((uint32_t)i < m) {
      const uint64_t c = kb[i];
      const uint32_t dp = (uint32_t)c;
      const float dist = reorder ? exact_distance(ix, sq, dp) : ord2f((uint32_t)(c >> 32));
      k = make_key(dist, dp);
      if (a.part_ids) {
        const size_t o = (size_t)q * a.part_cap + i;
        a.part_ids[o] = dp;
        a.part_tie[o] = ix.disjoint ? src[i] : c;
        a.part_ah[o] = ord2f((uint32_t)(c >> 32));
        a.part_exact[o] = dist;
      }
    } else if (a.part_ids && (uint32_t)i < a.part_cap) {
      const size_t o = (size_t)q * a.part_cap + i;
      a.part_ids[o] = 0xFFFFFFFFu;
      a.part_tie[o] = kKeyMax;
      a.part_ah[o] = INFINITY;
      a.part_exact[o] = INFINITY;
    }
    ka[i] = k;
  }
  if (a.part_ids) {
    for (uint32_t i = np2 + tid; i < a.part_cap; i += kFinThreads) {
      const size_t o = (size_t)q * a.part_cap + i;
      a.part_ids[o] = 0xFFFFFFFFu;
      a.part_tie[o] = kKeyMax;
      a.part_ah[o] = INFINITY;
      a.part_exact[o] = INFINITY;
    }
  }
  __syncthreads();
  if (!a.out_idx) return;
  block_bitonic_sort(ka, np2);
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
  const size_t smem = (size_t)np2 * 16 + (((size_t)ix.d + 3) & ~(size_t)3) * 4;
  cudaError_t e = cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  finalize_kernel<<<w.nq, kFinThreads, smem, s>>>(ix, w, a, np2);
  return cudaGetLastError();
}

// Merge `world` all-gathered partial lists: global top-N' by (AH score, tie-break key), then
// top-k by (exact distance, id).  SURVEY.md section 8e.
__global__ void __launch_bounds__(kFinThreads)
merge_partials_kernel(int distance, uint32_t nq, int world, int n_cand, const uint32_t* __restrict__ ids,
                      const uint64_t* __restrict__ tie, const float* __restrict__ ah,
                      const float* __restrict__ exact, uint32_t npre, uint32_t k, uint32_t* out_idx,
                      float* out_dist, uint32_t out_k, int np2) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* ka = reinterpret_cast<uint64_t*>(smem);   // [np2] sort keys
  uint32_t* pay = reinterpret_cast<uint32_t*>(ka + np2);  // [np2] payload index
  (void)ah;
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x;
  const int total = world * n_cand;
  // 1) rank all candidates by the tie-break key (score bits in the high half): since a u64
  //    bitonic sort cannot carry a payload, sort (key) and find payloads by a second pass.
  for (int i = tid; i < np2; i += kFinThreads) {
    uint64_t key = kKeyMax;
    if (i < total) {
      const int r = i / n_cand, c = i % n_cand;
      const size_t o = ((size_t)r * nq + q) * n_cand + c;
      if (ids[o] != 0xFFFFFFFFu) key = tie[o];
    }
    ka[i] = key;
  }
  __syncthreads();
  block_bitonic_sort(ka, np2);
  // threshold key = npre-th smallest
  const uint32_t navail = min((uint32_t)total, (uint32_t)np2);
  uint64_t thr = kKeyMax;
  {
    uint32_t cnt_valid = 0;
    // count valid via binary search for first kKeyMax
    int lo = 0, hi = (int)navail;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ka[mid] == kKeyMax) hi = mid; else lo = mid + 1; }
    cnt_valid = (uint32_t)lo;
    const uint32_t m = min(cnt_valid, npre);
    thr = m ? ka[m - 1] : 0;
    if (m == 0) thr = 0;
    __syncthreads();
    // 2) gather the selected ones as (exact distance, id) keys
    for (int i = tid; i < np2; i += kFinThreads) pay[i] = 0;
    __syncthreads();
    for (int i = tid; i < np2; i += kFinThreads) {
      uint64_t key = kKeyMax;
      if (i < total && m) {
        const int r = i / n_cand, c = i % n_cand;
        const size_t o = ((size_t)r * nq + q) * n_cand + c;
        if (ids[o] != 0xFFFFFFFFu && tie[o] <= thr) key = make_key(exact[o], ids[o]);
      }
      ka[i] = key;
    }
    __syncthreads();
    block_bitonic_sort(ka, np2);
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
}

cudaError_t launch_merge_partials(const DevIndex& ix, uint32_t nq, int world, int n_cand,
                                  const uint32_t* ids, const uint64_t* tie, const float* ah,
                                  const float* exact, uint32_t npre, uint32_t k, uint32_t* out_idx,
                                  float* out_dist, uint32_t out_k, cudaStream_t s) {
  int np2 = 2;
  while (np2 < world * n_cand) np2 <<= 1;
  const size_t smem = (size_t)np2 * 12;
  cudaError_t e = cudaFuncSetAttribute(merge_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  merge_partials_kernel<<<nq, kFinThreads, smem, s>>>(ix.distance, nq, world, n_cand, ids, tie, ah, exact,
                                                      npre, k, out_idx, out_dist, out_k, np2);
  return cudaGetLastError();
}

}  // namespace sb
