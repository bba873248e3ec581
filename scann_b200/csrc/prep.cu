// prep.cu -- query preparation kernels: partition tokenization and AH lookup tables.
//
// (a3/a4) KMeansTreePartitioner::TokensForDatapointWithSpillingBatched
//         partitioning/kmeans_tree_partitioner.cc:642-730, many_to_many_impl.inc:522-567
// (a5)    AsymmetricQueryer::CreateLookupTable hashes/asymmetric_hashing2/querying.h:284-329,
//         hashes/internal/asymmetric_hashing_impl.cc:505-645
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"
#include "lut.cuh"

namespace sb {

// ---------------------------------------------------------------------------------------
// Tokenization: dist[q][l] = sequential-in-dim fnmadd chain (bit-identical to the CPU
// kernel's FMA order), as an fp32 SIMT tile GEMM.  64x64 output tile, 4x4 per thread.
// The k loop runs in ascending dim order for every accumulator, so tiling over k does not
// change a single rounding.
// ---------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16, TPAD = 4;

__global__ void __launch_bounds__(256)
tokenize_kernel(const float* __restrict__ q, const float* __restrict__ c,
                const float* __restrict__ cnorm, float* __restrict__ out, int nq, int L, int D,
                int sql2) {
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];
  __shared__ float qn[TM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4];
  if (sql2) {
    // acc = ||c||^2 + ||q||^2 (many_to_many_impl.inc:530-541); ||q||^2 = float(SquaredL2Norm(q)) (:417-426)
    if (tid < TM) {
      const int r = m0 + tid;
      qn[tid] = r < nq ? squared_l2_norm_strided(q + (size_t)r * D, (uint32_t)D) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 4 + j;
        acc[i][j] = __fadd_rn(col < L ? cnorm[col] : 0.f, qn[ty * 4 + i]);
      }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const float bscale = sql2 ? 2.0f : 1.0f;
  for (int k0 = 0; k0 < D; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 4, kk = idx & 15;
      const int gr = m0 + row, gk = k0 + kk;
      As[kk][row] = (gr < nq && gk < D) ? -q[(size_t)gr * D + gk] : 0.f;
      const int gc = n0 + row;
      Bs[kk][row] = (gc < L && gk < D) ? __fmul_rn(c[(size_t)gc * D + gk], bscale) : 0.f;
    }
    __syncthreads();
    const int kmax = min(TK, D - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= nq) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < L) out[(size_t)r * L + col] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Int8 (fixed point) tokenization: query_tokenization_type FIXED_POINT_INT8.  The reference leaves its batched GEMM
// path for this type (KMeansTreePartitioner::SupportsLowLevelQueryBatching, partitioning/kmeans_tree_partitioner.h:
// 230-236) and runs KMeansTreeNode::GetAllDistancesInt8 per query (trees/kmeans_tree/kmeans_tree_node.h:222-256):
//   q'[k] = q[k] * inv_mult[k]  (squared L2: q[k] * (inv_mult[k] * 2))
//   val[l] = DenseDotProductDistanceOneToManyInt8Float(q', int8 centres)[l]
//   dot product: val[l];  squared L2: val[l] + (float(SquaredL2Norm(q)) + float(SquaredL2Norm(float centre l)))
// val is OneToManyAsymmetricTemplate<dims, 3, .., int8_t> (one_to_many_asymmetric_impl.inc:531-671): centres
// [0, 3 (L / 3)) in the three-at-a-time AVX2 order -- eight fnmadd lanes over whole groups of 8 dims, one 4-wide step
// into lanes 0..3, ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)), the remaining dims on the scalar -- and the last L mod 3
// centres through the one-to-one kernel (tokenize_i8_tail_kernel).  Here: a 64-query x 32-centre SIMT tile, four
// queries x two centres per thread, the eight lane accumulators of every (query, centre) pair in registers, so each
// lane sees its dims in the reference's order whatever the tiling over k.  These exact kernels serve small trees; from
// 256 centres launch_tokenize_topp takes the tensor-core route (the int8 centres are an exact bf16 operand) and only
// re-scores the candidates of the top P with the same chains (i8_center_distance below).
// ---------------------------------------------------------------------------------------
constexpr int I8M = 64, I8N = 32, I8K = 16, I8PAD = 4;

// q'[k] = q[k] * scale[k], scale = inv_mult (dot product) or inv_mult * 2 (squared L2: `q *= inv_mult * 2`,
// kmeans_tree_node.h:239-241); the index stores the scale (DevIndex::cen_qscale)
__device__ __forceinline__ float i8_scaled_query(float q, float scale) { return __fmul_rn(q, scale); }

__global__ void __launch_bounds__(256)
tokenize_i8_kernel(const float* __restrict__ q, const float* __restrict__ qscale, const int8_t* __restrict__ c,
                   const float* __restrict__ csq, float* __restrict__ out, int nq, int L, int D, int sql2) {
  __shared__ __align__(16) float As[I8K][I8M + I8PAD];  // -q'
  __shared__ __align__(16) float Bs[I8K][I8N + I8PAD];  // float(centre)
  __shared__ float qn[I8M];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * I8M, n0 = blockIdx.x * I8N;
  if (sql2 && tid < I8M) {
    const int r = m0 + tid;
    qn[tid] = r < nq ? squared_l2_norm_strided(q + (size_t)r * D, (uint32_t)D) : 0.f;
  }
  float acc[4][2][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int l = 0; l < 8; ++l) acc[i][j][l] = 0.f;
  const int D8 = D & ~7;
  // stages dims [k0, k0 + I8K) that are < klim (zeros elsewhere)
  auto stage = [&](int k0, int klim) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 4, kk = idx & 15;
      const int gr = m0 + row, gk = k0 + kk;
      As[kk][row] = (gr < nq && gk < klim) ? -i8_scaled_query(q[(size_t)gr * D + gk], qscale[gk]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 4, kk = idx & 15;
      const int gc = n0 + row, gk = k0 + kk;
      Bs[kk][row] = (gc < L && gk < klim) ? (float)c[(size_t)gc * D + gk] : 0.f;
    }
  };
  for (int k0 = 0; k0 < D8; k0 += I8K) {
    stage(k0, D8);
    __syncthreads();
    const int groups = min(2, (D8 - k0) >> 3);
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const float4 a = *reinterpret_cast<const float4*>(&As[g * 8 + l][ty * 4]);
        const float2 b = *reinterpret_cast<const float2*>(&Bs[g * 8 + l][tx * 2]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[2] = {b.x, b.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) acc[i][j][l] = __fmaf_rn(av[i], bv[j], acc[i][j][l]);
      }
    }
    __syncthreads();
  }
  // the last D mod 8 dims: a 4-wide step into lanes 0..3 when at least four are left, the rest after the lane sum
  const int rem = D - D8;
  if (rem) stage(D8, D);
  __syncthreads();  // also orders qn[]
  int t = 0;
  if (rem >= 4) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float4 a = *reinterpret_cast<const float4*>(&As[l][ty * 4]);
      const float2 b = *reinterpret_cast<const float2*>(&Bs[l][tx * 2]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[2] = {b.x, b.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][l] = __fmaf_rn(av[i], bv[j], acc[i][j][l]);
    }
    t = 4;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float* a = acc[i][j];
      float v = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[4]), __fadd_rn(a[2], a[6])),
                          __fadd_rn(__fadd_rn(a[1], a[5]), __fadd_rn(a[3], a[7])));
      for (int kk = t; kk < rem; ++kk) v = __fmaf_rn(As[kk][ty * 4 + i], Bs[kk][tx * 2 + j], v);
      const int col = n0 + tx * 2 + j;
      if (r < nq && col < L) out[(size_t)r * L + col] = sql2 ? __fadd_rn(v, __fadd_rn(qn[ty * 4 + i], csq[col])) : v;
    }
  }
}

// The last L mod 3 centres: ComputeOneToOneScore<0, false> = -DenseDotProductInt8FloatAvxImpl (one_to_many_asymmetric_
// impl.inc:629-639,223-258; distance_measures/one_to_one/dot_product_impl.inc:3-54): two 8-lane fmadd accumulators over
// whole groups of 16 dims, one 8-wide fmadd step into the first, one 4-wide step as a rounded product ADDED to lanes
// 0..3 of the first, Sum8(acc0 + acc1) = ((x0+x4)+(x2+x6)) + ((x1+x5)+(x3+x7)), the last < 4 dims fused on the scalar.
// One thread per (query, tail centre).
__global__ void __launch_bounds__(128)
tokenize_i8_tail_kernel(const float* __restrict__ q, const float* __restrict__ qscale, const int8_t* __restrict__ c,
                        const float* __restrict__ csq, float* __restrict__ out, int nq, int L, int D, int sql2, int L3) {
  const int nt = L - L3;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nq * nt) return;
  const int r = g / nt, col = L3 + g % nt;
  const float* qr = q + (size_t)r * D;
  auto qp = [&](uint32_t k) { return i8_scaled_query(qr[k], qscale[k]); };
  const float val = neg_dot_i8_one_to_one(qp, c + (size_t)col * D, (uint32_t)D);
  out[(size_t)r * L + col] = sql2 ? __fadd_rn(val, __fadd_rn(squared_l2_norm_strided(qr, (uint32_t)D), csq[col])) : val;
}

void launch_tokenize(const DevIndex& ix, const float* q, uint32_t nq, float* dist, cudaStream_t s) {
  if (ix.centers_i8) {
    const int sql2 = ix.distance == 1;
    dim3 grid((ix.L + I8N - 1) / I8N, (nq + I8M - 1) / I8M);
    tokenize_i8_kernel<<<grid, 256, 0, s>>>(q, ix.cen_qscale, ix.centers_i8, ix.cen_sqnorm, dist, (int)nq, (int)ix.L,
                                            (int)ix.d, sql2);
    const int L3 = (int)(ix.L / 3 * 3), nt = (int)ix.L - L3;
    if (nt)
      tokenize_i8_tail_kernel<<<((int)nq * nt + 127) / 128, 128, 0, s>>>(q, ix.cen_qscale, ix.centers_i8, ix.cen_sqnorm,
                                                                         dist, (int)nq, (int)ix.L, (int)ix.d, sql2, L3);
    return;
  }
  dim3 grid((ix.L + TN - 1) / TN, (nq + TM - 1) / TM);
  tokenize_kernel<<<grid, 256, 0, s>>>(q, ix.centers, ix.center_sqnorm, dist, (int)nq, (int)ix.L,
                                       (int)ix.d, ix.distance == 1);
}

// ---------------------------------------------------------------------------------------
// Top-P leaves per query: exact P smallest (distance, leaf) keys, sorted ascending.
// FastTopNeighbors semantics (utils/fast_top_neighbors_impl.inc:345-374): ties at the cut
// go to the smaller index.  MSB radix select on the order-preserving u32 image of the
// distance (4 x 8-bit digits), then an index-ordered pass that takes everything below the
// pivot value and the first `need` elements equal to it, then a bitonic sort of the P keys.
// ---------------------------------------------------------------------------------------
constexpr int kToppThreads = 256;

__global__ void __launch_bounds__(kToppThreads)
topp_kernel(const float* __restrict__ dist, int L, int P, int Ppow2, int32_t* __restrict__ leaves,
            float* __restrict__ bias) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_need, s_count, s_base_eq;
  __shared__ uint32_t warp_sums[kToppThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = dist + (size_t)blockIdx.x * L;

  uint32_t prefix = 0, mask = 0, need = (uint32_t)P;
  if (P < L) {
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < L; i += kToppThreads) {
        const uint32_t o = f2ord(row[i]);
        if ((o & mask) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        uint32_t cum = 0;
        int dsel = 255;
        for (int dgt = 0; dgt < 256; ++dgt) {
          if (cum + hist[dgt] >= need) { dsel = dgt; break; }
          cum += hist[dgt];
        }
        s_prefix = prefix | ((uint32_t)dsel << shift);
        s_need = need - cum;
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      mask |= 0xFFu << shift;
      __syncthreads();
    }
  } else {
    prefix = 0xFFFFFFFFu;  // everything is "less or equal"; take all
    need = 0xFFFFFFFFu;
  }
  if (tid == 0) { s_count = 0; s_base_eq = 0; }
  for (int i = tid; i < Ppow2; i += kToppThreads) skeys[i] = kKeyMax;
  __syncthreads();
  for (int t0 = 0; t0 < L; t0 += kToppThreads) {
    const int i = t0 + tid;
    const bool valid = i < L;
    const uint32_t o = valid ? f2ord(row[i]) : 0xFFFFFFFFu;
    const bool less = valid && o < prefix;
    const bool eq = valid && o == prefix;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, eq);
    if (lane == 0) warp_sums[warp] = __popc(m);
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kToppThreads / 32; ++w) {
      const uint32_t v = warp_sums[w];
      if (w < warp) woff += v;
      total += v;
    }
    const uint32_t rank = s_base_eq + woff + __popc(m & ((1u << lane) - 1u));
    if (less || (eq && rank < need)) {
      const uint32_t pos = atomicAdd(&s_count, 1u);
      if (pos < (uint32_t)Ppow2) skeys[pos] = ((uint64_t)o << 32) | (uint32_t)i;
    }
    __syncthreads();
    if (tid == 0) s_base_eq += total;
    __syncthreads();
  }
  block_bitonic_sort(skeys, Ppow2);
  for (int i = tid; i < P; i += kToppThreads) {
    const uint64_t k = skeys[i];
    const uint32_t l = (uint32_t)k;
    leaves[(size_t)blockIdx.x * P + i] = (k == kKeyMax) ? -1 : (int32_t)l;
    bias[(size_t)blockIdx.x * P + i] = (k == kKeyMax) ? 0.f : row[l];
  }
}

// Warp-per-query variant of the same selection (no block barriers): used when P <= 256, which
// covers every configuration of BASELINE.json.  8 queries per 256-thread block.
constexpr int kToppWarpMaxP = 256;

// MSB radix select by one warp: returns the order-image `prefix` of the P-th smallest value of
// ordfn(0..L-1) and, in `need`, how many elements equal to it belong to the P smallest.
template <typename F>
__device__ __forceinline__ uint32_t warp_radix_select(F ordfn, int L, uint32_t P, uint32_t* hist, int lane,
                                                      uint32_t& need_out) {
  uint32_t prefix = 0, mask = 0, need = P;
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[lane * 8 + k] = 0;
    __syncwarp();
    for (int i = lane; i < L; i += 32) {
      const uint32_t o = ordfn(i);
      if ((o & mask) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
    }
    __syncwarp();
    uint32_t c[8], tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; tot += c[k]; }
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t hit = __ballot_sync(0xFFFFFFFFu, incl >= need);
    const int tl = hit ? (__ffs(hit) - 1) : 31;
    uint32_t digit = 255, nneed = need;
    if (lane == tl) {
      uint32_t cum = incl - tot;
      digit = (uint32_t)lane * 8 + 7;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (cum + c[k] >= need) { digit = (uint32_t)lane * 8 + k; break; }
        cum += c[k];
      }
      nneed = need - cum;
    }
    digit = __shfl_sync(0xFFFFFFFFu, digit, tl);
    need = __shfl_sync(0xFFFFFFFFu, nneed, tl);
    prefix |= digit << shift;
    mask |= 0xFFu << shift;
    __syncwarp();
  }
  need_out = need;
  return prefix;
}

__device__ __forceinline__ void warp_bitonic_sort(uint64_t* skeys, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        const int l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int r = l | j;
        const uint64_t a = skeys[l], b = skeys[r];
        const bool up = (l & k) == 0;
        if ((a > b) == up) { skeys[l] = b; skeys[r] = a; }
      }
      __syncwarp();
    }
  }
}

// Exact top-P of one row of distances by one warp; writes P sorted (leaf, distance) pairs.
__device__ __forceinline__ void warp_topp_exact(const float* row, int L, int P, int Ppow2, uint32_t* hist,
                                                uint64_t* skeys, int lane, int32_t* leaves, float* bias) {
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t prefix = 0xFFFFFFFFu, need = 0xFFFFFFFFu;
  if (P < L) prefix = warp_radix_select([&](int i) { return f2ord(row[i]); }, L, (uint32_t)P, hist, lane, need);
  for (int i = lane; i < Ppow2; i += 32) skeys[i] = kKeyMax;
  __syncwarp();
  uint32_t count = 0, base_eq = 0;
  for (int t0 = 0; t0 < L; t0 += 32) {
    const int i = t0 + lane;
    const bool valid = i < L;
    const uint32_t o = valid ? f2ord(row[i]) : 0xFFFFFFFFu;
    const bool less = valid && o < prefix;
    const bool eq = valid && o == prefix;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, eq);
    const uint32_t rank = base_eq + __popc(m & lt);
    const bool take = less || (eq && rank < need);
    const uint32_t tm = __ballot_sync(0xFFFFFFFFu, take);
    if (take) {
      const uint32_t pos = count + __popc(tm & lt);
      if (pos < (uint32_t)Ppow2) skeys[pos] = ((uint64_t)o << 32) | (uint32_t)i;
    }
    count += __popc(tm);
    base_eq += __popc(m);
  }
  __syncwarp();
  warp_bitonic_sort(skeys, Ppow2, lane);
  for (int i = lane; i < P; i += 32) {
    const uint64_t k = skeys[i];
    const uint32_t l = (uint32_t)k;
    leaves[i] = (k == kKeyMax) ? -1 : (int32_t)l;
    bias[i] = (k == kKeyMax) ? 0.f : row[l];
  }
}

__global__ void __launch_bounds__(256)
topp_warp_kernel(const float* __restrict__ dist, int nq, int L, int P, int Ppow2,
                 int32_t* __restrict__ leaves, float* __restrict__ bias) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 8 + warp;
  if (q >= nq) return;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw) + warp * 256;
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + 8 * 256 * 4) + (size_t)warp * Ppow2;
  warp_topp_exact(dist + (size_t)q * L, L, P, Ppow2, hist, skeys, lane, leaves + (size_t)q * P, bias + (size_t)q * P);
}

void launch_topp(const DevIndex& ix, const float* dist, uint32_t nq, uint32_t P, int32_t* leaves,
                 float* bias, cudaStream_t s) {
  int pp = 1;
  while (pp < (int)P) pp <<= 1;
  if (pp < 2) pp = 2;
  if (pp <= kToppWarpMaxP) {
    const size_t smem = 8 * 256 * 4 + (size_t)8 * pp * 8;
    topp_warp_kernel<<<(nq + 7) / 8, 256, smem, s>>>(dist, (int)nq, (int)ix.L, (int)P, pp, leaves, bias);
  } else {
    topp_kernel<<<nq, kToppThreads, (size_t)pp * 8, s>>>(dist, (int)ix.L, (int)P, pp, leaves, bias);
  }
}

// ---------------------------------------------------------------------------------------
// Tokenization on the tensor cores (north-star item 1): the query x centre contraction runs as a
// bf16 tcgen05 GEMM (bruteforce.cu) and only the few centres that can belong to the top P are
// re-scored with the reference's exact fp32 FMA chain, so the result is bit-identical to the SIMT
// path above.
//   * q = qh + ql (+ O(2^-18)), c = ch + cl (+ O(2^-18)) in bf16; operands are concatenated along
//     K as A = [qh | ql | qh], B = [ch | ch | cl], so one plain GEMM yields qh.ch + ql.ch + qh.cl.
//   * |approx - chain| <= eps(q) := (K * 2^-21 + 2^-15) * |q| * max|c| (dropped terms <= 4 * 2^-18,
//     fp32 accumulation of K exact products <= K * 2^-22, the chain's own rounding <= D * 2^-24, all
//     relative to sum |q_d c_d| <= |q| |c|), plus the fp32 roundings of |q|^2 + |c|^2 - 2S for L2.
//   * With T = the P-th smallest approximate distance, every centre of the exact top P has an
//     approximate distance <= T + 2 eps (otherwise P centres would be strictly closer).  Those
//     candidates (P plus a handful) get the exact chain and are sorted by (distance, leaf).
//   * If a query has more candidates than the warp's buffer (degenerate inputs: zero query, equal
//     centres), the warp falls back to exact distances for all L centres.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
split_rows_kernel(const float* __restrict__ src, uint32_t rows, uint32_t d, uint32_t kp, int lo_term,
                  __nv_bfloat16* __restrict__ out, const float* __restrict__ scale) {
  const uint32_t r = blockIdx.x;
  __nv_bfloat16* o = out + (size_t)r * kp;
  for (uint32_t k = threadIdx.x; k < d; k += 128) {
    float x = r < rows ? src[(size_t)r * d + k] : 0.f;
    if (scale) x = __fmul_rn(x, scale[k]);  // int8 tokenization: the scaled query q' (the same rounding as the exact chain's)
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    o[k] = hi;
    o[d + k] = lo_term == 1 ? lo : hi;
    o[2 * d + k] = lo_term == 2 ? lo : hi;
  }
  for (uint32_t k = 3 * d + threadIdx.x; k < kp; k += 128) o[k] = __float2bfloat16_rn(0.f);
}

uint32_t tokenize_kpitch(uint32_t d) { return (3 * d + 63) / 64 * 64; }
size_t tokenize_operand_bytes(uint32_t rows, uint32_t d) {
  return (size_t)((rows + 127) / 128 * 128) * tokenize_kpitch(d) * 2;
}
cudaError_t build_tokenize_operand(const float* src, uint32_t rows, uint32_t d, int lo_term, void* out, cudaStream_t s,
                                   const float* scale) {
  const uint32_t kp = tokenize_kpitch(d), rows_pad = (rows + 127) / 128 * 128;
  split_rows_kernel<<<rows_pad, 128, 0, s>>>(src, rows, d, kp, lo_term, reinterpret_cast<__nv_bfloat16*>(out), scale);
  return cudaGetLastError();
}

// The exact distance of int8 tokenization for one centre, as the refinement kernels re-score it: `sq` = the scaled
// query q' (shared memory), the reference's three-at-a-time order for centres below L3 = 3 (L / 3), its one-to-one order
// for the last L mod 3 (exact_math.cuh; tokenize_i8_kernel / tokenize_i8_tail_kernel compute the same bits);
// squared L2: val + (|q|^2 + |float centre|^2).
__device__ __forceinline__ float i8_center_distance(const DevIndex& ix, const float* __restrict__ sq, int idx, int D, int L3,
                                                    bool sql2, float qn) {
  const int8_t* __restrict__ c = ix.centers_i8 + (size_t)idx * D;
  auto lsq = [&](uint32_t k) { return sq[k]; };
  float val;
  if (idx >= L3) {
    val = neg_dot_i8_one_to_one(lsq, c, (uint32_t)D);
  } else if ((D & 3) == 0) {
    // rows are 4-byte aligned: one 32-bit load per four dims (the lanes and their order are neg_dot_asym_order's)
    const uint32_t* __restrict__ cw = reinterpret_cast<const uint32_t*>(c);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int j = 0;
    for (; j + 8 <= D; j += 8) {
      const uint32_t w0 = __ldg(cw + (j >> 2)), w1 = __ldg(cw + (j >> 2) + 1);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        a[l] = __fmaf_rn(-sq[j + l], (float)(int8_t)(w0 >> (8 * l)), a[l]);
        a[l + 4] = __fmaf_rn(-sq[j + 4 + l], (float)(int8_t)(w1 >> (8 * l)), a[l + 4]);
      }
    }
    if (j + 4 <= D) {
      const uint32_t w0 = __ldg(cw + (j >> 2));
#pragma unroll
      for (int l = 0; l < 4; ++l) a[l] = __fmaf_rn(-sq[j + l], (float)(int8_t)(w0 >> (8 * l)), a[l]);
    }
    val = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[4]), __fadd_rn(a[2], a[6])),
                    __fadd_rn(__fadd_rn(a[1], a[5]), __fadd_rn(a[3], a[7])));
  } else {
    auto lc = [&](uint32_t k) { return (float)c[k]; };
    val = neg_dot_asym_order(lsq, lc, (uint32_t)D);
  }
  return sql2 ? __fadd_rn(val, __fadd_rn(qn, ix.cen_sqnorm[idx])) : val;
}

constexpr int kRefineMaxCand = 512;

// One 128-thread CTA per query.  The row of approximate distances is mapped to a fixed-point image
//   u = min(uint((a - min) * 2^31 * (1 - 2^-10) / (max - min)), 2^31 - 1),
// which is monotone in the distance and spreads the row evenly over the 256 bins of the first
// radix pass (the leading sign/exponent bits of the float image are common to nearly all distances
// of a query; a histogram over them only serialises shared-memory atomic conflicts).  The MSB radix
// select on u stops as soon as everything up to the end of the pivot's bucket fits the candidate
// buffer: an upper bound hi_u of the P-th smallest is all that is needed.  Candidates are the
// centres with u <= hi_u + ceil(2 eps * scale) + 1024; the 1024 covers the fp32 roundings of the
// map (2 ulp of a value < 2^31), so this is a superset of {a <= a_P + 2 eps}.
// kSmemRow: u is staged in shared memory once (L <= kRefineSmemL, L % 4 == 0), so the passes touch
// no global memory; otherwise any L, u recomputed from the row on every pass.
constexpr int kRefineSmemL = 49152;  // 192 KB of shared memory for the row image

template <bool kSmemRow, int kRefineThreads>
__global__ void __launch_bounds__(kRefineThreads)
topp_refine_kernel(DevIndex ix, const float* __restrict__ q, float* __restrict__ S, int nq, int P, int Ppow2, int Cp,
                   float eps_rel, int32_t* __restrict__ leaves, float* __restrict__ bias, uint32_t* fallbacks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t hist[256];
  constexpr int kWarps = kRefineThreads / 32;
  __shared__ float red_a[kWarps], red_b[kWarps];
  __shared__ uint32_t s_sel[4];  // digit, cum, bucket, candidate counter
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qi = blockIdx.x;
  const int L = (int)ix.L, D = (int)ix.d, Dp = (D + 3) & ~3;
  const bool sql2 = ix.distance == 1;
  // int8 tokenization (see tokenize_i8_kernel): S = <q', float(int8 centre)>, the exact function is the int8 chain
  const bool i8 = ix.centers_i8 != nullptr;
  const float* __restrict__ cnp = i8 ? ix.cen_sqnorm : ix.center_sqnorm;
  const int Lp = (L + 4 * kRefineThreads - 1) / (4 * kRefineThreads) * (4 * kRefineThreads);  // whole LDS.128 iterations of the block
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);
  float* sq = reinterpret_cast<float*>(smem_raw + (size_t)Cp * 8);
  uint32_t* srow = reinterpret_cast<uint32_t*>(smem_raw + (size_t)Cp * 8 + (size_t)Dp * 4);
  float* row = S + (size_t)qi * L;

  float ssq = 0.f;
  for (int k = tid; k < D; k += kRefineThreads) {
    float v = q[(size_t)qi * D + k];
    if (i8) v = __fmul_rn(v, ix.cen_qscale[k]);  // q' (sq holds the scaled query; ||q||^2 below reads the raw one)
    sq[k] = v;
    ssq = fmaf(v, v, ssq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xFFFFFFFFu, ssq, o);
  if (lane == 0) red_a[warp] = ssq;
  __syncthreads();
  ssq = 0.f;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) ssq += red_a[i];
  float qn = 0.f;
  if (sql2) {  // ||q||^2 exactly as tokenize_kernel: float(SquaredL2Norm(q))
    if (tid == 0) red_b[0] = squared_l2_norm_strided(i8 ? q + (size_t)qi * D : sq, (uint32_t)D);
    __syncthreads();
    qn = red_b[0];
  }
  __syncthreads();
  const float qnorm = sqrtf(ssq) * 1.001f, cmax = i8 ? ix.cen_i8_max_norm : ix.center_max_norm;
  float eps = eps_rel * qnorm * cmax;
  if (sql2 && !i8) eps = 2.f * eps + (float)(D + 8) * 1.1920929e-7f * (qn + cmax * cmax + 2.f * qnorm * cmax);
  // int8, squared L2: approx = (cn + qn) - S and exact = val + (qn + cn) share the rounded (qn + cn); what differs is
  // S against -val (eps: q' carries the factor 2 already) and the two final roundings, each <= 2^-24 of the sum
  if (sql2 && i8) eps = eps + 8.f * 1.1920929e-7f * (qn + ix.cen_sqnorm_max + qnorm * cmax);
  auto approx = [&](float sdot, float cn) -> float {
    return sql2 ? __fsub_rn(__fadd_rn(cn, qn), i8 ? sdot : __fmul_rn(2.f, sdot)) : -sdot;
  };

  // ---- pass A: approximate distances (to shared memory), their min and max ----
  float amin = __int_as_float(0x7F800000), amax = __int_as_float(0xFF800000);
  if constexpr (kSmemRow) {
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const float4* cn4 = reinterpret_cast<const float4*>(cnp);
#pragma unroll 4
    for (int i4 = tid; i4 < (Lp >> 2); i4 += kRefineThreads) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i4 < (L >> 2)) {
        const float4 v = row4[i4];
        float4 cn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sql2) cn = cn4[i4];
        a = make_float4(approx(v.x, cn.x), approx(v.y, cn.y), approx(v.z, cn.z), approx(v.w, cn.w));
        amin = fminf(fminf(amin, a.x), fminf(a.y, fminf(a.z, a.w)));
        amax = fmaxf(fmaxf(amax, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
      }
      reinterpret_cast<float4*>(srow)[i4] = a;
    }
  } else {
#pragma unroll 4
    for (int i = tid; i < L; i += kRefineThreads) {
      const float a = approx(row[i], sql2 ? cnp[i] : 0.f);
      amin = fminf(amin, a);
      amax = fmaxf(amax, a);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amin = fminf(amin, __shfl_xor_sync(0xFFFFFFFFu, amin, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
  }
  if (lane == 0) { red_a[warp] = amin; red_b[warp] = amax; }
  __syncthreads();
  amin = red_a[0];
  amax = red_b[0];
#pragma unroll
  for (int i = 1; i < kWarps; ++i) { amin = fminf(amin, red_a[i]); amax = fmaxf(amax, red_b[i]); }
  const float span = __fsub_rn(amax, amin);
  const float scale = (span > 0.f && span < __int_as_float(0x7F800000)) ? __fdiv_rn(2145386496.f, span) : 0.f;
  auto ufn = [&](float a) -> uint32_t {
    return min(__float2uint_rz(__fmul_rn(__fsub_rn(a, amin), scale)), 0x7FFFFFFFu);
  };
  // fn(i, u) over the row in block-uniform trip counts; slots past L carry u = 0xFFFFFFFF
  auto for_each_u = [&](auto&& fn) {
    if constexpr (kSmemRow) {
      for (int i4 = tid; i4 < (Lp >> 2); i4 += kRefineThreads) {
        const uint4 u = reinterpret_cast<const uint4*>(srow)[i4];
        fn(i4 * 4, u.x); fn(i4 * 4 + 1, u.y); fn(i4 * 4 + 2, u.z); fn(i4 * 4 + 3, u.w);
      }
    } else {
      for (int t0 = 0; t0 < L; t0 += kRefineThreads) {
        const int i = t0 + tid;
        fn(i, i < L ? ufn(approx(row[i], sql2 ? cnp[i] : 0.f)) : 0xFFFFFFFFu);
      }
    }
  };

  uint32_t lim_u = 0x7FFFFFFFu;
  if (P < L) {
    uint32_t mask = 0, prefix = 0, need = (uint32_t)P, below = 0;
    for (int shift = 23;; ) {
      for (int i = tid; i < 256; i += kRefineThreads) hist[i] = 0;
      __syncthreads();
      if (kSmemRow && shift == 23) {
        // first pass: convert the staged distances to u in place (each thread revisits its own slots)
        for (int i4 = tid; i4 < (Lp >> 2); i4 += kRefineThreads) {
          const float4 a = reinterpret_cast<const float4*>(srow)[i4];
          uint4 u = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
          if (i4 < (L >> 2)) {
            u = make_uint4(ufn(a.x), ufn(a.y), ufn(a.z), ufn(a.w));
            atomicAdd(&hist[u.x >> 23], 1u); atomicAdd(&hist[u.y >> 23], 1u);
            atomicAdd(&hist[u.z >> 23], 1u); atomicAdd(&hist[u.w >> 23], 1u);
          }
          reinterpret_cast<uint4*>(srow)[i4] = u;
        }
      } else {
        for_each_u([&](int, uint32_t u) {
          if (u <= 0x7FFFFFFFu && (u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        });
      }
      __syncthreads();
      if (warp == 0) {
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; tot += c[k]; }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= o) incl += t;
        }
        const uint32_t hit = __ballot_sync(0xFFFFFFFFu, incl >= need);
        const int tl = hit ? (__ffs(hit) - 1) : 31;
        if (lane == tl) {
          uint32_t cum = incl - tot, digit = (uint32_t)lane * 8 + 7, bucket = c[7];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (cum + c[k] >= need) { digit = (uint32_t)lane * 8 + k; bucket = c[k]; break; }
            cum += c[k];
          }
          s_sel[0] = digit; s_sel[1] = cum; s_sel[2] = bucket;
        }
      }
      __syncthreads();
      const uint32_t digit = s_sel[0], cum = s_sel[1], bucket = s_sel[2];
      prefix |= digit << shift;
      mask |= 0xFFu << shift;
      below += cum;
      need -= cum;
      if (shift == 0 || below + bucket + 16 <= (uint32_t)Cp) break;
      shift = shift > 8 ? shift - 8 : 0;
    }
    const uint32_t hi_u = prefix | (~mask & 0x7FFFFFFFu);
    const float widen = __fmul_ru(__fmul_ru(2.f * eps, scale), 1.000001f);
    const uint32_t extra = widen < 2.0e9f ? __float2uint_ru(widen) + 1024u : 0x7FFFFFFFu;  // NaN -> everything
    lim_u = hi_u + min(extra, 0x7FFFFFFFu - hi_u);
  } else if constexpr (kSmemRow) {
    // every centre is a candidate: only mark the valid slots
    for (int i4 = tid; i4 < (Lp >> 2); i4 += kRefineThreads) {
      const uint32_t f = i4 < (L >> 2) ? 0u : 0xFFFFFFFFu;
      reinterpret_cast<uint4*>(srow)[i4] = make_uint4(f, f, f, f);
    }
  }
  if (tid == 0) s_sel[3] = 0;
  __syncthreads();
  // ~P of L slots qualify: a per-thread reservation is cheaper than warp-aggregating every slot
  for_each_u([&](int i, uint32_t u) {
    if (u <= lim_u) {
      const uint32_t pos = atomicAdd(&s_sel[3], 1u);
      if (pos < (uint32_t)Cp) skeys[pos] = (uint64_t)(uint32_t)i;
    }
  });
  __syncthreads();
  const uint32_t count = s_sel[3];

  // int8 tokenization: the reference's int8 chain -- the three-at-a-time order for centres below 3 (L / 3), the
  // one-to-one order for the last L mod 3 (exact_math.cuh; tokenize_i8_kernel / tokenize_i8_tail_kernel compute the same)
  const int L3 = L / 3 * 3;
  auto exact_i8 = [&](int idx) -> float { return i8_center_distance(ix, sq, idx, D, L3, sql2, qn); };
  // the reference's exact fp32 chain for one centre (scalar loads: fallback and D % 4 != 0)
  auto exact = [&](int idx) -> float {
    if (i8) return exact_i8(idx);
    const float* c = ix.centers + (size_t)idx * D;
    float acc = sql2 ? __fadd_rn(ix.center_sqnorm[idx], qn) : 0.f;
    const float scale2 = sql2 ? 2.0f : 1.0f;
#pragma unroll 4
    for (int k = 0; k < D; ++k) acc = __fmaf_rn(-sq[k], __fmul_rn(__ldg(c + k), scale2), acc);
    return acc;
  };
  // the same chain with 16-byte loads, 16 dims in flight before the first FMA of a chunk (D % 4 == 0)
  auto exact_v4 = [&](int idx) -> float {
    const float* c = ix.centers + (size_t)idx * D;
    float acc = sql2 ? __fadd_rn(ix.center_sqnorm[idx], qn) : 0.f;
    for (int k0 = 0; k0 < D; k0 += 16) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(c + min(k0 + 4 * u, D - 4)));
      if (sql2) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          v[u] = make_float4(__fmul_rn(v[u].x, 2.f), __fmul_rn(v[u].y, 2.f), __fmul_rn(v[u].z, 2.f), __fmul_rn(v[u].w, 2.f));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 4 * u;
        if (k < D) {
          const float4 qv = *reinterpret_cast<const float4*>(sq + k);
          acc = __fmaf_rn(-qv.x, v[u].x, acc);
          acc = __fmaf_rn(-qv.y, v[u].y, acc);
          acc = __fmaf_rn(-qv.z, v[u].z, acc);
          acc = __fmaf_rn(-qv.w, v[u].w, acc);
        }
      }
    }
    return acc;
  };

  int32_t* lout = leaves + (size_t)qi * P;
  float* bout = bias + (size_t)qi * P;
  if (count <= (uint32_t)Cp) {
    int ns = 2;
    while ((uint32_t)ns < count) ns <<= 1;
    for (int j0 = warp * 32; j0 < ns; j0 += kRefineThreads) {  // warp-uniform trip count
      const int j = j0 + lane;
      const bool live = (uint32_t)j < count;
      const uint32_t idx = live ? (uint32_t)skeys[j] : 0u;
      float e = 0.f;
      if (live) e = ((D & 3) == 0 && !i8) ? exact_v4((int)idx) : exact((int)idx);
      if (j < ns) skeys[j] = live ? (((uint64_t)f2ord(e) << 32) | idx) : kKeyMax;
    }
    __syncthreads();
    if (ns <= 128) {
      // one warp sorts up to 128 keys in registers (4 per lane, key index = 4 * lane + r)
      if (warp == 0) {
        uint64_t kr[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) kr[r] = (4 * lane + r) < ns ? skeys[4 * lane + r] : kKeyMax;
#pragma unroll
        for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
          for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 4) {
              const int lj = j >> 2;  // partner lane distance
              const bool up = ((4 * lane) & k) == 0;
              const bool lower = (lane & lj) == 0;
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, kr[r], lj);
                const bool take_min = lower == up;
                kr[r] = take_min ? (kr[r] < o ? kr[r] : o) : (kr[r] > o ? kr[r] : o);
              }
            } else {
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                if ((r & j) == 0) {
                  const int i = 4 * lane + r;
                  const bool up = (i & k) == 0;
                  const uint64_t x = kr[r], y = kr[r | j];
                  if ((x > y) == up) { kr[r] = y; kr[r | j] = x; }
                }
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) skeys[4 * lane + r] = kr[r];
      }
      __syncthreads();
    } else {
      block_bitonic_sort(skeys, ns);
    }
    for (int i = tid; i < P; i += kRefineThreads) {
      const uint64_t k = i < max(ns, 128) && i < Cp ? skeys[i] : kKeyMax;
      lout[i] = (k == kKeyMax) ? -1 : (int32_t)(uint32_t)k;
      float bv = (k == kKeyMax) ? 0.f : ord2f((uint32_t)(k >> 32));
      // the key's order-preserving image folds -0.0 into +0.0.  The only zero with a sign: the one-to-one kernel of
      // the last L mod 3 int8 centres returns -(+0.0) for a zero dot product (its sums start at +0.0, so they never end
      // at -0.0); the three-at-a-time kernel and every squared-L2 distance end at +0.0
      if (i8 && !sql2 && bv == 0.f && k != kKeyMax && (int)(uint32_t)k >= L3) bv = -0.0f;
      bout[i] = bv;
    }
  } else {
    if (tid == 0 && fallbacks) atomicAdd(fallbacks, 1u);
    if (tid == 0 && ix.tok_fallback_flag) ix.tok_fallback_flag[qi] = 1;
    for (int i = tid; i < L; i += kRefineThreads) row[i] = exact(i);
    __syncthreads();
    if (warp == 0) warp_topp_exact(row, L, P, Ppow2, hist, skeys, lane, lout, bout);
  }
}

// Streaming refinement for long rows and P <= 128: one warp per query, two coalesced passes over the row instead of a
// radix select on a shared-memory image of it.  Pass 1 keeps the minimum of 128 strided sub-rows (4 per lane); the
// P-th smallest of those minima is an upper bound U of the P-th smallest approximate distance (P different sub-rows
// each hold an element <= U).  Pass 2 collects the centres with approx <= U + 2 eps -- a superset of the exact top P,
// see above -- which get the exact chain and a sort by (distance, leaf).  More than kStreamCand candidates (degenerate
// rows) fall back to exact distances for all centres.
// MEASURED AND REJECTED as the default (kept behind SCANN_B200_TOKENIZE=stream, bit-exact, tested): with one warp per
// query the two passes are a serial chain of L / 256 round trips to HBM per warp; 10k queries x 40k centres take
// 3.2 ms against 2.65 ms for the radix refinement (CTA per query, row staged in shared memory), 8000 centres 0.71
// against 0.38 ms, and the index-build stage's tokenization 1.24 s against 0.75 s per 20M rows.
constexpr int kStreamCand = 256;
__global__ void __launch_bounds__(256)
topp_stream_kernel(DevIndex ix, const float* __restrict__ q, float* __restrict__ S, int nq, int P, int Ppow2,
                   float eps_rel, int32_t* __restrict__ leaves, float* __restrict__ bias, uint32_t* fallbacks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qi = blockIdx.x * 8 + warp;
  if (qi >= nq) return;
  const int L = (int)ix.L, D = (int)ix.d, Dp = (D + 3) & ~3;
  unsigned char* base = smem_raw + (size_t)warp * ((size_t)kStreamCand * 8 + 1024 + (size_t)Dp * 4);
  uint64_t* skeys = reinterpret_cast<uint64_t*>(base);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + (size_t)kStreamCand * 8);
  float* mins = reinterpret_cast<float*>(hist);  // 128 floats; dead before `hist` is used
  float* sq = reinterpret_cast<float*>(base + (size_t)kStreamCand * 8 + 1024);
  float* row = S + (size_t)qi * L;
  const bool sql2 = ix.distance == 1;
  const uint32_t lt = (1u << lane) - 1u;

  float ssq = 0.f;
  for (int k = lane; k < Dp; k += 32) {
    const float v = k < D ? q[(size_t)qi * D + k] : 0.f;
    sq[k] = v;
    ssq = fmaf(v, v, ssq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xFFFFFFFFu, ssq, o);
  __syncwarp();
  float qn = 0.f;
  if (sql2) {  // ||q||^2 exactly as tokenize_kernel: float(SquaredL2Norm(q))
    float acc = 0.f;
    if (lane == 0) acc = squared_l2_norm_strided(sq, (uint32_t)D);
    qn = __shfl_sync(0xFFFFFFFFu, acc, 0);
  }
  const float qnorm = sqrtf(ssq) * 1.001f, cmax = ix.center_max_norm;
  float eps = eps_rel * qnorm * cmax;
  if (sql2) eps = 2.f * eps + (float)(D + 8) * 1.1920929e-7f * (qn + cmax * cmax + 2.f * qnorm * cmax);
  auto approx = [&](int i) -> float {
    const float sdot = row[i];
    return sql2 ? __fsub_rn(__fadd_rn(__ldg(ix.center_sqnorm + i), qn), __fmul_rn(2.f, sdot)) : -sdot;
  };
  // ---- pass 1: minima of 128 strided sub-rows ----
  const float inf = __int_as_float(0x7F800000);
  float m[4] = {inf, inf, inf, inf};
  for (int i0 = 0; i0 < L; i0 += 256) {  // eight independent loads in flight per lane
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + 32 * j + lane;
      v[j] = i < L ? approx(i) : inf;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j & 3] = fminf(m[j & 3], v[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) mins[4 * lane + j] = m[j];
  __syncwarp();
  float U = inf;
  {
    int rank[4] = {0, 0, 0, 0};
    for (int t = 0; t < 128; ++t) {
      const float w = mins[t];
#pragma unroll
      for (int j = 0; j < 4; ++j) rank[j] += (w < m[j] || (w == m[j] && t < 4 * lane + j)) ? 1 : 0;
    }
    float val = inf;
    bool has = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (rank[j] == P - 1) { has = true; val = m[j]; }
    const uint32_t hm = __ballot_sync(0xFFFFFFFFu, has);
    if (hm) U = __shfl_sync(0xFFFFFFFFu, val, __ffs(hm) - 1);
  }
  __syncwarp();
  const float thr = __fadd_ru(U, __fmul_ru(2.f, eps));
  // ---- pass 2: candidates ----
  uint32_t count = 0;
  for (int i0 = 0; i0 < L; i0 += 256) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + 32 * j + lane;
      v[j] = i < L ? approx(i) : inf;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + 32 * j + lane;
      const bool c = i < L && !(v[j] > thr);  // NaN is a candidate
      const uint32_t mk = __ballot_sync(0xFFFFFFFFu, c);
      if (mk) {
        if (c) {
          const uint32_t pos = count + __popc(mk & lt);
          if (pos < (uint32_t)kStreamCand) skeys[pos] = (uint64_t)(uint32_t)i;
        }
        count += __popc(mk);
      }
    }
  }
  __syncwarp();
  auto exact = [&](int idx) -> float {
    const float* c = ix.centers + (size_t)idx * D;
    float acc = sql2 ? __fadd_rn(ix.center_sqnorm[idx], qn) : 0.f;
    const float scale2 = sql2 ? 2.0f : 1.0f;
    if ((D & 3) == 0) {
      for (int k = 0; k < D; k += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(c + k));
        acc = __fmaf_rn(-sq[k], __fmul_rn(v.x, scale2), acc);
        acc = __fmaf_rn(-sq[k + 1], __fmul_rn(v.y, scale2), acc);
        acc = __fmaf_rn(-sq[k + 2], __fmul_rn(v.z, scale2), acc);
        acc = __fmaf_rn(-sq[k + 3], __fmul_rn(v.w, scale2), acc);
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < D; ++k) acc = __fmaf_rn(-sq[k], __fmul_rn(__ldg(c + k), scale2), acc);
    }
    return acc;
  };
  int32_t* lout = leaves + (size_t)qi * P;
  float* bout = bias + (size_t)qi * P;
  if (count <= (uint32_t)kStreamCand) {
    int ns = 2;
    while ((uint32_t)ns < count) ns <<= 1;
    for (int j = lane; j < ns; j += 32) {
      const bool live = (uint32_t)j < count;
      const uint32_t idx = live ? (uint32_t)skeys[j] : 0u;
      const float e = live ? exact((int)idx) : 0.f;
      skeys[j] = live ? (((uint64_t)f2ord(e) << 32) | idx) : kKeyMax;
    }
    __syncwarp();
    warp_bitonic_sort(skeys, ns, lane);
    for (int i = lane; i < P; i += 32) {
      const uint64_t k = i < ns ? skeys[i] : kKeyMax;
      lout[i] = (k == kKeyMax) ? -1 : (int32_t)(uint32_t)k;
      bout[i] = (k == kKeyMax) ? 0.f : ord2f((uint32_t)(k >> 32));
    }
  } else {
    if (lane == 0 && fallbacks) atomicAdd(fallbacks, 1u);
    if (lane == 0 && ix.tok_fallback_flag) ix.tok_fallback_flag[qi] = 1;
    for (int i = lane; i < L; i += 32) row[i] = exact(i);
    __syncwarp();
    warp_topp_exact(row, L, P, Ppow2, hist, skeys, lane, lout, bout);
  }
}

// Chunk pre-selection for long rows (P <= 128, L / 32 >= 2 P): the GEMM's store epilogue also writes the maximum of
// every 32-centre chunk of the row (of 2 S - ||c||^2 for squared L2), so the refinement reads L / 32 values instead of L.  The P-th smallest
// chunk minimum U (= -chunk maximum; exact, by the warp radix select) is an upper bound of the P-th smallest
// approximate distance -- P different chunks each hold an element <= U -- and only chunks whose minimum is <= U + 2 eps
// can hold a candidate.  Those chunks (about P of them) are read back, 128 bytes each, their elements <= U + 2 eps get
// the exact chain and the sort by (distance, leaf): the same superset argument and the same result as the radix
// refinement, with 5 KB + ~P * 128 B read per query instead of 160 KB at L = 40k, one warp per query and no row image.
constexpr int kChunkMaxChunks = 2048;  // L <= 65536
__global__ void __launch_bounds__(256)
topp_chunk_kernel(DevIndex ix, const float* __restrict__ q, float* __restrict__ S, const float* __restrict__ cmax_ws,
                  int nq, int P, int Ppow2, float eps_rel, int32_t* __restrict__ leaves, float* __restrict__ bias,
                  uint32_t* fallbacks, int have_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qi = blockIdx.x * 8 + warp;
  if (qi >= nq) return;
  const int L = (int)ix.L, D = (int)ix.d, Dp = (D + 3) & ~3, Lc = (L + 31) >> 5, Lcp = (Lc + 3) & ~3;
  unsigned char* base = smem_raw + (size_t)warp * ((size_t)kStreamCand * 8 + 1024 + (size_t)Dp * 4 + (size_t)Lcp * 4);
  uint64_t* skeys = reinterpret_cast<uint64_t*>(base);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + (size_t)kStreamCand * 8);
  float* sq = reinterpret_cast<float*>(base + (size_t)kStreamCand * 8 + 1024);
  float* cm = sq + Dp;  // [Lc] chunk minima of the approximate distance
  float* row = S + (size_t)qi * L;
  const uint32_t lt = (1u << lane) - 1u;

  const bool sql2 = ix.distance == 1;
  const bool i8 = ix.centers_i8 != nullptr;  // int8 tokenization: S = <q', float(int8 centre)>, see topp_refine_kernel
  const int L3 = L / 3 * 3;
  float ssq = 0.f;
  for (int k = lane; k < Dp; k += 32) {
    float v = k < D ? q[(size_t)qi * D + k] : 0.f;
    if (i8 && k < D) v = __fmul_rn(v, ix.cen_qscale[k]);
    sq[k] = v;
    ssq = fmaf(v, v, ssq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xFFFFFFFFu, ssq, o);
  __syncwarp();
  float qn = 0.f;
  if (sql2) {  // ||q||^2 exactly as tokenize_kernel: float(SquaredL2Norm(q))
    float acc = 0.f;
    if (lane == 0) acc = squared_l2_norm_strided(i8 ? q + (size_t)qi * D : sq, (uint32_t)D);
    qn = __shfl_sync(0xFFFFFFFFu, acc, 0);
  }
  // chunk minima of the approximate distance: -max(S) (dot product), -max(2 S - ||c||^2) + ||q||^2 (squared L2)
  const float* cmr = cmax_ws + (size_t)qi * Lc;
  // int8, squared L2: the statistic is max(2 S - 2 |c|^2) (the GEMM gets 2 |c|^2 as its bias), the approximate
  // distance (|c|^2 + |q|^2) - S: chunk minimum = |q|^2 - statistic / 2
  for (int j = lane; j < Lc; j += 32) cm[j] = sql2 ? __fadd_rn(i8 ? __fmul_rn(-0.5f, cmr[j]) : -cmr[j], qn) : -cmr[j];
  __syncwarp();
  const float qnorm = sqrtf(ssq) * 1.001f, cmaxn = i8 ? ix.cen_i8_max_norm : ix.center_max_norm;
  float eps = eps_rel * qnorm * cmaxn;
  if (sql2 && i8) eps = eps + 16.f * 1.1920929e-7f * (qn + ix.cen_sqnorm_max + qnorm * cmaxn);
  // squared L2: as the radix refinement, plus the few ulps by which "(b - 2S) + |q|^2" of the chunk statistic and
  // "(b + |q|^2) - 2S" of the element test may differ
  if (sql2 && !i8) eps = 2.f * eps + (float)(D + 16) * 1.1920929e-7f * (qn + cmaxn * cmaxn + 2.f * qnorm * cmaxn);
  auto approx = [&](int i) -> float {
    const float sdot = row[i];
    if (i8) return sql2 ? __fsub_rn(__fadd_rn(__ldg(ix.cen_sqnorm + i), qn), sdot) : -sdot;
    return sql2 ? __fsub_rn(__fadd_rn(__ldg(ix.center_sqnorm + i), qn), __fmul_rn(2.f, sdot)) : -sdot;
  };
  uint32_t need = 0;
  const uint32_t prefix = warp_radix_select([&](int j) { return f2ord(cm[j]); }, Lc, (uint32_t)P, hist, lane, need);
  const float U = ord2f(prefix);
  const float thr = __fadd_ru(U, __fmul_ru(2.f, eps));
  auto exact = [&](int idx) -> float {  // the reference's sequential fnmadd chain (tokenize_kernel)
    if (i8) return i8_center_distance(ix, sq, idx, D, L3, sql2, qn);
    const float* c = ix.centers + (size_t)idx * D;
    float acc = sql2 ? __fadd_rn(ix.center_sqnorm[idx], qn) : 0.f;
    const float scale2 = sql2 ? 2.0f : 1.0f;
    if ((D & 3) == 0) {
      for (int k = 0; k < D; k += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(c + k));
        acc = __fmaf_rn(-sq[k], __fmul_rn(v.x, scale2), acc);
        acc = __fmaf_rn(-sq[k + 1], __fmul_rn(v.y, scale2), acc);
        acc = __fmaf_rn(-sq[k + 2], __fmul_rn(v.z, scale2), acc);
        acc = __fmaf_rn(-sq[k + 3], __fmul_rn(v.w, scale2), acc);
      }
    } else {
#pragma unroll 4
      for (int k = 0; k < D; ++k) acc = __fmaf_rn(-sq[k], __fmul_rn(__ldg(c + k), scale2), acc);
    }
    return acc;
  };
  // Without the stored matrix (have_rows == 0: the GEMM only wrote the chunk maxima) the element test runs on the
  // exact chain itself: every centre of a candidate chunk is evaluated (32 consecutive centre rows, one per lane) and
  // kept if exact <= U + eps -- an element of the exact top P has chain <= T* <= U + eps, T* being the P-th smallest
  // chain distance, because the P chunks with minimum <= U each hold an element with chain <= U + eps.  The keys carry
  // the exact distance already (bit 63 of the buffered key marks them).
  const float thr_exact = __fadd_ru(U, eps);
  uint32_t count = 0;
  for (int j0 = 0; j0 < Lc; j0 += 32) {
    const int j = j0 + lane;
    uint32_t hits = __ballot_sync(0xFFFFFFFFu, j < Lc && !(cm[j] > thr));
    while (hits) {
      const int jj = j0 + __ffs(hits) - 1;
      hits &= hits - 1;
      const int i = jj * 32 + lane;
      float e = 0.f;
      bool c;
      if (have_rows) {
        c = i < L && !(approx(i) > thr);
      } else {
        if (i < L) e = exact(i);
        c = i < L && !(e > thr_exact);
      }
      const uint32_t mk = __ballot_sync(0xFFFFFFFFu, c);
      if (c) {
        const uint32_t pos = count + __popc(mk & lt);
        if (pos < (uint32_t)kStreamCand)
          skeys[pos] = have_rows ? (uint64_t)(uint32_t)i : (((uint64_t)f2ord(e) << 32) | (uint32_t)i);
      }
      count += __popc(mk);
    }
  }
  __syncwarp();
  int32_t* lout = leaves + (size_t)qi * P;
  float* bout = bias + (size_t)qi * P;
  if (count <= (uint32_t)kStreamCand) {
    int ns = 2;
    while ((uint32_t)ns < count) ns <<= 1;
    for (int j = lane; j < ns; j += 32) {
      const bool live = (uint32_t)j < count;
      if (have_rows) {
        const uint32_t idx = live ? (uint32_t)skeys[j] : 0u;
        const float e = live ? exact((int)idx) : 0.f;
        skeys[j] = live ? (((uint64_t)f2ord(e) << 32) | idx) : kKeyMax;
      } else if (!live) {
        skeys[j] = kKeyMax;
      }
    }
    __syncwarp();
    warp_bitonic_sort(skeys, ns, lane);
    for (int i = lane; i < P; i += 32) {
      const uint64_t k = i < ns ? skeys[i] : kKeyMax;
      lout[i] = (k == kKeyMax) ? -1 : (int32_t)(uint32_t)k;
      float bv = (k == kKeyMax) ? 0.f : ord2f((uint32_t)(k >> 32));
      if (i8 && !sql2 && bv == 0.f && k != kKeyMax && (int)(uint32_t)k >= L3) bv = -0.0f;  // see topp_refine_kernel
      bout[i] = bv;
    }
  } else {
    if (lane == 0 && fallbacks) atomicAdd(fallbacks, 1u);
    if (lane == 0 && ix.tok_fallback_flag) ix.tok_fallback_flag[qi] = 1;
    for (int i = lane; i < L; i += 32) row[i] = exact(i);
    __syncwarp();
    warp_topp_exact(row, L, P, Ppow2, hist, skeys, lane, lout, bout);
  }
}

bool tokenize_tensor_path(const DevIndex& ix, uint32_t P) {
  // int8 tokenization takes the same route (the centre operand then holds the int8 centres, exact in bf16): tcgen05
  // pre-filter + radix refinement with the int8 chain, or the exact SIMT kernels (launch_tokenize)
  if (!ix.tok_b || P + 32 > (uint32_t)kRefineMaxCand || ix.d > 2048) return false;
  const char* e = getenv("SCANN_B200_TOKENIZE");
  if (e && !strcmp(e, "simt")) return false;
  if (e && (!strcmp(e, "tcgen05") || !strcmp(e, "stream") || !strcmp(e, "chunk"))) return true;
  return ix.L >= 256;  // below that the SIMT GEMM is already negligible
}

cudaError_t launch_tokenize_topp(const DevIndex& ix, const float* q, uint32_t nq, uint32_t P, float* dist, void* a_ws,
                                 int32_t* leaves, float* bias, uint32_t* fallbacks, cudaStream_t s, int* launches) {
  if (!tokenize_tensor_path(ix, P)) {
    launch_tokenize(ix, q, nq, dist, s);
    launch_topp(ix, dist, nq, P, leaves, bias, s);
    if (launches) *launches += (ix.centers_i8 && ix.L % 3) ? 3 : 2;
    return cudaGetLastError();
  }
  const bool i8 = ix.centers_i8 != nullptr;
  cudaError_t e = build_tokenize_operand(q, nq, ix.d, 1, a_ws, s, i8 ? ix.cen_qscale : nullptr);
  if (e != cudaSuccess) return e;
  // chunk pre-selection: P <= 128, at least 2 P chunks; default from 4096 centres
  // (SCANN_B200_TOKENIZE=chunk forces it where it applies, =tcgen05 the radix refinement)
  const uint32_t n_chunks = (ix.L + 31) / 32;
  bool chunked = ix.tok_cmax_ws && P <= 128 && n_chunks >= 2 * P && n_chunks <= (uint32_t)kChunkMaxChunks;
  {
    const char* env = getenv("SCANN_B200_TOKENIZE");
    if (env && (!strcmp(env, "tcgen05") || !strcmp(env, "stream"))) chunked = false;
    else if (!(env && !strcmp(env, "chunk")) && ix.L < 4096) chunked = false;
  }
  // SCANN_B200_TOKENIZE_ROWS=0: do not store the matrix when nobody else reads it (chunk pre-selection without
  // tok_need_rows); the element test then evaluates the exact chain for whole candidate chunks.  Measured and rejected
  // as the default: at 40k centres the GEMM is bound by its operand traffic from L2, not by the store (1.19 -> 1.11 ms
  // at P = 24), and the 32 exact chains per candidate chunk cost more than the stored row saves (1.40 -> 3.08 ms at
  // P = 80).
  bool store_rows = true;
  if (const char* env = getenv("SCANN_B200_TOKENIZE_ROWS")) store_rows = !(env[0] == '0' && chunked && ix.tok_need_rows == 0);
  e = gemm_bf16_nt(a_ws, nq, (nq + 127) / 128 * 128, ix.tok_b, ix.L, ix.tok_kp, store_rows ? dist : nullptr, ix.L, s,
                   chunked ? ix.tok_cmax_ws : nullptr, n_chunks, ix.distance == 1 ? (i8 ? ix.cen_sqnorm2 : ix.center_sqnorm) : nullptr);
  if (e != cudaSuccess) return e;
  if (chunked) {
    int sp = 2;
    while (sp < (int)P) sp <<= 1;
    const float er = (float)ix.tok_kp * 4.76837158e-7f + 3.05175781e-5f;  // K * 2^-21 + 2^-15
    const size_t per_warp = (size_t)kStreamCand * 8 + 1024 + (size_t)((ix.d + 3) & ~3u) * 4 + (size_t)((n_chunks + 3) & ~3u) * 4;
    const size_t bytes = per_warp * 8;
    e = cudaFuncSetAttribute(topp_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    topp_chunk_kernel<<<(nq + 7) / 8, 256, bytes, s>>>(ix, q, dist, ix.tok_cmax_ws, (int)nq, (int)P, sp, er, leaves, bias,
                                                       fallbacks, store_rows ? 1 : 0);
    if (launches) *launches += 3;
    return cudaGetLastError();
  }
  int pp = 2, cp = 64;
  while (pp < (int)P) pp <<= 1;
  while (cp < (int)P + 32) cp <<= 1;
  if (cp < pp) cp = pp;
  const float eps_rel = (float)ix.tok_kp * 4.76837158e-7f + 3.05175781e-5f;  // K * 2^-21 + 2^-15
  {
    // SCANN_B200_TOKENIZE=stream: the streaming refinement (P <= 128)
    const char* env = getenv("SCANN_B200_TOKENIZE");
    const bool force_stream = env && !strcmp(env, "stream"), force_radix = env && !strcmp(env, "tcgen05");
    (void)force_radix;
    if (P <= 128 && force_stream && !i8) {  // measured slower than the radix refinement (see the kernel's comment): opt-in only
      int sp = 2;
      while (sp < (int)P) sp <<= 1;
      const size_t per_warp = (size_t)kStreamCand * 8 + 1024 + (size_t)((ix.d + 3) & ~3u) * 4;
      const size_t bytes = per_warp * 8;
      e = cudaFuncSetAttribute(topp_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess) return e;
      topp_stream_kernel<<<(nq + 7) / 8, 256, bytes, s>>>(ix, q, dist, (int)nq, (int)P, sp, eps_rel, leaves, bias, fallbacks);
      if (launches) *launches += 3;
      return cudaGetLastError();
    }
  }
  if (cp < 128) cp = 128;  // the in-register sort writes 128 keys back
  const size_t smem = (size_t)cp * 8 + (size_t)((ix.d + 3) & ~3u) * 4;
#define SB_REFINE(kS, kT, bytes)                                                                                  \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(topp_refine_kernel<kS, kT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
    if (e != cudaSuccess) return e;                                                                               \
    topp_refine_kernel<kS, kT><<<nq, kT, (bytes), s>>>(ix, q, dist, (int)nq, (int)P, pp, cp, eps_rel, leaves, bias, \
                                                       fallbacks);                                                \
  } while (0)
  // more threads per query for longer rows (the row image stays in shared memory up to L = 49152)
  auto rowbytes = [&](uint32_t threads) { return (size_t)((ix.L + 4 * threads - 1) / (4 * threads) * (4 * threads)) * 4; };
  const bool smem_row = ix.L <= (uint32_t)kRefineSmemL && (ix.L & 3u) == 0;
  if (smem_row && ix.L <= 4096) SB_REFINE(true, 128, smem + rowbytes(128));
  else if (smem_row && ix.L <= 16384) SB_REFINE(true, 256, smem + rowbytes(256));
  else if (smem_row) SB_REFINE(true, 512, smem + rowbytes(512));
  else SB_REFINE(false, 256, smem);
#undef SB_REFINE
  if (launches) *launches += 3;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// AH lookup tables: raw[b][c] = lookup_distance(q_b, centre c), mult = 127 / max|raw|,
// lut = u8(round(raw * mult) + 128).  The generic one-to-many path of the reference is
// used for 16-centre codebooks: centres 0..14 go through the accumulating kernel (Highway
// lanes for dims < 8, AVX2 FMA lanes otherwise), centre 15 through the SSE4 one-to-one dot
// (one_to_many_symmetric.h:704-705,793-799).
// Output layout: lut[q][8W][16] u8 with rows >= B zeroed (so padded blocks add 0).
// ---------------------------------------------------------------------------------------
constexpr int kLutThreads = 256;

__global__ void __launch_bounds__(kLutThreads)
lut_kernel(DevIndex ix, const float* __restrict__ q, uint8_t* __restrict__ lut,
           float* __restrict__ mult_out, float* __restrict__ inv_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);          // [D]
  float* raw = sq + ((ix.d + 3) & ~3u);                     // [B*16]
  __shared__ float red[kLutThreads / 32];
  __shared__ float s_mult;
  const int tid = threadIdx.x;
  const uint32_t qi = blockIdx.x;
  for (uint32_t k = tid; k < ix.d; k += kLutThreads) sq[k] = q[(size_t)qi * ix.d + k];
  __syncthreads();
  const uint32_t ne = ix.B * 16;
  float mx = 0.f;
  for (uint32_t e = tid; e < ne; e += kLutThreads) {
    const float r = lut_raw_entry(ix, sq, e);
    raw[e] = r;
    mx = fmaxf(mx, fabsf(r));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int w = 0; w < kLutThreads / 32; ++w) m = fmaxf(m, red[w]);
    const float mult = lut_multiplier(m);
    s_mult = mult;
    mult_out[qi] = mult;
    inv_out[qi] = lut_inverse_multiplier(ix, mult);
  }
  __syncthreads();
  const float mult = s_mult;
  uint8_t* out = lut + (size_t)qi * ix.W * 8 * 16;
  const uint32_t npad = ix.W * 8 * 16;
  for (uint32_t e = tid; e < npad; e += kLutThreads) {
    uint8_t v = 0;
    if (e < ne) v = (uint8_t)lut_quantize(raw[e], mult);
    out[e] = v;
  }
}

void launch_lut(const DevIndex& ix, const float* q, uint32_t nq, uint8_t* lut, float* mult,
                float* inv_mult, cudaStream_t s) {
  const size_t smem = (((size_t)ix.d + 3) & ~(size_t)3) * 4 + (size_t)ix.B * 16 * 4;
  lut_kernel<<<nq, kLutThreads, smem, s>>>(ix, q, lut, mult, inv_mult);
}

}  // namespace sb
