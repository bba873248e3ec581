// encode.cu -- the deterministic part of index construction on the GPU:
// database tokenization, SOAR secondary assignment, residuals and AH encoding (plain and noise-shaped),
// behind scann_b200_encode_database (include/scann_b200.h).
//
// Host-side mirror of, in the reference (paths relative to /root/reference/scann/):
//   KMeansTreePartitioner::TokenizeDatabase                         partitioning/kmeans_tree_partitioner.cc:475-560
//   ... ::OrthogonalityAmplifiedTokenForDatapointBatched            partitioning/kmeans_tree_partitioner.cc:925-997
//   DenseManyToManyOrthogonalityAmplified                           distance_measures/many_to_many/many_to_many_impl.inc:729-781
//   TreeAHHybridResidual::BuildLeafSearchers (get_hashed_datapoint) tree_x_hybrid/tree_ah_hybrid_residual.cc:395-428
//   Indexer::Hash / HashWithNoiseShaping                            hashes/asymmetric_hashing2/indexing.cc:87-246
//   AhImpl::IndexDatapoint / IndexDatapointNoiseShaped               hashes/internal/asymmetric_hashing_impl.cc:199-244,434-503
// The trainers (k-means tree, AH codebooks) are random-initialised and are not part of this stage: the
// caller passes trained centres and a codebook.  Every output is bit-identical to oracle/scann_oracle.c
// (so_assign_primary, so_assign_soar, so_encode).  There is no CPU path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/scann_b200.h"
#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"

namespace sb {
void set_last_error(const char* msg);

constexpr int kTeam = 16;          // lanes per (datapoint, token) pair: one lane per codebook centre
constexpr int kEncodeThreads = 128;
constexpr int kSoarThreads = 256;  // 8 warps, one datapoint per warp
constexpr int kSoarCand = 64;      // per-warp candidate list (two rounds of 32)
constexpr uint32_t kFull = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// SOAR secondary assignment.  One warp per datapoint.
//   cost(c) = t1 + (lambda * t2) * t2,  t1 = sum fma(diff, diff), t2 = sum fma(diff, rhat), diff = x - c,
// sequentially in the dimension, fp32, first strict minimum over ALL centres (the primary included).
// Evaluating that chain for every centre is O(N L D) on the fp32 pipe; instead the row of approximate
// squared distances that the tokenization GEMM left in `row` prunes the centres: cost(c) >= t1(c) >=
// ||x - c||^2 (1 - e1) and ||x - c||^2 >= approx(c) - E, so a centre with (approx(c) - E)(1 - e1) > m
// cannot beat (nor tie) the best cost m found so far.  m starts from the exact cost of the `P` nearest
// centres; survivors are collected per warp and evaluated 32 at a time, which tightens m as the scan goes.
// The result is the exact argmin by (cost, centre) whatever the pruning order.
// ---------------------------------------------------------------------------------------------
struct SoarArgs {
  const float* x;          // [n][D]
  const float* centers;    // [L][D]
  const float* cnorm;      // [L] ||c||^2 (fnmadd chain)
  const float* row;        // [n][L]: x.c (row_is_dot) or squared distances
  const float* row2;       // [n][L]: rhat.c from the second GEMM, or NULL (no projection pruning)
  const uint8_t* row_exact; // [n] or NULL: 1 = this row of `row` holds exact squared distances (tokenizer fallback)
  float* rhat;             // [n][D] normalised residuals (written by rhat_kernel, read by soar_kernel)
  const int32_t* near;     // [n][P] nearest centres, sorted by (distance, centre); near[.][0] is the primary
  int32_t* sec;            // [n] out
  unsigned long long* evaluated;  // exact cost evaluations, summed (statistics)
  uint32_t n, L, D, P;
  int row_is_dot;
  float lambda, eps_rel, cmax;
};

__device__ __forceinline__ float soar_cost(const float* __restrict__ xs, const float* __restrict__ rh,
                                           const float* __restrict__ c, uint32_t D, float lambda) {
  float t1 = 0.f, t2 = 0.f;
  if ((D & 3u) == 0) {
    for (uint32_t k = 0; k < D; k += 4) {
      const float4 cv = __ldg(reinterpret_cast<const float4*>(c + k));
      float diff = __fsub_rn(xs[k], cv.x);
      t1 = __fmaf_rn(diff, diff, t1); t2 = __fmaf_rn(diff, rh[k], t2);
      diff = __fsub_rn(xs[k + 1], cv.y);
      t1 = __fmaf_rn(diff, diff, t1); t2 = __fmaf_rn(diff, rh[k + 1], t2);
      diff = __fsub_rn(xs[k + 2], cv.z);
      t1 = __fmaf_rn(diff, diff, t1); t2 = __fmaf_rn(diff, rh[k + 2], t2);
      diff = __fsub_rn(xs[k + 3], cv.w);
      t1 = __fmaf_rn(diff, diff, t1); t2 = __fmaf_rn(diff, rh[k + 3], t2);
    }
  } else {
    for (uint32_t k = 0; k < D; ++k) {
      const float diff = __fsub_rn(xs[k], __ldg(c + k));
      t1 = __fmaf_rn(diff, diff, t1);
      t2 = __fmaf_rn(diff, rh[k], t2);
    }
  }
  return __fadd_rn(t1, __fmul_rn(__fmul_rn(lambda, t2), t2));
}

// ComputeNormalizedResidual (orthogonality_amplification_utils.h:27-46), one warp per datapoint:
// out = float(double(x) - double(c)); sqnorm = sequential double sum of out^2; zero if sqnorm < 1e-7, else
// out *= float(1 / sqrt(sqnorm)).  The rows are the A operand of the second GEMM (rhat . c for every centre).
__global__ void __launch_bounds__(kSoarThreads)
rhat_kernel(SoarArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * (kSoarThreads / 32) + warp;
  if (i >= a.n) return;
  const float* x = a.x + (size_t)i * a.D;
  const float* pc = a.centers + (size_t)a.near[(size_t)i * a.P] * a.D;
  float* out = a.rhat + (size_t)i * a.D;
  for (uint32_t k = lane; k < a.D; k += 32) out[k] = (float)__dsub_rn((double)x[k], (double)__ldg(pc + k));
  __syncwarp();
  double sqnorm = 0.0;
  if (lane == 0)
    for (uint32_t k = 0; k < a.D; ++k) sqnorm = __dadd_rn(sqnorm, __dmul_rn((double)out[k], (double)out[k]));
  sqnorm = __shfl_sync(kFull, sqnorm, 0);
  const bool degenerate = sqnorm < 1e-7;
  const float inv_norm = degenerate ? 0.f : (float)(1.0 / sqrt(sqnorm));
  for (uint32_t k = lane; k < a.D; k += 32) out[k] = degenerate ? 0.f : __fmul_rn(out[k], inv_norm);
}

__global__ void __launch_bounds__(kSoarThreads)
soar_kernel(SoarArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t Dp = (a.D + 3) & ~3u;
  float* xs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (2 * Dp + kSoarCand);
  float* rh = xs + Dp;
  int32_t* cand = reinterpret_cast<int32_t*>(rh + Dp);
  const uint32_t i = blockIdx.x * (kSoarThreads / 32) + warp;
  if (i >= a.n) return;
  const float* x = a.x + (size_t)i * a.D;
  const int32_t prim = a.near[(size_t)i * a.P];
  float ssq = 0.f;
  double xr64 = 0.0;  // <x, rhat>: the centre-independent half of the projection <x - c, rhat>
  for (uint32_t k = lane; k < a.D; k += 32) {
    const float v = x[k], r = a.rhat[(size_t)i * a.D + k];
    xs[k] = v;
    rh[k] = r;
    ssq = fmaf(v, v, ssq);
    xr64 = fma((double)v, (double)r, xr64);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssq += __shfl_xor_sync(kFull, ssq, o);
    xr64 += __shfl_xor_sync(kFull, xr64, o);
  }
  __syncwarp();
  double qn64 = 0.0;
  if (lane == 0)  // ||x||^2 as the tokenizer computes it (sequential double accumulation, narrowed)
    for (uint32_t k = 0; k < a.D; ++k) qn64 = __dadd_rn(qn64, __dmul_rn((double)xs[k], (double)xs[k]));
  const float qn = (float)__shfl_sync(kFull, qn64, 0);
  const float xr = (float)xr64;

  float m = __int_as_float(0x7F800000);
  int32_t best = 0x7FFFFFFF;
  unsigned long long evals = 0;
  auto evaluate = [&](int32_t c) {  // every lane passes its own centre (or -1), the warp keeps the minimum
    float cost = __int_as_float(0x7F800000);
    int32_t idx = 0x7FFFFFFF;
    if (c >= 0) { cost = soar_cost(xs, rh, a.centers + (size_t)c * a.D, a.D, a.lambda); idx = c; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float oc = __shfl_xor_sync(kFull, cost, o);
      const int32_t oi = __shfl_xor_sync(kFull, idx, o);
      if (oc < cost || (oc == cost && oi < idx)) { cost = oc; idx = oi; }
    }
    if (cost < m || (cost == m && idx < best)) { m = cost; best = idx; }
  };
  {
    const uint32_t np = min(a.P, 32u);
    const int32_t c = (uint32_t)lane < np ? a.near[(size_t)i * a.P + lane] : -1;
    evaluate(c);
    evals += np;
  }
  // E1 bounds |approx - ||x - c||^2| (the tokenization pre-filter's eps, prep.cu, plus the chains' own rounding);
  // E2 bounds |(xr - S2) - t2| for the fp32 chain t2 = sum fma(diff, rhat): GEMM error of rhat.c (|rhat| <= 1),
  // the rounding of xr and of the chain ((D + 2) 2^-24 |x - c| |rhat|).  Then, with p = max(|xr - S2| - E2, 0):
  //   cost = fl(t1 + fl(fl(lambda t2) t2)) >= ((approx - E1) shrink + lambda p^2 (1 - 2^-22)) (1 - 2^-23).
  const float qnorm = sqrtf(ssq) * 1.001f;
  const float scale = qn + a.cmax * a.cmax + 2.f * qnorm * a.cmax;
  const float ulp = (float)(a.D + 8) * 1.1920929e-7f;
  const float E1 = 2.f * a.eps_rel * qnorm * a.cmax + 2.f * ulp * scale;
  const float E2 = 1.5f * a.eps_rel * a.cmax + 2.f * ulp * (qnorm + a.cmax);
  const float shrink = 1.f - ulp;  // t1 >= ||x - c||^2 * shrink
  const float lam_lo = a.lambda * (1.f - 4.76837158e-7f);
  const float* row = a.row + (size_t)i * a.L;
  const float* row2 = a.row2 ? a.row2 + (size_t)i * a.L : nullptr;
  const bool is_dot = a.row_is_dot && !(a.row_exact && a.row_exact[i]);
  uint32_t ncand = 0;
  // four 32-centre groups per iteration: the loads of all of them are in flight before the first test (the scan is a
  // serial chain of round trips per warp otherwise: issue 48 %, DRAM 17 % in the first capture)
  for (uint32_t base = 0; base < a.L; base += 128) {
    float v[4], w[4], cn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t l = base + 32 * u + lane;
      const bool in = l < a.L;
      v[u] = in ? row[l] : 0.f;
      w[u] = (in && row2) ? row2[l] : 0.f;
      cn[u] = (in && is_dot) ? __ldg(a.cnorm + l) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t l = base + 32 * u + lane;
      bool pass = false;
      if (l < a.L) {
        const float ap = is_dot ? __fsub_rn(__fadd_rn(cn[u], qn), __fmul_rn(2.f, v[u])) : v[u];
        float lb = __fmul_rd(__fsub_rd(ap, E1), shrink);
        if (row2) {
          const float p = fmaxf(__fsub_rd(fabsf(__fsub_rn(xr, w[u])), E2), 0.f);
          lb = __fmul_rd(__fadd_rd(lb, __fmul_rd(__fmul_rd(p, p), lam_lo)), 1.f - 1.1920929e-7f);
        }
        pass = !(lb > m);  // NaN passes
      }
      const uint32_t mask = __ballot_sync(kFull, pass);
      if (mask) {
        if (pass) cand[ncand + __popc(mask & ((1u << lane) - 1u))] = (int32_t)l;
        ncand += __popc(mask);
        __syncwarp();
        if (ncand >= 32) {
          evaluate(cand[lane]);
          evals += 32;
          const int32_t carry = (uint32_t)lane + 32 < ncand ? cand[lane + 32] : -1;
          __syncwarp();
          ncand -= 32;
          if ((uint32_t)lane < ncand) cand[lane] = carry;
          __syncwarp();
        }
      }
    }
  }
  if (ncand) {
    evaluate((uint32_t)lane < ncand ? cand[lane] : -1);
    evals += ncand;
  }
  if (lane == 0) {
    if (best == 0x7FFFFFFF) best = prim;  // every cost NaN (non-finite input): leave the datapoint unspilled
    a.sec[i] = best;
    atomicAdd(a.evaluated, evals);
  }
}

// ---------------------------------------------------------------------------------------------
// Serialized token layout (scann_ops/cc/scann.cc:533-551): without SOAR tokens[i] = primary; with SOAR
// tokens[2i] = the lower-numbered leaf, tokens[2i + 1] = the other leaf, or -1 when the secondary equals the
// primary (the datapoint is not spilled, kmeans_tree_partitioner.cc:527-531).
// ---------------------------------------------------------------------------------------------
__global__ void tokens_kernel(const int32_t* __restrict__ near, uint32_t P, const int32_t* __restrict__ sec, uint32_t n,
                              int32_t* __restrict__ tokens, unsigned long long* spilled) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t p = near[(size_t)i * P];
  if (!sec) { tokens[i] = p; return; }
  const int32_t s = sec[i];
  if (s == p) { tokens[2 * i] = p; tokens[2 * i + 1] = -1; return; }
  tokens[2 * i] = min(p, s);
  tokens[2 * i + 1] = max(p, s);
  atomicAdd(spilled, 1ull);
}

// ---------------------------------------------------------------------------------------------
// AH encoding.  A team of 16 lanes owns one (datapoint, token) pair; lane c owns codebook centre c.
//   plain (threshold NaN): per block, squared L2 to the 16 centres with the arithmetic of the reference's
//     one-to-many kernel (centres 0..14 accumulating kernel, centre 15 the SSE4 one-to-one kernel, as in the LUT
//     build, prep.cu), first minimum.
//   noise-shaped: residual statistics per (block, centre) in double, initial codes = minimum residual norm, blocks
//     visited in descending order of that norm, <= 10 rounds of coordinate descent on
//     eta * parallel^2 + perpendicular^2 (asymmetric_hashing_impl.cc:376-503).  All double operations are written
//     with explicit round-to-nearest intrinsics (no FMA contraction: the reference's translation unit is compiled
//     without FMA).
// ---------------------------------------------------------------------------------------------
struct EncodeArgs {
  const float* x;           // [n][D]
  const float* centers;     // [L][D] or NULL (hash x itself)
  const int32_t* tokens;    // [n][npd]
  const float* codebook;    // [B][16][S]
  const int32_t* block_dims;
  const uint32_t* block_off;
  uint8_t* codes;           // [n][B]  pair 0
  uint8_t* soar_codes;      // [n][B]  pair 1 (npd == 2)
  unsigned long long* ties; // datapoints whose initial block norms tie
  uint32_t n, D, B, S, npd;
  double threshold;         // NaN = plain
};

__host__ __device__ __forceinline__ size_t encode_team_bytes(uint32_t D, uint32_t B, bool shaped) {
  const size_t Dp = (D + 3) & ~3u;
  size_t b = 2 * Dp * sizeof(float);                        // res, orig
  if (shaped) b += (size_t)B * sizeof(double)               // initial norms
                   + 2 * Dp * sizeof(double);               // res, orig widened once (the descent re-reads them)
  b += ((size_t)B * 3 + 15) & ~(size_t)15;                  // code u8, order u16
  return (b + 15) & ~(size_t)15;
}

// u64 image of a double that orders like the double (NaNs at the two ends)
__device__ __forceinline__ uint64_t d2ord(double v) {
  const uint64_t u = (uint64_t)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord2d(uint64_t o) {
  return __longlong_as_double((long long)((o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o));
}
// Lowest lane of the team (mask tmask, first lane tbase) that holds the minimum of v; *vmin = that minimum.
// Two 32-bit REDUX.MIN on the ordered image and one ballot instead of a 4-level shuffle tree of (double, index).
__device__ __forceinline__ int team_argmin(uint32_t tmask, int tbase, double v, double* vmin) {
  const uint64_t key = d2ord(v);
  const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
  const uint32_t mh = __reduce_min_sync(tmask, hi);
  const uint32_t ml = __reduce_min_sync(tmask, hi == mh ? lo : 0xFFFFFFFFu);
  const uint32_t win = __ballot_sync(tmask, hi == mh && lo == ml) & tmask;
  *vmin = ord2d(((uint64_t)mh << 32) | ml);
  return __ffs(win) - 1 - tbase;
}

template <bool kShaped>
__global__ void __launch_bounds__(kEncodeThreads)
encode_kernel(EncodeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int team = tid / kTeam, c = tid % kTeam;      // lane within the team = codebook centre
  const int tbase = (tid & 31) & ~(kTeam - 1);        // first lane of the team inside its warp
  const uint32_t tmask = 0xFFFFu << tbase;
  const uint32_t Dp = (a.D + 3) & ~3u;
  unsigned char* base = smem_raw + (size_t)team * encode_team_bytes(a.D, a.B, kShaped);
  float* res = reinterpret_cast<float*>(base);
  float* orig = res + Dp;
  double* n0 = reinterpret_cast<double*>(orig + Dp);
  double* dres = n0 + (kShaped ? a.B : 0);
  double* dorig = dres + (kShaped ? Dp : 0);
  uint8_t* code = reinterpret_cast<uint8_t*>(dorig + (kShaped ? Dp : 0));
  uint16_t* order = reinterpret_cast<uint16_t*>(code + ((a.B + 1) & ~1u));
  const uint64_t pair = (uint64_t)blockIdx.x * (kEncodeThreads / kTeam) + team;
  const bool active = pair < (uint64_t)a.n * a.npd;
  // inactive teams run the same instruction stream on a dummy pair so that the warp stays converged
  const uint32_t i = active ? (uint32_t)(pair / a.npd) : 0u;
  const uint32_t which = active ? (uint32_t)(pair % a.npd) : 0u;
  const int32_t tok = a.tokens[(size_t)i * a.npd + which];
  uint8_t* out = (which ? a.soar_codes : a.codes) + (size_t)i * a.B;
  if (tok < 0) {  // not spilled: the SOAR row stays zero (CombineLeafDatasets, tree_x_hybrid/internal/utils.h:87-106)
    if (active) for (uint32_t b = c; b < a.B; b += kTeam) out[b] = 0;
  }
  const bool work = tok >= 0;
  const float* x = a.x + (size_t)i * a.D;
  const float* cen = (a.centers && work) ? a.centers + (size_t)tok * a.D : nullptr;
  for (uint32_t k = c; k < a.D; k += kTeam) {
    const float v = x[k];
    orig[k] = v;
    const float r = cen ? __fsub_rn(v, __ldg(cen + k)) : v;   // ComputeResiduals: float subtraction
    res[k] = r;
    if constexpr (kShaped) { dres[k] = (double)r; dorig[k] = (double)v; }
  }
  __syncwarp();

  if constexpr (!kShaped) {
    for (uint32_t b = 0; b < a.B; ++b) {
      const uint32_t nd = (uint32_t)a.block_dims[b];
      const float* rb = res + a.block_off[b];
      const float* cx = a.codebook + ((size_t)b * 16 + c) * a.S;
      auto lq = [&](uint32_t k) { return rb[k]; };
      auto lx = [&](uint32_t k) { return __ldg(cx + k); };
      float dist;
      if (c < 15) dist = nd < 8 ? sql2_small(lq, lx, nd) : sql2_avx2_order(lq, lx, nd);
      else dist = sql2_sse4_order(lq, lx, nd);
      int idx = c;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(kFull, dist, o, kTeam);
        const int oi = __shfl_xor_sync(kFull, idx, o, kTeam);
        if (od < dist || (od == dist && oi < idx)) { dist = od; idx = oi; }
      }
      if (c == 0 && active && work) out[b] = (uint8_t)idx;
    }
    return;
  } else {
    const uint32_t B = a.B, D = a.D;
    // ---- ||orig|| (sequential double sum) and SquaredL2Norm(orig) (four strided accumulators) ----
    double chunked = 0.0, sqn = 0.0;
    if (c == 0) {
      for (uint32_t k = 0; k < D; ++k) chunked = __dadd_rn(chunked, __dmul_rn((double)orig[k], (double)orig[k]));
    } else if (c == 1) {  // DenseSingleAccumulate (utils/reduction.h:357-390)
      double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
      uint32_t k = 0;
      for (; k + 4 <= D; k += 4) {
        r0 = __dadd_rn(r0, __dmul_rn((double)orig[k], (double)orig[k]));
        r1 = __dadd_rn(r1, __dmul_rn((double)orig[k + 1], (double)orig[k + 1]));
        r2 = __dadd_rn(r2, __dmul_rn((double)orig[k + 2], (double)orig[k + 2]));
        r3 = __dadd_rn(r3, __dmul_rn((double)orig[k + 3], (double)orig[k + 3]));
      }
      r2 = __dadd_rn(r2, r3);
      if (k + 2 <= D) {
        r0 = __dadd_rn(r0, __dmul_rn((double)orig[k], (double)orig[k]));
        r1 = __dadd_rn(r1, __dmul_rn((double)orig[k + 1], (double)orig[k + 1]));
        k += 2;
      }
      r1 = __dadd_rn(r1, r2);
      if (k < D) r0 = __dadd_rn(r0, __dmul_rn((double)orig[k], (double)orig[k]));
      sqn = __dadd_rn(r0, r1);
    }
    chunked = __shfl_sync(kFull, chunked, 0, kTeam);
    sqn = __shfl_sync(kFull, sqn, 1, kTeam);
    const double inv_norm = __ddiv_rn(1.0, __dsqrt_rn(chunked));
    const double t2 = __dmul_rn(a.threshold, a.threshold);
    const double ratio = __ddiv_rn(t2, sqn);
    const double mult = __ddiv_rn(ratio, __ddiv_rn(__dsub_rn(1.0, ratio), __dsub_rn((double)D, 1.0)));
    // ComputeResidualStatsForCluster for (block b, this lane's centre).  The statistics are NOT kept: B * 16 pairs of
    // doubles per team would cap the SM at 8 resident warps; recomputing them costs 6 double operations per dimension
    // (the same operations in the same order, so the same bits) and keeps the team's footprint at ~1.4 KB.
    auto stats = [&](uint32_t b, double* rn_out, double* par_out) {
      const uint32_t nd = (uint32_t)a.block_dims[b], off = a.block_off[b];
      const float* cx = a.codebook + ((size_t)b * 16 + c) * a.S;
      double rn = 0.0, par = 0.0;
      for (uint32_t k = 0; k < nd; ++k) {
        const double rc = __dsub_rn(dres[off + k], (double)__ldg(cx + k));
        rn = __dadd_rn(rn, __dmul_rn(rc, rc));
        par = __dadd_rn(par, __dmul_rn(__dmul_rn(rc, dorig[off + k]), inv_norm));
      }
      *rn_out = rn;
      *par_out = par;
    };
    // ---- InitializeToMinResidualNorm (first minimum) + ComputeParallelResidualComponent (sequential over blocks) ----
    double par = 0.0;
    for (uint32_t b = 0; b < B; ++b) {
      double rn, pc, nmin;
      stats(b, &rn, &pc);
      const int idx = team_argmin(tmask, tbase, rn, &nmin);
      par = __dadd_rn(par, __shfl_sync(kFull, pc, idx, kTeam));
      if (c == 0) { code[b] = (uint8_t)idx; n0[b] = nmin; }
    }
    __syncwarp();
    // ---- visiting order: descending initial norm, ties by ascending block (rank by counting) ----
    bool tie = false;
    for (uint32_t b = c; b < B; b += kTeam) {
      const double v = n0[b];
      uint32_t rank = 0;
      for (uint32_t o = 0; o < B; ++o) {
        const double w = n0[o];
        rank += (w > v || (w == v && o < b)) ? 1u : 0u;
        tie |= (w == v && o != b);
      }
      order[rank] = (uint16_t)b;
    }
    __syncwarp();
    if (__ballot_sync(kFull, tie) & tmask)
      if (c == 0 && active && work) atomicAdd(a.ties, 1ull);
    // ---- coordinate descent (OptimizeSingleSubspace over the blocks, <= 10 rounds) ----
    bool changes = true;
    for (int round = 0; round < 10; ++round) {
      // the two teams of a warp stay in lockstep: a finished team keeps iterating without effect (a round without
      // changes is a fixed point), so the loop ends when neither team changed anything
      if (!__any_sync(kFull, changes)) break;
      changes = false;
      for (uint32_t ii = 0; ii < B; ++ii) {
        const uint32_t b = order[ii];
        const int cur = code[b];
        double rn, pc;
        stats(b, &rn, &pc);
        const double old_norm = __shfl_sync(kFull, rn, cur, kTeam), old_par = __shfl_sync(kFull, pc, cur, kTeam);
        const double new_par = __dadd_rn(__dsub_rn(par, old_par), pc);
        const double par_delta = __dsub_rn(__dmul_rn(new_par, new_par), __dmul_rn(par, par));
        const double norm_delta = __dsub_rn(rn, old_norm);
        const double perp_delta = __dsub_rn(norm_delta, par_delta);
        const double cost_delta = __dadd_rn(__dmul_rn(mult, par_delta), perp_delta);
        const bool valid = c != cur && !(par_delta > 0.0) && cost_delta < 0.0;
        if (__ballot_sync(kFull, valid) & tmask) {  // team-uniform
          double vmin;
          const int idx = team_argmin(tmask, tbase, valid ? cost_delta : 0.0, &vmin);  // valid lanes are < 0: they win
          par = __shfl_sync(tmask, new_par, tbase + idx);
          changes = true;
          if (c == 0) code[b] = (uint8_t)idx;
        }
      }
      __syncwarp();  // code[] of this round is visible to the next one (a block is visited once per round)
    }
    __syncwarp();
    if (active && work) for (uint32_t b = c; b < B; b += kTeam) out[b] = code[b];
  }
}

}  // namespace sb

// ---- host side: the C ABI entry point -----------------------------------------------------------
namespace {

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  sb::set_last_error(buf);
  return code;
}

#define CU(expr)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(SCANN_B200_INTERNAL, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),    \
                  __FILE__, __LINE__, cudaGetErrorString(_e));                                \
  } while (0)

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 16); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct Stream {
  cudaStream_t s = nullptr;
  ~Stream() { if (s) cudaStreamDestroy(s); }
};
struct Events {
  cudaEvent_t e[8] = {};
  ~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
};
struct HostPin {  // pins the caller's arrays for the duration of the call so that the chunk copies are DMA'd
  void* p = nullptr;
  ~HostPin() { if (p) cudaHostUnregister(p); }
  void pin(const void* q, size_t bytes, unsigned flags) {
    // Only large arrays: that is where DMA from the caller's pages pays, and a large numpy / malloc block is its own
    // mapping.  Small arrays share pages with unrelated heap objects; page-locking and unlocking those is avoided.
    if (bytes < ((size_t)32 << 20)) return;
    if (cudaHostRegister(const_cast<void*>(q), bytes, flags) == cudaSuccess) { p = const_cast<void*>(q); return; }
    (void)cudaGetLastError();
    if (flags != cudaHostRegisterDefault && cudaHostRegister(const_cast<void*>(q), bytes, cudaHostRegisterDefault) == cudaSuccess) {
      p = const_cast<void*>(q);
      return;
    }
    (void)cudaGetLastError();  // not fatal: the copies fall back to staged transfers
  }
};

}  // namespace

extern "C" int scann_b200_encode_database(const scann_b200_encode_desc* d, int32_t* tokens_out, uint8_t* codes_out,
                                          uint8_t* soar_codes_out, scann_b200_encode_stats* stats_out) {
  if (!d || !tokens_out || !codes_out) return fail(SCANN_B200_INVALID_ARGUMENT, "null argument");
  const uint32_t N = d->n, D = d->d, L = d->n_leaves, B = d->n_blocks, S = d->dims_per_block;
  if (!d->dataset || !d->centers || !d->codebook) return fail(SCANN_B200_INVALID_ARGUMENT, "dataset, centers and codebook are required");
  if (!D || !L || !B || !S) return fail(SCANN_B200_INVALID_ARGUMENT, "empty dimensionality / centres / codebook");
  if (B > 256) return fail(SCANN_B200_INVALID_ARGUMENT, "at most 256 AH blocks (got %u)", B);
  const bool soar = !std::isnan(d->soar_lambda);
  if (soar && !d->residual) return fail(SCANN_B200_INVALID_ARGUMENT, "SOAR requires residual quantization (dot product tree-AH)");
  if (soar && !soar_codes_out) return fail(SCANN_B200_INVALID_ARGUMENT, "soar_codes_out is required with SOAR");
  const bool shaped = !std::isnan(d->noise_shaping_threshold);
  std::vector<int32_t> bdims(B);
  std::vector<uint32_t> boff(B + 1, 0);
  for (uint32_t b = 0; b < B; ++b) {
    bdims[b] = d->block_dims ? d->block_dims[b] : (int32_t)S;
    if (bdims[b] <= 0 || (uint32_t)bdims[b] > S) return fail(SCANN_B200_INVALID_ARGUMENT, "block %u has %d dims (stride %u)", b, bdims[b], S);
    boff[b + 1] = boff[b] + (uint32_t)bdims[b];
  }
  if (boff[B] != D) return fail(SCANN_B200_INVALID_ARGUMENT, "AH blocks cover %u dims, dimensionality is %u", boff[B], D);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    (void)cudaGetLastError();
    return fail(SCANN_B200_FAILED_PRECONDITION, "no CUDA device: scann_b200 has no CPU path");
  }
  if (d->device < 0 || d->device >= ndev) return fail(SCANN_B200_INVALID_ARGUMENT, "device %d out of range", d->device);
  CU(cudaSetDevice(d->device));
  if (stats_out) memset(stats_out, 0, sizeof *stats_out);
  if (!N) return SCANN_B200_OK;

  Stream st;
  CU(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
  cudaStream_t s = st.s;
  Events ev;
  for (auto& e : ev.e) CU(cudaEventCreate(&e));

  // ---- centres, norms, bf16 operand, codebook ----
  Buf centers, cnorm, tok_b, codebook, block_dims, block_off;
  CU(centers.alloc(sizeof(float) * (size_t)L * D));
  CU(cudaMemcpyAsync(centers.p, d->centers, sizeof(float) * (size_t)L * D, cudaMemcpyHostToDevice, s));
  std::vector<float> cn(L);
  double cmax2 = 0.0;
  for (uint32_t l = 0; l < L; ++l) {
    // many_to_many_impl.inc:236-257: ||c||^2 = -(fnmadd chain over dims)
    float acc = 0.f;
    double a64 = 0.0;
    for (uint32_t k = 0; k < D; ++k) {
      const float c = d->centers[(size_t)l * D + k];
      acc = fmaf(-c, c, acc);
      a64 += (double)c * (double)c;
    }
    cn[l] = acc * -1.0f;
    cmax2 = std::max(cmax2, a64);
  }
  CU(cnorm.alloc(sizeof(float) * L));
  CU(cudaMemcpyAsync(cnorm.p, cn.data(), sizeof(float) * L, cudaMemcpyHostToDevice, s));
  CU(codebook.alloc(sizeof(float) * (size_t)B * 16 * S));
  CU(cudaMemcpyAsync(codebook.p, d->codebook, sizeof(float) * (size_t)B * 16 * S, cudaMemcpyHostToDevice, s));
  CU(block_dims.alloc(sizeof(int32_t) * B));
  CU(cudaMemcpyAsync(block_dims.p, bdims.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s));
  CU(block_off.alloc(sizeof(uint32_t) * (B + 1)));
  CU(cudaMemcpyAsync(block_off.p, boff.data(), sizeof(uint32_t) * (B + 1), cudaMemcpyHostToDevice, s));

  sb::DevIndex v{};
  v.distance = SCANN_B200_SQUARED_L2;  // the builder's partitioning distance (scann_builder.py:213-238)
  v.n = N; v.d = D; v.L = L;
  v.centers = centers.as<float>();
  v.center_sqnorm = cnorm.as<float>();
  v.center_max_norm = (float)(std::sqrt(cmax2) * 1.0001);
  v.tok_kp = sb::tokenize_kpitch(D);
  CU(tok_b.alloc(sb::tokenize_operand_bytes(L, D)));
  CU(sb::build_tokenize_operand(v.centers, L, D, 2, tok_b.p, s));
  v.tok_b = tok_b.p;

  // ---- per-chunk workspace ----
  const uint32_t P = soar ? std::min<uint32_t>(L, 8) : 1;
  const uint32_t npd = soar ? 2 : 1;
  uint32_t R = 16384;
  { const char* e = getenv("SCANN_B200_ENCODE_CHUNK"); if (e && atoi(e) > 0) R = (uint32_t)atoi(e); }
  R = std::min(R, N);
  Buf xbuf[2], dist, tok_a, near, bias, sec, toks[2], codes[2], scodes[2], counters, rhat, dist2, fbflag, tok_cmax;
  for (int b = 0; b < 2; ++b) {
    CU(xbuf[b].alloc(sizeof(float) * (size_t)R * D));
    CU(toks[b].alloc(sizeof(int32_t) * (size_t)R * npd));
    CU(codes[b].alloc((size_t)R * B));
    CU(scodes[b].alloc((size_t)R * B));
  }
  CU(dist.alloc(sizeof(float) * (size_t)R * L));
  CU(tok_a.alloc(sb::tokenize_operand_bytes(R, D)));
  CU(near.alloc(sizeof(int32_t) * (size_t)R * P));
  CU(bias.alloc(sizeof(float) * (size_t)R * P));
  CU(sec.alloc(sizeof(int32_t) * R));
  CU(fbflag.alloc(R));
  v.tok_fallback_flag = fbflag.as<uint8_t>();
  CU(tok_cmax.alloc(sizeof(float) * (size_t)R * ((L + 31) / 32)));
  v.tok_cmax_ws = tok_cmax.as<float>();
  v.tok_need_rows = soar ? 1 : 0;  // the SOAR pruning reads the distance matrix
  const bool row_is_dot = sb::tokenize_tensor_path(v, P);
  // SOAR: the projection term is pruned with a second tensor-core GEMM (rhat x centres) when the centre operand
  // exists; small trees (SIMT tokenization) prune by distance only
  bool soar_gemm = soar && row_is_dot && d->soar_lambda >= 0.f;
  { const char* e = getenv("SCANN_B200_SOAR_GEMM"); if (e && e[0] == '0') soar_gemm = false; }
  if (soar) CU(rhat.alloc(sizeof(float) * (size_t)R * D));
  if (soar_gemm) CU(dist2.alloc(sizeof(float) * (size_t)R * L));
  CU(counters.alloc(sizeof(unsigned long long) * 4 + sizeof(uint32_t) * 4));
  CU(cudaMemsetAsync(counters.p, 0, sizeof(unsigned long long) * 4 + sizeof(uint32_t) * 4, s));
  unsigned long long* c_evaluated = counters.as<unsigned long long>();
  unsigned long long* c_spilled = c_evaluated + 1;
  unsigned long long* c_ties = c_evaluated + 2;
  uint32_t* c_fallbacks = reinterpret_cast<uint32_t*>(c_evaluated + 4);

  HostPin pin_x, pin_t, pin_c, pin_s;
  { const char* e = getenv("SCANN_B200_ENCODE_PIN");
    if (!(e && e[0] == '0')) {
      pin_x.pin(d->dataset, sizeof(float) * (size_t)N * D, cudaHostRegisterReadOnly);
      pin_t.pin(tokens_out, sizeof(int32_t) * (size_t)N * npd, cudaHostRegisterDefault);
      pin_c.pin(codes_out, (size_t)N * B, cudaHostRegisterDefault);
      if (soar) pin_s.pin(soar_codes_out, (size_t)N * B, cudaHostRegisterDefault);
    } }

  const float eps_rel = (float)v.tok_kp * 4.76837158e-7f + 3.05175781e-5f;  // as launch_tokenize_topp
  const size_t team_bytes = sb::encode_team_bytes(D, B, shaped);
  const size_t enc_smem = team_bytes * (sb::kEncodeThreads / sb::kTeam);
  if (enc_smem > 200 * 1024) return fail(SCANN_B200_UNIMPLEMENTED, "encode: %zu bytes of shared memory per CTA (D = %u, B = %u)", enc_smem, D, B);
  if (shaped) CU(cudaFuncSetAttribute(sb::encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem));
  else CU(cudaFuncSetAttribute(sb::encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem));
  const size_t soar_smem = (size_t)(sb::kSoarThreads / 32) * (2 * ((D + 3) & ~3u) + sb::kSoarCand) * 4;
  if (soar) {
    if (soar_smem > 200 * 1024) return fail(SCANN_B200_UNIMPLEMENTED, "SOAR assignment: D = %u is too large", D);
    CU(cudaFuncSetAttribute(sb::soar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)soar_smem));
  }

  // ---- chunk pipeline: copy-in, compute and copy-out on three streams, two buffers ----
  // chunk c uses buffer b = c & 1.  copy-in of c waits for the compute of c - 2 (it overwrites x[b]); compute of c
  // waits for its copy-in and for the copy-out of c - 2 (it overwrites toks/codes[b]); copy-out of c waits for its
  // compute.  The host only blocks on the stage times of chunk c - 2.
  float ms[4] = {0, 0, 0, 0};  // tokenize, soar, encode, total
  Stream st_in, st_out;
  CU(cudaStreamCreateWithFlags(&st_in.s, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&st_out.s, cudaStreamNonBlocking));
  struct Slot {
    cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr, t[4] = {};
    bool used = false;
    ~Slot() {
      for (cudaEvent_t e : {in_done, comp_done, out_done, t[0], t[1], t[2], t[3]}) if (e) cudaEventDestroy(e);
    }
  } slot[2];
  for (auto& sl : slot) {
    CU(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&sl.comp_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
    for (auto& e : sl.t) CU(cudaEventCreate(&e));
  }
  auto collect = [&](Slot& sl) -> cudaError_t {  // stage times of the chunk that last used this slot
    cudaError_t e = cudaEventSynchronize(sl.t[3]);
    float t;
    for (int k = 0; k < 3 && e == cudaSuccess; ++k) {
      e = cudaEventElapsedTime(&t, sl.t[k], sl.t[k + 1]);
      ms[k] += t;
    }
    return e;
  };
  CU(cudaStreamSynchronize(s));  // centres, operand and codebook are in place before the other streams start
  CU(cudaEventRecord(ev.e[6], s));
  uint64_t chunk = 0;
  for (uint64_t r0 = 0; r0 < N; r0 += R, ++chunk) {
    const uint32_t nr = (uint32_t)std::min<uint64_t>(R, N - r0);
    const int b = (int)(chunk & 1);
    Slot& sl = slot[b];
    float* xb = xbuf[b].as<float>();
    int32_t* tb = toks[b].as<int32_t>();
    uint8_t* cb = codes[b].as<uint8_t>();
    uint8_t* sb_ = scodes[b].as<uint8_t>();
    if (sl.used) {
      CU(collect(sl));
      CU(cudaStreamWaitEvent(st_in.s, sl.comp_done, 0));
      CU(cudaStreamWaitEvent(s, sl.out_done, 0));
    }
    CU(cudaMemcpyAsync(xb, d->dataset + r0 * D, sizeof(float) * (size_t)nr * D, cudaMemcpyHostToDevice, st_in.s));
    CU(cudaEventRecord(sl.in_done, st_in.s));
    CU(cudaStreamWaitEvent(s, sl.in_done, 0));
    CU(cudaEventRecord(sl.t[0], s));
    CU(cudaMemsetAsync(fbflag.p, 0, nr, s));
    int launches = 0;
    CU(sb::launch_tokenize_topp(v, xb, nr, P, dist.as<float>(), tok_a.p, near.as<int32_t>(), bias.as<float>(),
                                c_fallbacks, s, &launches));
    CU(cudaEventRecord(sl.t[1], s));
    if (soar) {
      sb::SoarArgs a{};
      a.x = xb; a.centers = v.centers; a.cnorm = v.center_sqnorm; a.row = dist.as<float>();
      a.near = near.as<int32_t>(); a.sec = sec.as<int32_t>(); a.evaluated = c_evaluated;
      a.n = nr; a.L = L; a.D = D; a.P = P; a.row_is_dot = row_is_dot ? 1 : 0;
      a.lambda = d->soar_lambda; a.eps_rel = eps_rel; a.cmax = v.center_max_norm;
      a.rhat = rhat.as<float>();
      a.row_exact = fbflag.as<uint8_t>();
      const unsigned wgrid = (nr + sb::kSoarThreads / 32 - 1) / (sb::kSoarThreads / 32);
      sb::rhat_kernel<<<wgrid, sb::kSoarThreads, 0, s>>>(a);
      CU(cudaGetLastError());
      if (soar_gemm) {  // the tokenization's A operand is free again: reuse it for rhat
        CU(sb::build_tokenize_operand(a.rhat, nr, D, 1, tok_a.p, s));
        CU(sb::gemm_bf16_nt(tok_a.p, nr, (nr + 127) / 128 * 128, v.tok_b, L, v.tok_kp, dist2.as<float>(), L, s));
        a.row2 = dist2.as<float>();
      }
      sb::soar_kernel<<<wgrid, sb::kSoarThreads, soar_smem, s>>>(a);
      CU(cudaGetLastError());
    }
    sb::tokens_kernel<<<(nr + 255) / 256, 256, 0, s>>>(near.as<int32_t>(), P, soar ? sec.as<int32_t>() : nullptr, nr, tb,
                                                       c_spilled);
    CU(cudaGetLastError());
    CU(cudaEventRecord(sl.t[2], s));
    {
      sb::EncodeArgs a{};
      a.x = xb; a.centers = d->residual ? v.centers : nullptr; a.tokens = tb;
      a.codebook = codebook.as<float>(); a.block_dims = block_dims.as<int32_t>(); a.block_off = block_off.as<uint32_t>();
      a.codes = cb; a.soar_codes = sb_; a.ties = c_ties;
      a.n = nr; a.D = D; a.B = B; a.S = S; a.npd = npd; a.threshold = d->noise_shaping_threshold;
      const uint64_t pairs = (uint64_t)nr * npd;
      const unsigned grid = (unsigned)((pairs + sb::kEncodeThreads / sb::kTeam - 1) / (sb::kEncodeThreads / sb::kTeam));
      if (shaped) sb::encode_kernel<true><<<grid, sb::kEncodeThreads, enc_smem, s>>>(a);
      else sb::encode_kernel<false><<<grid, sb::kEncodeThreads, enc_smem, s>>>(a);
      CU(cudaGetLastError());
    }
    CU(cudaEventRecord(sl.t[3], s));
    CU(cudaEventRecord(sl.comp_done, s));
    CU(cudaStreamWaitEvent(st_out.s, sl.comp_done, 0));
    CU(cudaMemcpyAsync(tokens_out + r0 * npd, tb, sizeof(int32_t) * (size_t)nr * npd, cudaMemcpyDeviceToHost, st_out.s));
    CU(cudaMemcpyAsync(codes_out + r0 * B, cb, (size_t)nr * B, cudaMemcpyDeviceToHost, st_out.s));
    if (soar) CU(cudaMemcpyAsync(soar_codes_out + r0 * B, sb_, (size_t)nr * B, cudaMemcpyDeviceToHost, st_out.s));
    CU(cudaEventRecord(sl.out_done, st_out.s));
    sl.used = true;
  }
  for (auto& sl : slot) {
    if (!sl.used) continue;
    CU(collect(sl));
    CU(cudaStreamWaitEvent(s, sl.out_done, 0));
  }
  CU(cudaEventRecord(ev.e[7], s));
  CU(cudaStreamSynchronize(s));
  CU(cudaStreamSynchronize(st_out.s));
  CU(cudaEventElapsedTime(&ms[3], ev.e[6], ev.e[7]));
  if (stats_out) {
    unsigned long long hc[4];
    uint32_t hf[4];
    CU(cudaMemcpy(hc, counters.p, sizeof hc, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hf, c_fallbacks, sizeof hf, cudaMemcpyDeviceToHost));
    stats_out->ms_tokenize = ms[0];
    stats_out->ms_soar = ms[1];
    stats_out->ms_encode = ms[2];
    stats_out->ms_total = ms[3];
    stats_out->soar_evaluated = hc[0];
    stats_out->spilled = hc[1];
    stats_out->norm_ties = hc[2];
    stats_out->tokenize_fallbacks = hf[0];
    stats_out->chunk_rows = R;
  }
  return SCANN_B200_OK;
}
