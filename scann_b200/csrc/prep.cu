// prep.cu -- query preparation kernels: partition tokenization and AH lookup tables.
//
// (a3/a4) KMeansTreePartitioner::TokensForDatapointWithSpillingBatched
//         partitioning/kmeans_tree_partitioner.cc:642-730, many_to_many_impl.inc:522-567
// (a5)    AsymmetricQueryer::CreateLookupTable hashes/asymmetric_hashing2/querying.h:284-329,
//         hashes/internal/asymmetric_hashing_impl.cc:505-645
#include <float.h>

#include "common.cuh"
#include "exact_math.cuh"
#include "kernels.h"

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
    // acc = ||c||^2 + ||q||^2 (many_to_many_impl.inc:530-541); ||q||^2 accumulated in double.
    if (tid < TM) {
      double s = 0.0;
      const int r = m0 + tid;
      if (r < nq)
        for (int k = 0; k < D; ++k) s += (double)q[(size_t)r * D + k] * (double)q[(size_t)r * D + k];
      qn[tid] = (float)s;
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

void launch_tokenize(const DevIndex& ix, const float* q, uint32_t nq, float* dist, cudaStream_t s) {
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

__global__ void __launch_bounds__(256)
topp_warp_kernel(const float* __restrict__ dist, int nq, int L, int P, int Ppow2,
                 int32_t* __restrict__ leaves, float* __restrict__ bias) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 8 + warp;
  if (q >= nq) return;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw) + warp * 256;
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw + 8 * 256 * 4) + (size_t)warp * Ppow2;
  const float* row = dist + (size_t)q * L;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t prefix = 0, mask = 0, need = (uint32_t)P;
  if (P < L) {
    for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) hist[lane * 8 + k] = 0;
      __syncwarp();
      for (int i = lane; i < L; i += 32) {
        const uint32_t o = f2ord(row[i]);
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
  } else {
    prefix = 0xFFFFFFFFu;
    need = 0xFFFFFFFFu;
  }
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
  for (int k = 2; k <= Ppow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (Ppow2 >> 1); t += 32) {
        const int l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int r = l | j;
        const uint64_t a = skeys[l], b = skeys[r];
        const bool up = (l & k) == 0;
        if ((a > b) == up) { skeys[l] = b; skeys[r] = a; }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < P; i += 32) {
    const uint64_t k = skeys[i];
    const uint32_t l = (uint32_t)k;
    leaves[(size_t)q * P + i] = (k == kKeyMax) ? -1 : (int32_t)l;
    bias[(size_t)q * P + i] = (k == kKeyMax) ? 0.f : row[l];
  }
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
    const uint32_t b = e >> 4, c = e & 15;
    const uint32_t n = (uint32_t)ix.block_dims[b];
    const float* qb = sq + ix.block_off[b];
    const float* cx = ix.codebook + ((size_t)b * 16 + c) * ix.dpb;
    auto lq = [&](uint32_t i) { return qb[i]; };
    auto lx = [&](uint32_t i) { return cx[i]; };
    float r;
    if (ix.distance == 0) {
      if (c < 15) r = n < 8 ? neg_dot_small(lq, lx, n) : neg_dot_avx2_order(lq, lx, n);
      else r = -dot_sse4_order(lq, lx, n);
    } else {
      if (c < 15) r = n < 8 ? sql2_small(lq, lx, n) : sql2_avx2_order(lq, lx, n);
      else r = sql2_sse4_order(lq, lx, n);
    }
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
    const float floor_ = __fsqrt_rn(FLT_EPSILON);
    const float denom = m > floor_ ? m : floor_;
    const float mult = __fdiv_rn(127.0f, denom);
    s_mult = mult;
    mult_out[qi] = mult;
    // lut16_avx2.inc:429 (dot, double division) vs querying.h:450 (squared L2, 1.0f / mult)
    inv_out[qi] = ix.key_by_dp ? __fdiv_rn(1.0f, mult) : (float)(1.0 / (double)mult);
  }
  __syncthreads();
  const float mult = s_mult;
  uint8_t* out = lut + (size_t)qi * ix.W * 8 * 16;
  const uint32_t npad = ix.W * 8 * 16;
  for (uint32_t e = tid; e < npad; e += kLutThreads) {
    uint8_t v = 0;
    if (e < ne) {
      const float f = __fadd_rn(roundf(__fmul_rn(raw[e], mult)), 128.0f);
      v = (uint8_t)(int)f;
    }
    out[e] = v;
  }
}

void launch_lut(const DevIndex& ix, const float* q, uint32_t nq, uint8_t* lut, float* mult,
                float* inv_mult, cudaStream_t s) {
  const size_t smem = (((size_t)ix.d + 3) & ~(size_t)3) * 4 + (size_t)ix.B * 16 * 4;
  lut_kernel<<<nq, kLutThreads, smem, s>>>(ix, q, lut, mult, inv_mult);
}

}  // namespace sb
