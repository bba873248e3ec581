// lut.cuh -- the per-entry arithmetic of the AH lookup table, shared by lut_kernel (prep.cu) and the fused
// LUT build in the pilot (scan.cu) so that both produce the same bits.
//   AsymmetricQueryer::CreateLookupTable (hashes/asymmetric_hashing2/querying.h:284-329),
//   AhImpl::CreateRawFloatLookupTable / ConvertLookupToFixedPoint<uint8_t>
//   (hashes/internal/asymmetric_hashing_impl.cc:505-645), paths relative to /root/reference/scann/.
#pragma once
#include <float.h>

#include "exact_math.cuh"
#include "kernels.h"

namespace sb {

// raw[b][c] = lookup_distance(q_block b, centre c): centres 0..14 through the accumulating one-to-many kernel
// (Highway lanes for dims < 8, AVX2 FMA lanes otherwise), centre 15 through the SSE4 one-to-one kernel
// (one_to_many_symmetric.h:704-705,793-799).  `sq` is the query in shared memory, e = b * 16 + c.
__device__ __forceinline__ float lut_raw_entry(const DevIndex& ix, const float* sq, uint32_t e) {
  const uint32_t b = e >> 4, c = e & 15;
  const uint32_t n = (uint32_t)ix.block_dims[b];
  const float* qb = sq + ix.block_off[b];
  const float* cx = ix.codebook + ((size_t)b * 16 + c) * ix.dpb;
  auto lq = [&](uint32_t i) { return qb[i]; };
  auto lx = [&](uint32_t i) { return cx[i]; };
  if (ix.distance == 0) {
    if (c < 15) return n < 8 ? neg_dot_small(lq, lx, n) : neg_dot_avx2_order(lq, lx, n);
    return -dot_sse4_order(lq, lx, n);
  }
  if (c < 15) return n < 8 ? sql2_small(lq, lx, n) : sql2_avx2_order(lq, lx, n);
  return sql2_sse4_order(lq, lx, n);
}

// mult = 127 / max(sqrt(FLT_EPSILON), max |raw|) (ComputeMultiplierByQuantile at quantile 1.0)
__device__ __forceinline__ float lut_multiplier(float max_abs_raw) {
  const float floor_ = __fsqrt_rn(FLT_EPSILON);
  const float denom = max_abs_raw > floor_ ? max_abs_raw : floor_;
  return __fdiv_rn(127.0f, denom);
}
// lut16_avx2.inc:429 (dot product: double division, narrowed) vs querying.h:450 (squared L2: 1.0f / mult)
__device__ __forceinline__ float lut_inverse_multiplier(const DevIndex& ix, float mult) {
  return ix.key_by_dp ? __fdiv_rn(1.0f, mult) : (float)(1.0 / (double)mult);
}
// u8(round_half_away(raw * mult) + 128)
__device__ __forceinline__ uint32_t lut_quantize(float raw, float mult) {
  return (uint32_t)(uint8_t)(int)__fadd_rn(roundf(__fmul_rn(raw, mult)), 128.0f);
}

}  // namespace sb
