// scan_common.cuh -- pieces shared by the SIMT scan (scan.cu) and the tensor-core scan (scan_tc.cu).
#pragma once
#include "common.cuh"
#include "exact_math.cuh"

namespace sb {

// ---- packed code layout -------------------------------------------------------------------
// Per 32-slot group: W 32-bit words per slot (nibble k of word j = code of block 8j+k), stored
// as planes so that a warp's loads are contiguous: floor(W/4) uint4 planes (512 B each), then
// one uint2 plane if W%4 >= 2, then one u32 plane if W is odd.  W*128 bytes per group.
template <int W>
__device__ __forceinline__ void load_codes(const uint32_t* __restrict__ gbase, int lane,
                                           uint32_t (&w)[W]) {
  constexpr int N4 = W / 4, R = W % 4;
#pragma unroll
  for (int p = 0; p < N4; ++p) {
    const uint4 v = ldg_stream_v4(gbase + p * 128 + lane * 4);
    w[4 * p + 0] = v.x; w[4 * p + 1] = v.y; w[4 * p + 2] = v.z; w[4 * p + 3] = v.w;
  }
  if constexpr (R >= 2) {
    const uint2 v = ldg_stream_v2(gbase + N4 * 128 + lane * 2);
    w[4 * N4 + 0] = v.x; w[4 * N4 + 1] = v.y;
  }
  if constexpr (R & 1) {
    w[W - 1] = ldg_stream_u32(gbase + N4 * 128 + ((R >= 2) ? 64 : 0) + lane);
  }
}

// Runtime-W variant (the generic kernels for 16 < W <= 32, i.e. 128 < B <= 256 blocks): word j of this lane.
__device__ __forceinline__ uint32_t load_code_word(const uint32_t* __restrict__ gbase, int lane, int W, int j) {
  const int N4 = W / 4, R = W % 4;
  if (j < 4 * N4) return __ldg(gbase + (j >> 2) * 128 + lane * 4 + (j & 3));
  if (R >= 2 && j < 4 * N4 + 2) return __ldg(gbase + N4 * 128 + lane * 2 + (j - 4 * N4));
  return __ldg(gbase + N4 * 128 + ((R >= 2) ? 64 : 0) + lane);
}

// Largest accumulator value whose float score is <= the score of `tau` (conservative integer
// pre-filter; the reference's trunc((eps - bias) * mult) of lut16_avx2.inc:432-438 may drop a
// candidate that is strictly better than eps, this one never does).
__device__ __forceinline__ int acc_threshold(uint64_t tau, float mult, float inv, float bias) {
  if (tau == kKeyMax) return 40000;
  const float ts = ord2f((uint32_t)(tau >> 32));
  const float est = __fmul_rn(__fsub_rn(ts, bias), mult);
  int t;
  if (!(est < 40000.f)) t = 32767;
  else if (!(est > -40000.f)) t = -32769;
  else t = (int)floorf(est);
  t = min(t, 32767);
  t = max(t, -32769);
  while (t < 32767 && ah_float_score(t + 1, inv, bias) <= ts) ++t;
  while (t >= -32768 && ah_float_score(t, inv, bias) > ts) --t;
  return t;
}

}  // namespace sb
