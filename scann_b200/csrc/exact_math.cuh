// exact_math.cuh -- float kernels whose rounding sequence is part of the parity contract.
//
// Each function reproduces, operation by operation, the arithmetic of one CPU kernel of the
// reference so that LUT bytes, centre distances and reorder distances are bit-identical to
// the oracle (oracle/scann_oracle.c holds the same restatements for the CPU).  All float
// operations are written with explicit round-to-nearest intrinsics so that nvcc can neither
// contract a mul+add into an FMA nor split an FMA.
//
// Paths are relative to /root/reference/scann/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

// float(SquaredL2Norm(v)) (distance_measures/one_to_one/l2_distance.h:108-120 -> DenseSingleAccumulate, utils/reduction.h:
// 357-390): four strided double accumulators, folded r2 += r3, a 2-wide step, r1 += r2, the last element, r0 + r1.  The
// squares of floats are exact in double, so fusing or not cannot matter.  Pinned to the reference's compiled
// DenseSingleAccumulate through the oracle (tests/test_oracle_ref.py).
__device__ __forceinline__ float squared_l2_norm_strided(const float* __restrict__ v, uint32_t D) {
  double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  uint32_t k = 0;
  for (; k + 4 <= D; k += 4) {
    r0 = __dadd_rn(r0, __dmul_rn((double)v[k], (double)v[k]));
    r1 = __dadd_rn(r1, __dmul_rn((double)v[k + 1], (double)v[k + 1]));
    r2 = __dadd_rn(r2, __dmul_rn((double)v[k + 2], (double)v[k + 2]));
    r3 = __dadd_rn(r3, __dmul_rn((double)v[k + 3], (double)v[k + 3]));
  }
  r2 = __dadd_rn(r2, r3);
  if (k + 2 <= D) {
    r0 = __dadd_rn(r0, __dmul_rn((double)v[k], (double)v[k]));
    r1 = __dadd_rn(r1, __dmul_rn((double)v[k + 1], (double)v[k + 1]));
    k += 2;
  }
  r1 = __dadd_rn(r1, r2);
  if (k < D) r0 = __dadd_rn(r0, __dmul_rn((double)v[k], (double)v[k]));
  return (float)__dadd_rn(r0, r1);
}

// distance_measures/one_to_many/one_to_many_symmetric.h:373-503 (AVX2 one-to-many, dims >= 8):
// eight fnmadd lanes, top+bottom fold, 4-wide and 2-wide steps, (x0+x2)+(x1+x3), fused tail.
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float neg_dot_avx2_order(LoadQ q, LoadX x, uint32_t n) {
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) a[l] = __fmaf_rn(-q(j + l), x(j + l), a[l]);
  }
  float b[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) b[l] = __fadd_rn(a[l + 4], a[l]);
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) b[l] = __fmaf_rn(-q(j + l), x(j + l), b[l]);
    j += 4;
  }
  if (j + 2 <= n) {
    b[2] = __fmaf_rn(-q(j), x(j), b[2]);
    b[3] = __fmaf_rn(-q(j + 1), x(j + 1), b[3]);
    j += 2;
  }
  float r = __fadd_rn(__fadd_rn(b[0], b[2]), __fadd_rn(b[1], b[3]));
  if (j < n) r = __fmaf_rn(-q(j), x(j), r);
  return r;
}

// distance_measures/one_to_many/one_to_many_asymmetric_impl.inc:296-353 (OneToManyAsymmetricTemplate<.., int16_t> /
// <.., int8_t> on AVX2: f32 query x bf16 or int8 row): eight fnmadd lanes over whole groups of 8 dims, one 4-wide step
// into lanes 0..3, HorizontalSum3X = ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)), the remaining dims fused on the scalar.
// Pinned to the reference's compiled kernel through the oracle (tests/test_oracle_ref.py).
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float neg_dot_asym_order(LoadQ q, LoadX x, uint32_t n) {
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) a[l] = __fmaf_rn(-q(j + l), x(j + l), a[l]);
  }
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = __fmaf_rn(-q(j + l), x(j + l), a[l]);
    j += 4;
  }
  float r = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[4]), __fadd_rn(a[2], a[6])),
                      __fadd_rn(__fadd_rn(a[1], a[5]), __fadd_rn(a[3], a[7])));
  for (; j < n; ++j) r = __fmaf_rn(-q(j), x(j), r);
  return r;
}

// FusedMultiplyOp<kIsSquaredL2 = true, int16_t> (:102-116): diff = q - x; acc = fma(diff, diff, acc), same lanes
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float sql2_asym_order(LoadQ q, LoadX x, uint32_t n) {
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float t = __fsub_rn(q(j + l), x(j + l));
      a[l] = __fmaf_rn(t, t, a[l]);
    }
  }
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float t = __fsub_rn(q(j + l), x(j + l));
      a[l] = __fmaf_rn(t, t, a[l]);
    }
    j += 4;
  }
  float r = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[4]), __fadd_rn(a[2], a[6])),
                      __fadd_rn(__fadd_rn(a[1], a[5]), __fadd_rn(a[3], a[7])));
  for (; j < n; ++j) {
    const float t = __fsub_rn(q(j), x(j));
    r = __fmaf_rn(t, t, r);
  }
  return r;
}

// -<q', float(c)> for one int8 row in the order of the reference's ONE-TO-ONE kernel, DenseDotProductInt8FloatAvxImpl
// (distance_measures/one_to_one/dot_product_impl.inc:3-54; the last n mod 3 rows of an asymmetric one-to-many call):
// two 8-lane fmadd accumulators over whole groups of 16 dims, one 8-wide fmadd step into the first, one 4-wide step as
// a rounded product ADDED to lanes 0..3 of the first, Sum8(acc0 + acc1) = ((x0+x4)+(x2+x6)) + ((x1+x5)+(x3+x7)), the
// last < 4 dims fused on the scalar; negated at the end.
template <typename LoadQ>
__device__ __forceinline__ float neg_dot_i8_one_to_one(LoadQ q, const int8_t* __restrict__ x, uint32_t n) {
  float a0[8], a1[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) a0[l] = a1[l] = 0.f;
  uint32_t j = 0;
  for (; j + 16 <= n; j += 16) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      a0[l] = __fmaf_rn((float)x[j + l], q(j + l), a0[l]);
      a1[l] = __fmaf_rn((float)x[j + 8 + l], q(j + 8 + l), a1[l]);
    }
  }
  if (j + 8 <= n) {
#pragma unroll
    for (int l = 0; l < 8; ++l) a0[l] = __fmaf_rn((float)x[j + l], q(j + l), a0[l]);
    j += 8;
  }
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) a0[l] = __fadd_rn(a0[l], __fmul_rn((float)x[j + l], q(j + l)));
    j += 4;
  }
  float v[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) v[l] = __fadd_rn(a0[l], a1[l]);
  float s = __fadd_rn(__fadd_rn(__fadd_rn(v[0], v[4]), __fadd_rn(v[2], v[6])),
                      __fadd_rn(__fadd_rn(v[1], v[5]), __fadd_rn(v[3], v[7])));
  for (; j < n; ++j) s = __fmaf_rn((float)x[j], q(j), s);
  return -s;
}

// SquaredL2DistanceLambdas::FmaTerm (one_to_many_symmetric.h:1043-1051).
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float sql2_avx2_order(LoadQ q, LoadX x, uint32_t n) {
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float t = __fsub_rn(q(j + l), x(j + l));
      a[l] = __fmaf_rn(t, t, a[l]);
    }
  }
  float b[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) b[l] = __fadd_rn(a[l + 4], a[l]);
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float t = __fsub_rn(q(j + l), x(j + l));
      b[l] = __fmaf_rn(t, t, b[l]);
    }
    j += 4;
  }
  if (j + 2 <= n) {
    const float t2 = __fsub_rn(q(j), x(j)), t3 = __fsub_rn(q(j + 1), x(j + 1));
    b[2] = __fmaf_rn(t2, t2, b[2]);
    b[3] = __fmaf_rn(t3, t3, b[3]);
    j += 2;
  }
  float r = __fadd_rn(__fadd_rn(b[0], b[2]), __fadd_rn(b[1], b[3]));
  if (j < n) {
    const float t = __fsub_rn(q(j), x(j));
    r = __fmaf_rn(t, t, r);
  }
  return r;
}

// one_to_many_symmetric.h:691-800 (Highway path for dims < 8, 4 lanes, no FMA).
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float neg_dot_small(LoadQ q, LoadX x, uint32_t n) {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 4 <= n; j += 4) {
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = __fsub_rn(a[l], __fmul_rn(q(j + l), x(j + l)));
  }
  if (j + 2 <= n) {
    a[0] = __fsub_rn(a[0], __fmul_rn(q(j), x(j)));
    a[1] = __fsub_rn(a[1], __fmul_rn(q(j + 1), x(j + 1)));
    j += 2;
  }
  float r = __fadd_rn(__fadd_rn(a[0], a[2]), __fadd_rn(a[1], a[3]));
  if (j < n) r = __fmaf_rn(-q(j), x(j), r);  // the scalar AccTerm `acc - a * b`: one expression, fused by the reference's build
  return r;
}
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float sql2_small(LoadQ q, LoadX x, uint32_t n) {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  for (; j + 4 <= n; j += 4) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float t = __fsub_rn(q(j + l), x(j + l));
      a[l] = __fadd_rn(a[l], __fmul_rn(t, t));
    }
  }
  if (j + 2 <= n) {
    const float t0 = __fsub_rn(q(j), x(j)), t1 = __fsub_rn(q(j + 1), x(j + 1));
    a[0] = __fadd_rn(a[0], __fmul_rn(t0, t0));
    a[1] = __fadd_rn(a[1], __fmul_rn(t1, t1));
    j += 2;
  }
  float r = __fadd_rn(__fadd_rn(a[0], a[2]), __fadd_rn(a[1], a[3]));
  if (j < n) {
    const float t = __fsub_rn(q(j), x(j));
    r = __fmaf_rn(t, t, r);  // scalar AccTerm `acc + tmp * tmp`: fused
  }
  return r;
}

// distance_measures/one_to_one/dot_product_sse4.cc:242-296 (DenseDotProductSse4, float).
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float dot_sse4_order(LoadQ q, LoadX x, uint32_t n) {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  if (n >= 8) {
    float a0[4], a1[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      a0[l] = __fmul_rn(q(l), x(l));
      a1[l] = __fmul_rn(q(4 + l), x(4 + l));
    }
    j = 8;
    for (; j + 8 <= n; j += 8) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        a0[l] = __fadd_rn(a0[l], __fmul_rn(q(j + l), x(j + l)));
        a1[l] = __fadd_rn(a1[l], __fmul_rn(q(j + 4 + l), x(j + 4 + l)));
      }
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = __fadd_rn(a0[l], a1[l]);
  }
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = __fadd_rn(a[l], __fmul_rn(q(j + l), x(j + l)));
    j += 4;
  }
  if (j + 2 <= n) {
    a[0] = __fadd_rn(a[0], 0.0f);
    a[1] = __fadd_rn(a[1], 0.0f);
    a[2] = __fadd_rn(a[2], __fmul_rn(q(j), x(j)));
    a[3] = __fadd_rn(a[3], __fmul_rn(q(j + 1), x(j + 1)));
    j += 2;
  }
  if (j < n) a[0] = __fmaf_rn(q(j), x(j), a[0]);  // `accumulator[0] += aptr[0] * bptr[0]`: fused
  return __fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3]));
}
template <typename LoadQ, typename LoadX>
__device__ __forceinline__ float sql2_sse4_order(LoadQ q, LoadX x, uint32_t n) {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t j = 0;
  if (n >= 8) {
    float a0[4], a1[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float t0 = __fsub_rn(q(l), x(l)), t1 = __fsub_rn(q(4 + l), x(4 + l));
      a0[l] = __fmul_rn(t0, t0);
      a1[l] = __fmul_rn(t1, t1);
    }
    j = 8;
    for (; j + 8 <= n; j += 8) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float t0 = __fsub_rn(q(j + l), x(j + l)), t1 = __fsub_rn(q(j + 4 + l), x(j + 4 + l));
        a0[l] = __fadd_rn(a0[l], __fmul_rn(t0, t0));
        a1[l] = __fadd_rn(a1[l], __fmul_rn(t1, t1));
      }
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) a[l] = __fadd_rn(a0[l], a1[l]);
  }
  if (j + 4 <= n) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float t = __fsub_rn(q(j + l), x(j + l));
      a[l] = __fadd_rn(a[l], __fmul_rn(t, t));
    }
    j += 4;
  }
  if (j + 2 <= n) {
    const float t2 = __fsub_rn(q(j), x(j)), t3 = __fsub_rn(q(j + 1), x(j + 1));
    a[2] = __fadd_rn(a[2], __fmul_rn(t2, t2));
    a[3] = __fadd_rn(a[3], __fmul_rn(t3, t3));
    j += 2;
  }
  if (j < n) {
    const float t = __fsub_rn(q(j), x(j));
    a[0] = __fmaf_rn(t, t, a[0]);  // `accumulator[0] += (a - b) * (a - b)`: fused
  }
  return __fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3]));
}

// lut16_avx2.inc:429,472-476: dist = float(acc) * float(1.0 / mult) + bias, two roundings.
__device__ __forceinline__ float ah_float_score(int acc, float inv_mult, float bias) {
  return __fadd_rn(__fmul_rn((float)acc, inv_mult), bias);
}

}  // namespace sb
