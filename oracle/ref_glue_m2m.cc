// oracle/_ref glue, part 5: the arithmetic of the reference's batched float tokenization --
// DenseDistanceManyToMany (KMeansTreePartitioner::TokensForDatapointWithSpillingBatched, partitioning/
// kmeans_tree_partitioner.cc:642-730; the database tokenization of the index build takes the same kernel).
//
// Test infrastructure (see oracle/scann_oracle.h): only tests/ may load the resulting library.  This file contains no
// reference code; it INCLUDES two member functions of /root/reference/scann/distance_measures/many_to_many/
// many_to_many_impl.inc, extracted by line range at BUILD time into oracle/_ref/gen/ (git-ignored) by oracle/Makefile:
//   _ref/gen/m2m_augment.inc    = :236-257  M2MTransposer::AugmentWithL2Norms (|c|^2 = -(fnmadd chain), rows doubled)
//   _ref/gen/m2m_accumulate.inc = :522-560  DenseManyToManyTransposed::DoAccumulationTransposedTemplate
//   _ref/gen/m2m_oa_accumulate.inc = :729-773  DenseManyToManyOrthogonalityAmplified::DoAccumulationTransposedTemplate
//       (the SOAR cost t1 + (lambda t2) t2 of the index build's secondary assignment)
// and partitioning/orthogonality_amplification_utils.h:27-46 (ComputeNormalizedResidual) as _ref/gen/oa_residual.inc
// as static members of a struct that supplies what they name from their classes (kIsSquaredL2, FloatT,
// kElementsPerRegister), with the reference's own AVX2 wrappers (utils/intrinsics/avx2.h, fma.inc).  The transposition
// (many_to_many_impl.inc:169-207: element moves, no arithmetic) and the loop over blocks of 2 x 8 datapoints are written
// here; the query norm is SquaredL2Norm (ref_squared_l2_norm of ref_glue_sym.cc: the reference's DenseSingleAccumulate),
// narrowed to float as many_to_many_impl.inc:421-427 stores it.
#include <immintrin.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "scann/utils/common.h"
#include "scann/utils/types.h"
#include "scann/utils/index_sequence.h"
#include "scann/utils/intrinsics/fma.h"
#include "scann/utils/intrinsics/simd.h"

extern "C" double ref_squared_l2_norm(const float* v, uint64_t n);

namespace research_scann {
namespace avx2 {
#define SCANN_SIMD_ATTRIBUTE SCANN_AVX2

template <bool kIsSquaredL2, typename FloatT>
struct M2MPieces {
  static constexpr size_t kElementsPerRegister = Simd<FloatT>::kElementsPerRegister;
#include "m2m_augment.inc"
#include "m2m_accumulate.inc"
};

struct M2MOaPieces {
  using FloatT = float;
  static constexpr size_t kElementsPerRegister = Simd<float>::kElementsPerRegister;
#include "m2m_oa_accumulate.inc"
};

// one datapoint ("query") with its normalised residual against centres [0, n): out[i] = SOAR cost of centre i
SCANN_AVX2_OUTLINE void RunOneSoarQuery(const float* x, const float* rhat, float lambda, const float* centers, size_t n,
                                        size_t dims, float* out) {
  constexpr size_t kE = M2MOaPieces::kElementsPerRegister;
  const size_t tsz = dims * kE;
  float* storage = static_cast<float*>(aligned_alloc(64, (2 * tsz * sizeof(float) + 63) / 64 * 64));
  float* t0 = storage;
  float* t1 = storage + tsz;
  for (size_t first = 0; first < n; first += 2 * kE) {
    const size_t cnt = std::min(n - first, 2 * kE);
    for (size_t i = 0; i < 2 * tsz; ++i) storage[i] = 0.0f;
    for (size_t j = 0; j < cnt; ++j) {
      float* t = j < kE ? t0 : t1;
      for (size_t dim = 0; dim < dims; ++dim) t[dim * kE + (j % kE)] = centers[(first + j) * dims + dim];
    }
    const float* qptrs[1] = {x};
    const float* rptrs[1] = {rhat};
    auto acc = M2MOaPieces::DoAccumulationTransposedTemplate<1>(t0, t1, qptrs, rptrs, lambda, dims);
    auto results = acc.Store();
    for (size_t j = 0; j < cnt; ++j) out[first + j] = results[0].data()[j];
  }
  free(storage);
}

// one query against rows [0, n): out[i] = the accumulator of datapoint i
template <bool kIsSquaredL2>
SCANN_AVX2_OUTLINE void RunOneQuery(const float* query, const float* db, size_t n, size_t dims, float* out) {
  using P = M2MPieces<kIsSquaredL2, float>;
  constexpr size_t kE = P::kElementsPerRegister;  // 8
  const size_t tsz = (dims + (kIsSquaredL2 ? 1 : 0)) * kE;
  float* storage = static_cast<float*>(aligned_alloc(64, (2 * tsz * sizeof(float) + 63) / 64 * 64));
  float* t0 = storage + (kIsSquaredL2 ? kE : 0);
  float* t1 = storage + tsz + (kIsSquaredL2 ? kE : 0);
  const float qnorm = kIsSquaredL2 ? static_cast<float>(ref_squared_l2_norm(query, dims)) : 0.0f;
  for (size_t first = 0; first < n; first += 2 * kE) {
    const size_t cnt = std::min(n - first, 2 * kE);
    for (size_t i = 0; i < 2 * tsz; ++i) storage[i] = 0.0f;
    for (size_t j = 0; j < cnt; ++j) {
      float* t = j < kE ? t0 : t1;
      for (size_t dim = 0; dim < dims; ++dim) t[dim * kE + (j % kE)] = db[(first + j) * dims + dim];
    }
    if constexpr (kIsSquaredL2) P::AugmentWithL2Norms(t0, t1, dims);
    const float* qptrs[1] = {query};
    auto acc = P::template DoAccumulationTransposedTemplate<1>(t0, t1, qptrs, &qnorm, dims);
    auto results = acc.Store();
    for (size_t j = 0; j < cnt; ++j) out[first + j] = results[0].data()[j];
  }
  free(storage);
}

#undef SCANN_SIMD_ATTRIBUTE
}  // namespace avx2
}  // namespace research_scann

namespace research_scann {
template <typename T>
class DatapointPtr {
 public:
  DatapointPtr(const T* values, size_t dims) : values_(values), d_(dims) {}
  const T* values() const { return values_; }
  size_t dimensionality() const { return d_; }
 private:
  const T* values_;
  size_t d_;
};
#include "oa_residual.inc"  // _ref/gen: partitioning/orthogonality_amplification_utils.h:27-46
}  // namespace research_scann

extern "C" {

// OrthogonalityAmplifiedTokenForDatapointBatched's arithmetic (partitioning/kmeans_tree_partitioner.cc:925-997): for every
// datapoint i, rhat = ComputeNormalizedResidual(x_i, centre primary[i]) -> out_rhat [n][dims] (optional), and the SOAR
// cost of every centre -> out_cost [n][L]
int ref_soar_costs(const float* x, uint64_t n, uint64_t dims, const float* centers, uint64_t L, const int32_t* primary,
                   float lambda, float* out_cost, float* out_rhat) {
  using namespace research_scann;
  std::vector<float> rhat(dims);
  for (uint64_t i = 0; i < n; ++i) {
    ComputeNormalizedResidual(DatapointPtr<float>(x + i * dims, dims),
                              DatapointPtr<float>(centers + (uint64_t)primary[i] * dims, dims),
                              MutableSpan<float>(rhat.data(), dims));
    if (out_rhat) memcpy(out_rhat + i * dims, rhat.data(), sizeof(float) * dims);
    avx2::RunOneSoarQuery(x + i * dims, rhat.data(), lambda, centers, L, dims, out_cost + i * L);
  }
  return 0;
}

// out[q][i] = DenseDistanceManyToMany(dot product | squared L2)(queries, db) as the reference accumulates it
int ref_many_to_many_f32(const float* queries, uint64_t nq, const float* db, uint64_t n, uint64_t dims, int squared_l2,
                         float* out) {
  for (uint64_t q = 0; q < nq; ++q) {
    if (squared_l2) research_scann::avx2::RunOneQuery<true>(queries + q * dims, db, n, dims, out + q * n);
    else research_scann::avx2::RunOneQuery<false>(queries + q * dims, db, n, dims, out + q * n);
  }
  return 0;
}

}  // extern "C"
