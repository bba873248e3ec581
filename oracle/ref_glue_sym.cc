// oracle/_ref glue, part 4: the reference's own SYMMETRIC float one-to-many kernel (f32 query x f32 rows, dims >= 8) --
// the arithmetic of the exact f32 reordering (utils/reordering_helper.cc ExactReorderingHelper -> DenseDotProductDistance
// OneToMany / DenseSquaredL2DistanceOneToMany) and of the AH lookup-table build for blocks of >= 8 dims.
//
// Test infrastructure (see oracle/scann_oracle.h): only tests/ may load the resulting library.  This file contains no
// reference code; it INCLUDES line ranges of /root/reference/scann/distance_measures/one_to_many/one_to_many_symmetric.h,
// extracted at BUILD time into oracle/_ref/gen/ (git-ignored) by oracle/Makefile:
//   _ref/gen/sym_sumtopbottom.inc = :233-237    SumTopBottomAvx
//   _ref/gen/sym_avx2.inc         = :373-503    DenseAccumulatingDistanceMeasureOneToManyInternalAvx2
//   _ref/gen/sym_lambdas.inc      = :983-1085   DotProductDistanceLambdas, SquaredL2DistanceLambdas
// and, for SquaredL2Norm (query norms of squared-L2 tokenization, dp_norms.npy, the centre norms of int8 tokenization):
//   _ref/gen/reduction_dense_single.inc = utils/reduction.h:357-390                         DenseSingleAccumulate
//   _ref/gen/l2_square.inc              = distance_measures/one_to_one/l2_distance.h:55-62  struct Square
// What this file supplies are the non-arithmetic names those ranges lean on and that live in headers this image cannot
// compile (abseil, Highway, DenseDataset, the thread pool): DatapointPtr (pointer + length), a row-major view, a serial
// ParallelFor, the Highway vector NAMES the lambda classes mention in templates that are never instantiated here, and
// distance classes whose one-to-one GetDistanceDense (the kernel of the last n mod 3 rows) is never called: the entry
// point below takes 3 m rows.
#include <immintrin.h>

#include <cmath>
#include <cstring>

#include "scann/utils/common.h"
#include "scann/utils/types.h"
#include "scann/utils/intrinsics/attributes.h"

namespace hwy {
namespace ref_shim_ns {
template <typename D> struct VecShim {};
template <typename D> using Vec = VecShim<D>;
template <typename V> V NegMulAdd(V, V, V);
template <typename V> V MulAdd(V, V, V);
}  // namespace ref_shim_ns
}  // namespace hwy
#define HWY_NAMESPACE ref_shim_ns

namespace research_scann {

class ThreadPool;

template <typename T>
class DatapointPtr {
 public:
  DatapointPtr(const T* values, size_t dims) : values_(values), d_(dims) {}
  const T* values() const { return values_; }
  size_t dimensionality() const { return d_; }
  size_t nonzero_entries() const { return d_; }
 private:
  const T* values_;
  size_t d_;
};
template <typename T>
DatapointPtr<T> MakeDatapointPtr(const T* values, size_t dims) { return DatapointPtr<T>(values, dims); }

struct DotProductDistance {
  double GetDistanceDense(const DatapointPtr<float>&, const DatapointPtr<float>&) const { return std::nan(""); }
};
struct SquaredL2Distance {
  double GetDistanceDense(const DatapointPtr<float>&, const DatapointPtr<float>&) const { return std::nan(""); }
};

template <size_t kBlock, typename SeqT, typename F>
SCANN_INLINE void ParallelFor(SeqT seq, ThreadPool*, F f) {
  for (size_t i : seq) f(i);
}

struct RowMajorViewF {
  const float* base;
  size_t dims;
  const float* GetPtr(size_t i) const { return base + i * dims; }
};

namespace one_to_many_low_level {

template <typename ValueT>
inline size_t GetDatapointIndex(MutableSpan<ValueT>, size_t index) { return index; }

struct StoreFloatIndexed {
  float* out;
  SCANN_INLINE void invoke(size_t index, float val) const { out[index] = val; }
  SCANN_INLINE void invoke(size_t index, double val) const { out[index] = (float)val; }
  SCANN_INLINE void prefetch(size_t) const {}
};

#include "sym_sumtopbottom.inc"
#include "sym_avx2.inc"
#include "sym_lambdas.inc"

}  // namespace one_to_many_low_level
}  // namespace research_scann

namespace research_scann {
// utils/reduction.h: AccumulatorTypeFor<float> is double (types.h AccumulatorTypeFor: floating point -> double)
template <typename T> using AccumulatorTypeFor = double;
#include "reduction_dense_single.inc"  // _ref/gen: utils/reduction.h:357-390 DenseSingleAccumulate
namespace l2_distance_internal {
#include "l2_square.inc"               // _ref/gen: distance_measures/one_to_one/l2_distance.h:55-62 struct Square
}  // namespace l2_distance_internal
}  // namespace research_scann

using namespace research_scann;
using namespace research_scann::one_to_many_low_level;

extern "C" {

// DenseAccumulatingDistanceMeasureOneToManyInternalAvx2 over rows [0, n) of a row-major f32 matrix, n a multiple of 3,
// dims >= 8: out[i] = -<query, row i> (dot product) or ||query - row i||^2 (squared L2).  Returns 1 on bad arguments.
int ref_one_to_many_f32(const float* query, const float* rows, uint64_t n, uint64_t dims, int squared_l2, float* out) {
  if (dims < 8 || n % 3 != 0) return 1;
  DatapointPtr<float> q(query, (size_t)dims);
  RowMajorViewF view{rows, (size_t)dims};
  StoreFloatIndexed cb{out};
  MutableSpan<float> result(out, (size_t)n);
  if (squared_l2) {
    SquaredL2DistanceLambdas<float> lambdas;
    DenseAccumulatingDistanceMeasureOneToManyInternalAvx2<float, RowMajorViewF, SquaredL2DistanceLambdas<float>, float,
                                                          false, StoreFloatIndexed>(q, &view, lambdas, result, &cb, nullptr);
  } else {
    DotProductDistanceLambdas<float> lambdas;
    DenseAccumulatingDistanceMeasureOneToManyInternalAvx2<float, RowMajorViewF, DotProductDistanceLambdas<float>, float,
                                                          false, StoreFloatIndexed>(q, &view, lambdas, result, &cb, nullptr);
  }
  return 0;
}

// SquaredL2Norm(ConstSpan<float>) (distance_measures/one_to_one/l2_distance.h:108-111): DenseSingleAccumulate(vec, Square())
double ref_squared_l2_norm(const float* v, uint64_t n) {
  return DenseSingleAccumulate(ConstSpan<float>(v, (size_t)n), l2_distance_internal::Square());
}

}  // extern "C"
