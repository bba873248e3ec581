// oracle/_ref glue, part 7: the reference's own SSE4 one-to-one float kernels -- DenseDotProductSse4 and
// DenseSquaredL2DistanceSse4 (float, float) -- which the one-to-many dispatch uses for the last n mod 3 rows of a call
// (lambdas.VectorVector = DistanceMeasure::GetDistanceDense): in the AH lookup-table build that is codebook centre 15
// of every block (16 mod 3 = 1).
//
// Test infrastructure (see oracle/scann_oracle.h): only tests/ may load the resulting library.  This file contains no
// reference code; it INCLUDES two function definitions extracted by line range at BUILD time into oracle/_ref/gen/
// (git-ignored) by oracle/Makefile:
//   _ref/gen/sse4_dot_f32.inc = distance_measures/one_to_one/dot_product_sse4.cc:241-296
//   _ref/gen/sse4_l2_f32.inc  = distance_measures/one_to_one/l2_distance_sse4.cc:165-221
// Compiled with -ffp-contract=off: the vector steps of these functions are separate mul / add (sub) intrinsics, which
// the reference's clang build does not fuse and g++'s "fast" mode would.  Their LAST step for an odd number of
// dimensions is one scalar expression (`accumulator[0] += a * b`) that clang does fuse and "off" does not, so the entry
// points below refuse odd lengths: for those the oracle follows the contraction rule unpinned.
#include <immintrin.h>

#include "scann/utils/common.h"
#include "scann/utils/types.h"
#include "scann/utils/intrinsics/attributes.h"

namespace research_scann {

template <typename T>
class DatapointPtr {
 public:
  DatapointPtr(const T* values, size_t dims) : values_(values), d_(dims) {}
  const T* values() const { return values_; }
  size_t nonzero_entries() const { return d_; }
  bool IsDense() const { return true; }
 private:
  const T* values_;
  size_t d_;
};

namespace dp_internal {
#include "sse4_dot_f32.inc"
}  // namespace dp_internal
namespace l2_internal {
#include "sse4_l2_f32.inc"
}  // namespace l2_internal
}  // namespace research_scann

extern "C" {

// <a, b> / ||a - b||^2 as DistanceMeasure::GetDistanceDense computes them for float vectors (returned as double, as the
// reference does); n even.  Returns NaN for an odd n (see the header).
double ref_dot_sse4_f32(const float* a, const float* b, uint64_t n) {
  if (n & 1) return __builtin_nan("");
  using research_scann::DatapointPtr;
  return research_scann::dp_internal::DenseDotProductSse4(DatapointPtr<float>(a, n), DatapointPtr<float>(b, n));
}
double ref_sql2_sse4_f32(const float* a, const float* b, uint64_t n) {
  if (n & 1) return __builtin_nan("");
  using research_scann::DatapointPtr;
  return research_scann::l2_internal::DenseSquaredL2DistanceSse4(DatapointPtr<float>(a, n), DatapointPtr<float>(b, n));
}

}  // extern "C"
