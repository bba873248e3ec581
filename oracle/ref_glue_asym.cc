// oracle/_ref glue, part 3: the reference's own asymmetric one-to-many kernels (f32 query x int8 / bf16 rows), compiled
// from where they lie -- the arithmetic behind int8 (FIXED_POINT_INT8) query tokenization
// (KMeansTreeNode::GetAllDistancesInt8, trees/kmeans_tree/kmeans_tree_node.h:247) and the int8 / bfloat16 reordering
// helpers (utils/reordering_helper.cc:430-441,610-618,745-757).
//
// Test infrastructure (see oracle/scann_oracle.h): only tests/ may load the resulting library.  This file contains no
// reference code; it INCLUDES
//   /root/reference/scann/distance_measures/one_to_many/one_to_many_asymmetric_impl.inc  (whole file, as the AVX2
//       instantiation of one_to_many_asymmetric.h:118-122 does: OneToManyAsymmetricTemplate, ComputeOneToManyScores,
//       ComputeOneToOneScore, HandleXDims, ...)
//   /root/reference/scann/distance_measures/one_to_one/dot_product_impl.inc              (DenseDotProductInt8FloatAvxImpl,
//       the one-to-one kernel of the last n mod 3 rows)
//   and through them utils/intrinsics/{attributes,sse4,avx1,avx2,fma,horizontal_sum}.h, utils/index_sequence.h,
//   utils/internal/{avx_funcs,avx2_funcs}.h, and utils/bfloat16_helpers.h:30-48 (extracted by line range into _ref/gen)
// against the shim headers of oracle/ref_shim/.  What this file supplies are the non-arithmetic names those files lean
// on and that live in headers this image cannot compile (protobuf, abseil, DenseDataset): DatapointPtr (a pointer +
// length), a row-major dataset view, the result callbacks of one_to_many_helpers.h:32-48,72-83,192-218 (index = position,
// store the value), and declarations of the SSE4 / AVX1 one-to-one kernels the AVX2 build names but never calls.
#include <immintrin.h>

#include <cstring>

#include "absl/base/casts.h"
#include "scann/utils/common.h"
#include "scann/utils/types.h"
#include "scann/utils/index_sequence.h"
#include "scann/utils/intrinsics/fma.h"
#include "scann/utils/intrinsics/horizontal_sum.h"
#include "scann/utils/intrinsics/simd.h"
#include "scann/utils/internal/avx2_funcs.h"
#include "scann/utils/internal/avx_funcs.h"

#ifndef ABSL_INTERNAL_UNALIGNED_LOAD32
static inline uint32_t scann_ref_unaligned_load32(const void* p) { uint32_t v; memcpy(&v, p, 4); return v; }
#define ABSL_INTERNAL_UNALIGNED_LOAD32(p) scann_ref_unaligned_load32(p)
#endif
#ifndef ABSL_ANNOTATE_MEMORY_IS_INITIALIZED
#define ABSL_ANNOTATE_MEMORY_IS_INITIALIZED(p, n) do { } while (0)
#endif

namespace research_scann {

#include "bf16.inc"  // _ref/gen: utils/bfloat16_helpers.h:30-48 (Bfloat16Quantize / Bfloat16Decompress)

template <typename T>
class DatapointPtr {
 public:
  DatapointPtr(const void*, const T* values, size_t nonzero_entries, size_t dimensionality)
      : values_(values), n_(nonzero_entries), d_(dimensionality) {}
  const T* values() const { return values_; }
  size_t nonzero_entries() const { return n_; }
  size_t dimensionality() const { return d_; }
  bool IsDense() const { return true; }
 private:
  const T* values_;
  size_t n_, d_;
};

template <typename T>
struct RowMajorView {
  const T* base;
  size_t dims;
  size_t dimensionality() const { return dims; }
  const T* GetPtr(size_t i) const { return base + i * dims; }
};

namespace one_to_many_low_level {
template <typename ValueT>
inline size_t GetDatapointIndex(MutableSpan<ValueT>, size_t index) { return index; }
template <typename CallbackT> struct NeedsBottomBitsSideData : std::false_type {};
template <typename CallbackT, typename ResultT, typename DataT>
SCANN_INLINE void InvokeCallback(const CallbackT& callback, size_t result_idx, ResultT val, size_t, const DataT*) {
  callback.invoke(result_idx, val);
}
struct StoreFloat {
  float* out;
  template <typename ValueT> SCANN_INLINE void invoke(size_t index, ValueT val) const { out[index] = (float)val; }
  SCANN_INLINE void prefetch(size_t) const {}
};
}  // namespace one_to_many_low_level

namespace dp_internal {
namespace avx2_impl {
#define SCANN_SIMD_ATTRIBUTE SCANN_AVX2
#include "scann/distance_measures/one_to_one/dot_product_impl.inc"
#undef SCANN_SIMD_ATTRIBUTE
}  // namespace avx2_impl
// distance_measures/one_to_one/dot_product_avx2.cc:34-41
SCANN_AVX2_OUTLINE double DenseDotProductAvx2(const DatapointPtr<int8_t>& a, const DatapointPtr<float>& b) {
  return avx2_impl::DenseDotProductInt8FloatAvxImpl<AvxFunctionsAvx2Fma>(a.values(), b.values(), a.nonzero_entries());
}
// named by the non-AVX2 branches of StaticallyInvokeOneToOneDenseDotProduct; never called in this build
inline double DenseDotProductAvx1(const DatapointPtr<int8_t>&, const DatapointPtr<float>&) { abort(); }
inline double DenseDotProductSse4(const DatapointPtr<int8_t>&, const DatapointPtr<float>&) { abort(); }
}  // namespace dp_internal

namespace avx2 {
#define SCANN_SIMD_ATTRIBUTE SCANN_AVX2
#include "asym_impl.inc"  // _ref/gen: the whole file, five vector-type casts added (see oracle/Makefile)
#undef SCANN_SIMD_ATTRIBUTE
}  // namespace avx2

}  // namespace research_scann

using namespace research_scann;

extern "C" {

// DenseDotProductDistanceOneToManyInt8Float(query, dataset, result) (one_to_many_asymmetric.cc:44-49): result[i] =
// -<query, float(rows[i])> for i in [0, n), the last n mod 3 rows through the one-to-one kernel.
int ref_one_to_many_int8_float(const float* query, const int8_t* rows, uint64_t n, uint64_t dims, float* out) {
  RowMajorView<int8_t> view{rows, (size_t)dims};
  avx2::OneToManyInt8FloatImpl<false, false>(query, view, (const float*)nullptr, (const uint32_t*)nullptr,
                                             MutableSpan<float>(out, (size_t)n), one_to_many_low_level::StoreFloat{out});
  return 0;
}

// the same kernel with an index list (kHasIndices): out[i] = -<query, float(rows[indices[i]])> -- the call of the int8
// reordering helper (DenseDotProductDistanceOneToManyInt8Float(query, dataset, indices, result))
int ref_one_to_many_int8_float_indexed(const float* query, const int8_t* rows, uint64_t dims, const uint32_t* indices,
                                       uint64_t n, float* out) {
  RowMajorView<int8_t> view{rows, (size_t)dims};
  avx2::OneToManyInt8FloatImpl<true, false>(query, view, (const float*)nullptr, indices,
                                            MutableSpan<float>(out, (size_t)n), one_to_many_low_level::StoreFloat{out});
  return 0;
}

// DenseDotProductDistanceOneToManyBf16Float / OneToManyBf16FloatSquaredL2 (one_to_many_asymmetric.cc): bf16 rows
int ref_one_to_many_bf16_float(const float* query, const int16_t* rows, uint64_t n, uint64_t dims, int squared_l2,
                               float* out) {
  RowMajorView<int16_t> view{rows, (size_t)dims};
  if (squared_l2)
    avx2::OneToManyBf16FloatImpl<false, true>(query, view, (const uint32_t*)nullptr, MutableSpan<float>(out, (size_t)n),
                                              one_to_many_low_level::StoreFloat{out});
  else
    avx2::OneToManyBf16FloatImpl<false, false>(query, view, (const uint32_t*)nullptr, MutableSpan<float>(out, (size_t)n),
                                               one_to_many_low_level::StoreFloat{out});
  return 0;
}

}  // extern "C"
