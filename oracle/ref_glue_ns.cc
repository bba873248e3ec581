// oracle/_ref glue, part 6: the reference's own noise-shaped (anisotropic) AH encoder -- the coordinate descent of
// AhImpl<T>::IndexDatapointNoiseShaped (hashes/internal/asymmetric_hashing_impl.cc:434-503), reached from
// Indexer::HashWithNoiseShaping (hashes/asymmetric_hashing2/indexing.cc:185-246) in the index build.
//
// Test infrastructure (see oracle/scann_oracle.h): only tests/ may load the resulting library.  This file contains no
// reference code; it INCLUDES line ranges of asymmetric_hashing_impl.cc, extracted at BUILD time into oracle/_ref/gen/
// (git-ignored) by oracle/Makefile:
//   _ref/gen/ns_helpers_a.inc = :263-298  Square, ComputeParallelCostMultiplier, SubspaceResidualStats,
//                                         ComputeResidualStatsForCluster
//   _ref/gen/ns_helpers_b.inc = :349-416  InitializeToMinResidualNorm, ComputeParallelResidualComponent,
//                                         CoordinateDescentResult, OptimizeSingleSubspace
//   _ref/gen/ns_body.inc      = :450-501  the body of IndexDatapointNoiseShaped after the residual statistics: cost
//                                         multiplier, initialisation, block order (ZipSortBranchOptimized on the
//                                         initial residual norms, descending), <= 10 rounds of coordinate descent
// and the reference's own utils/zip_sort.h (the block order, including how it orders EQUAL norms).  What is written
// here: the chunking of the datapoint into blocks (ComputeResidualStats :300-347 does it through ChunkingProjection,
// which needs the protobuf config) with the chunked norm of :317-325, and SquaredL2Norm through the reference's
// DenseSingleAccumulate (ref_squared_l2_norm of ref_glue_sym.cc).
// Compiled with -ffp-contract=off like the rest of the library: two single-expression multiply-adds of this code would
// be fused by the reference's documented clang build, g++'s "fast" mode would ALSO fuse the
// Square() accumulations, which clang does not; "off" is the mode the oracle restates.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "scann/utils/common.h"
#include "scann/utils/types.h"
#include "scann/utils/zip_sort.h"

extern "C" double ref_squared_l2_norm(const float* v, uint64_t n);

namespace research_scann {

template <typename T> using FloatingTypeFor = float;
template <typename C> SCANN_INLINE SeqRange IndicesOf(const C& c) { return Seq(c.size()); }

struct OriginalPtr {
  const float* p;
  size_t d;
  size_t dimensionality() const { return d; }
};
inline double SquaredL2Norm(const OriginalPtr& o) { return ref_squared_l2_norm(o.p, o.d); }

namespace {
#include "ns_helpers_a.inc"
#include "ns_helpers_b.inc"

// one datapoint: codes[B] from its (maybe residual) vector, the original vector and the codebook
void NoiseShapedOne(const float* maybe_residual, const float* original, size_t dims, const float* codebook, size_t B,
                    size_t stride, const int32_t* block_dims, double threshold, uint8_t* out) {
  const double eta = NAN;
  OriginalPtr original_dptr{original, dims};
  // ComputeResidualStats (:300-347): chunked norm of the original over the blocks in order, then the statistics of every
  // (block, centre)
  double chunked_norm = 0.0;
  for (size_t i = 0; i < dims; ++i) chunked_norm += Square<double>(original[i]);
  chunked_norm = std::sqrt(chunked_norm);
  double inverse_chunked_norm = 1.0 / chunked_norm;
  std::vector<std::vector<SubspaceResidualStats>> residual_stats(B);
  size_t off = 0;
  for (size_t b = 0; b < B; ++b) {
    const size_t bd = block_dims ? (size_t)block_dims[b] : stride;
    residual_stats[b].resize(16);
    for (size_t c = 0; c < 16; ++c) {
      residual_stats[b][c] = ComputeResidualStatsForCluster<float>(
          ConstSpan<float>(maybe_residual + off, bd), ConstSpan<float>(original + off, bd), inverse_chunked_norm,
          ConstSpan<float>(codebook + (b * 16 + c) * stride, bd));
    }
    off += bd;
  }
  std::vector<uint8_t> result_storage(B);
  MutableSpan<uint8_t> result(result_storage.data(), B);
#include "ns_body.inc"
  (void)final_residual_norm;
  memcpy(out, result_storage.data(), B);
}
}  // namespace
}  // namespace research_scann

extern "C" {

// codes [n][B] of AhImpl<float>::IndexDatapointNoiseShaped for rows x - centers[token] (centers NULL: x itself), the
// "original" being x (Indexer::HashWithNoiseShaping(maybe_residual, original, ...)); codebook [B][16][stride]
int ref_encode_noise_shaped(const float* x, uint64_t n, uint64_t dims, const float* centers, const int32_t* token,
                            const float* codebook, uint64_t B, uint64_t stride, const int32_t* block_dims,
                            double threshold, uint8_t* out) {
  std::vector<float> res(dims);
  for (uint64_t i = 0; i < n; ++i) {
    const float* xi = x + i * dims;
    for (uint64_t k = 0; k < dims; ++k) res[k] = centers ? xi[k] - centers[(uint64_t)token[i] * dims + k] : xi[k];
    research_scann::NoiseShapedOne(res.data(), xi, dims, codebook, B, stride, block_dims, threshold, out + i * B);
  }
  return 0;
}

}  // extern "C"
