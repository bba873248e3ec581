// oracle/_ref glue, part 1: the reference's own AVX2 LUT16 kernel, compiled from where it lies.
//
// Test infrastructure (see oracle/scann_oracle.h): only tests/ may load the resulting library.  This file contains no
// reference code; it INCLUDES /root/reference/scann/hashes/internal/lut16_avx2.inc (the whole file: BottomLoop :55-124,
// GetInt16Distances :274-292, GetTopFloatDistances :404-527) and, through it, the reference's SIMD wrappers
// utils/intrinsics/{sse4,avx1,avx2}.h, utils/bits.h and hashes/internal/lut16_args.h, against the shim headers of
// oracle/ref_shim/ (spans, logging, flags, a record-everything stand-in for FastTopNeighbors).  The extern "C" entry
// points below hand plain arrays to LUT16Avx2<>.
#include "scann/hashes/internal/lut16_avx2.inc"

namespace research_scann {
namespace asymmetric_hashing_internal {
template class LUT16Avx2<1, PrefetchStrategy::kOff>;
template class LUT16Avx2<2, PrefetchStrategy::kOff>;
template class LUT16Avx2<3, PrefetchStrategy::kOff>;
template class LUT16Avx2<3, PrefetchStrategy::kSeq>;
}  // namespace asymmetric_hashing_internal
}  // namespace research_scann

using namespace research_scann;
using namespace research_scann::asymmetric_hashing_internal;

extern "C" {

// LUT16Avx2<nq>::GetInt16Distances: packed [n32][num_blocks][16] bytes, luts nq x [num_blocks][16] u8,
// out nq x [32 * n32] int16.  nq in 1..3.
int ref_lut16_int16(const uint8_t* packed, uint64_t n32, uint64_t num_blocks, const uint8_t* const* luts, int nq,
                    int16_t* const* out) {
  LUT16Args<int16_t> a;
  a.packed_dataset = packed;
  a.num_32dp_simd_iters = n32;
  a.num_blocks = num_blocks;
  a.lookups = ConstSpan<const uint8_t*>(luts, (size_t)nq);
  a.distances = ConstSpan<int16_t*>(out, (size_t)nq);
  a.prefetch_strategy = PrefetchStrategy::kOff;
  switch (nq) {
    case 1: LUT16Avx2<1, PrefetchStrategy::kOff>::GetInt16Distances(a); return 0;
    case 2: LUT16Avx2<2, PrefetchStrategy::kOff>::GetInt16Distances(a); return 0;
    case 3: LUT16Avx2<3, PrefetchStrategy::kOff>::GetInt16Distances(a); return 0;
  }
  return 1;
}

// LUT16Avx2<nq>::GetTopFloatDistances with a top-N that keeps every push and an epsilon of `epsilon` (+inf: the
// int16 pre-filter passes everything below 32767): float scores (acc * (1 / mult) + bias) of the pushed datapoints.
// out_idx / out_dist: nq x [cap]; out_count[nq].  Returns 0, or 2 if a list did not fit.
int ref_lut16_top_float(const uint8_t* packed, uint64_t n32, uint64_t num_blocks, uint32_t num_datapoints,
                        const uint8_t* const* luts, int nq, const float* biases, const float* mults, float epsilon,
                        uint32_t first_dp_index, uint32_t* const* out_idx, float* const* out_dist, uint32_t cap,
                        uint32_t* out_count, int seq_prefetch) {
  std::vector<FastTopNeighbors<float>> tops;
  tops.reserve(nq);
  std::vector<FastTopNeighbors<float>*> ptrs;
  for (int j = 0; j < nq; ++j) { tops.emplace_back(epsilon); ptrs.push_back(&tops[j]); }
  LUT16ArgsTopN<float> a;
  a.packed_dataset = packed;
  a.num_32dp_simd_iters = n32;
  a.num_blocks = num_blocks;
  a.lookups = ConstSpan<const uint8_t*>(luts, (size_t)nq);
  a.first_dp_index = first_dp_index;
  a.num_datapoints = num_datapoints;
  a.fast_topns = ConstSpan<FastTopNeighbors<float>*>(ptrs.data(), ptrs.size());
  a.biases = ConstSpan<float>(biases, (size_t)nq);
  a.fixed_point_multipliers = ConstSpan<float>(mults, (size_t)nq);
  a.prefetch_strategy = seq_prefetch ? PrefetchStrategy::kSeq : PrefetchStrategy::kOff;
  switch (nq) {
    case 1: LUT16Avx2<1, PrefetchStrategy::kOff>::GetTopFloatDistances(std::move(a)); break;
    case 2: LUT16Avx2<2, PrefetchStrategy::kOff>::GetTopFloatDistances(std::move(a)); break;
    case 3:
      if (seq_prefetch) LUT16Avx2<3, PrefetchStrategy::kSeq>::GetTopFloatDistances(std::move(a));
      else LUT16Avx2<3, PrefetchStrategy::kOff>::GetTopFloatDistances(std::move(a));
      break;
    default: return 1;
  }
  int rc = 0;
  for (int j = 0; j < nq; ++j) {
    const auto& r = tops[j].results;
    out_count[j] = (uint32_t)r.size();
    if (r.size() > cap) { rc = 2; continue; }
    for (size_t i = 0; i < r.size(); ++i) { out_idx[j][i] = r[i].first; out_dist[j][i] = r[i].second; }
  }
  return rc;
}

}  // extern "C"
