// oracle/_ref glue, part 2: line ranges of the reference's asymmetric_hashing_impl.cc and bfloat16_helpers.h,
// extracted at BUILD time into oracle/_ref/gen/ (git-ignored) by oracle/Makefile and compiled here:
//   _ref/gen/ah_lut_convert.inc  = hashes/internal/asymmetric_hashing_impl.cc:571-645   ComputeMultiplierByQuantile,
//                                  ConvertLookupToFixedPointImpl, ConvertLookupToFixedPoint (float LUT -> u8 LUT)
//   _ref/gen/ah_pack.inc         = hashes/internal/asymmetric_hashing_impl.cc:690-737   CreatePackedDataset
//   _ref/gen/bf16.inc            = utils/bfloat16_helpers.h:30-48                       Bfloat16Quantize / Decompress
// Test infrastructure; no reference code is stored in the repository.  What this file supplies are the declarations
// those ranges lean on and that live in headers this image cannot compile (protobuf, highway, DenseDataset):
// the config message accessors, FixedPointBias (asymmetric_hashing_impl.h:170-173, one shift), MaxAbsValue
// (utils/util_functions.cc:36-75 computes max |x| with highway; max is exact, so any order gives the same float),
// a TopNAmortizedConstant stand-in for the quantile < 1 branch (never taken: the builder always writes quantile 1.0),
// and a row-major view with DenseDataset's operator[] / nonzero_entries() / values().
#include <algorithm>
#include <cmath>
#include <cstring>

#include "absl/base/casts.h"
#include "scann/utils/common.h"
#include "scann/utils/types.h"

namespace research_scann {

struct AsymmetricHasherConfig {
  struct FixedPointLUTConversionOptions {
    enum Method { TRUNCATE = 0, ROUND = 1 };
    float quantile = 1.0f;
    int method = ROUND;
    float multiplier_quantile() const { return quantile; }
    int float_to_int_conversion_method() const { return method; }
  };
};

inline float MaxAbsValue(ConstSpan<float> arr) {
  float m = 0.f;
  for (float v : arr) m = std::max(m, std::abs(v));
  return m;
}

template <typename T>
class TopNAmortizedConstant {
 public:
  explicit TopNAmortizedConstant(size_t k) : k_(k) {}
  void push(T v) { v_.push_back(v); }
  T exact_bottom() {
    std::nth_element(v_.begin(), v_.begin() + (k_ - 1), v_.end(), std::greater<T>());
    return v_[k_ - 1];
  }
 private:
  size_t k_;
  std::vector<T> v_;
};

// row-major [n][b] code matrix with the accessors CreatePackedDataset uses
struct CodeRow {
  const uint8_t* p;
  uint64_t b;
  uint64_t nonzero_entries() const { return b; }
  const uint8_t* values() const { return p; }
};
template <typename T>
struct DenseDataset;
template <>
struct DenseDataset<uint8_t> {
  const uint8_t* base;
  uint32_t n;
  uint64_t b;
  bool empty() const { return n == 0; }
  uint32_t size() const { return n; }
  CodeRow operator[](size_t i) const { return CodeRow{base + i * b, b}; }
};

namespace asymmetric_hashing_internal {
template <typename Uint>
inline constexpr Uint FixedPointBias() { return static_cast<Uint>(1) << ((sizeof(Uint) * 8) - 1); }

template <typename T>
std::vector<T> ConvertLookupToFixedPoint(ConstSpan<float> raw_lookup,
                                         const AsymmetricHasherConfig::FixedPointLUTConversionOptions& conversion_options,
                                         float* multiplier);

#include "ah_lut_convert.inc"

template std::vector<uint8_t> ConvertLookupToFixedPoint<uint8_t>(
    ConstSpan<float>, const AsymmetricHasherConfig::FixedPointLUTConversionOptions&, float*);

#include "ah_pack.inc"
}  // namespace asymmetric_hashing_internal

#include "bf16.inc"

}  // namespace research_scann

using namespace research_scann;

extern "C" {

// ConvertLookupToFixedPoint<uint8_t> with the builder's options (quantile 1.0, ROUND; truncate != 0 selects TRUNCATE)
int ref_lut_to_fixed_point(const float* raw, uint64_t n, int truncate, uint8_t* out, float* multiplier) {
  AsymmetricHasherConfig::FixedPointLUTConversionOptions opt;
  opt.method = truncate ? AsymmetricHasherConfig::FixedPointLUTConversionOptions::TRUNCATE
                        : AsymmetricHasherConfig::FixedPointLUTConversionOptions::ROUND;
  std::vector<uint8_t> r =
      asymmetric_hashing_internal::ConvertLookupToFixedPoint<uint8_t>(ConstSpan<float>(raw, n), opt, multiplier);
  std::memcpy(out, r.data(), n);
  return 0;
}

// CreatePackedDataset: codes [n][b] (one 4-bit code per byte) -> packed [ceil(n / 32)][b][16]; returns the size
uint64_t ref_pack_dataset(const uint8_t* codes, uint32_t n, uint64_t b, uint8_t* out, uint64_t out_cap) {
  DenseDataset<uint8_t> ds{codes, n, b};
  std::vector<uint8_t> p = asymmetric_hashing_internal::CreatePackedDataset(ds);
  if (p.size() <= out_cap) std::memcpy(out, p.data(), p.size());
  return p.size();
}

void ref_bf16_quantize(const float* x, uint64_t n, int16_t* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = Bfloat16Quantize(x[i]);
}
void ref_bf16_decompress(const int16_t* x, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = Bfloat16Decompress(x[i]);
}

}  // extern "C"
