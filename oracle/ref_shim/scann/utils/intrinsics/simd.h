// shim (oracle/_ref build only) for utils/intrinsics/simd.h: the reference's own SSE4 / AVX1 / AVX2 wrappers (the real
// files, from /root/reference), and the NAMES of the platforms this build does not compile (AVX-512, AMX, Highway:
// they appear in `if constexpr (IsSame<T, Avx512<float>>())` tests of the files compiled here).
#pragma once
#include "scann/utils/intrinsics/attributes.h"
#include "scann/utils/intrinsics/avx1.h"
#include "scann/utils/intrinsics/avx2.h"
#include "scann/utils/intrinsics/flags.h"
#include "scann/utils/intrinsics/highway.h"
#include "scann/utils/intrinsics/sse4.h"
namespace research_scann {
// complete (so that `Avx512<float>{...}` inside never-instantiated `if constexpr` branches parses), never used
template <typename T, size_t kNumRegisters = 1, size_t... kTensorNumRegisters>
class Avx512 {
 public:
  Avx512() {}
  template <typename U> Avx512(U) {}
};
}  // namespace research_scann
