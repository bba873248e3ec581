// shim (oracle/_ref build only) for utils/intrinsics/highway.h: the class template name only (the x86 builds of the
// reference never instantiate it; it appears in `if constexpr (IsSame<T, Highway<float>>())` tests).
#pragma once
#include "hwy/highway.h"
#include "scann/utils/common.h"
namespace research_scann {
template <typename T, size_t kNumRegisters = 1, size_t... kTensorNumRegisters> class Highway;
}  // namespace research_scann
