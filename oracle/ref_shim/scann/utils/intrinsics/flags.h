// shim (oracle/_ref build only) for utils/intrinsics/flags.h: the platform enum the SIMD wrappers name; no absl flags.
#pragma once
#include "scann/utils/common.h"
#include "scann/utils/types.h"
namespace research_scann {
inline bool RuntimeSupportsSse4() { return true; }
inline bool RuntimeSupportsAvx1() { return true; }
inline bool RuntimeSupportsAvx2() { return true; }
inline bool RuntimeSupportsAvx512() { return false; }
inline bool RuntimeSupportsAvx512Vnni() { return false; }
inline bool RuntimeSupportsAmx() { return false; }
enum PlatformGeneration {
  kFallbackForNonX86 = 99, kHighway = 98, kBaselineSse4 = 0, kSandyBridgeAvx1 = 1, kHaswellAvx2 = 2,
  kSkylakeAvx512 = 3, kCascadelakeAvx512Vnni = 4, kSapphireRapidsAmx = 5,
};
}  // namespace research_scann
