// shim (oracle/_ref build only) for /root/reference/scann/utils/types.h: index types only (the real header needs
// abseil and the generated protobuf headers).
#pragma once
#include "scann/utils/common.h"
namespace research_scann {
using DatapointIndex = uint32_t;
enum : DatapointIndex { kInvalidDatapointIndex = std::numeric_limits<DatapointIndex>::max() };
using DimensionIndex = uint64_t;
using NNResultsVector = std::vector<std::pair<DatapointIndex, float>>;
static constexpr int kNumDatapointsPerBlock = 32;   // utils/types.h:474
static constexpr int kPackedDatasetBlockSize = 1 << 4;  // utils/types.h:476
}  // namespace research_scann
