// shim (oracle/_ref build only): restricts are out of scope; LUT16ArgsTopNBase only needs the view type to exist.
#pragma once
#include "scann/utils/common.h"
namespace research_scann {
class RestrictAllowlistConstView {
 public:
  const size_t* data() const { return nullptr; }
  bool empty() const { return true; }
  size_t size() const { return 0; }
  bool IsWhitelisted(size_t) const { return true; }
};
}  // namespace research_scann
