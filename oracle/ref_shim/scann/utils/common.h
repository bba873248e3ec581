// shim (oracle/_ref build only) for /root/reference/scann/utils/common.h: the few infrastructure names the compiled
// reference files (lut16_avx2.inc, utils/intrinsics/*.h, utils/bits.h, the extracted line ranges) use -- spans,
// Seq(), DivRoundUp(), the type predicates, the inlining macros.  No arithmetic of the hot path lives here; the real
// header needs all of abseil (containers, status, strings, flags), which this image does not have.
#pragma once
#include <stddef.h>
#include <sys/types.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
#include <vector>

#include "absl/base/attributes.h"
#include "absl/base/optimization.h"
#include "absl/base/prefetch.h"
#include "absl/log/check.h"
#include "absl/types/span.h"

namespace research_scann {
using ::std::array;
using ::std::numeric_limits;
using ::std::pair;
using ::std::string;
using ::std::string_view;
using ::std::vector;
using ::std::make_signed_t;
using ::std::make_unsigned_t;
using ::std::conditional_t;
using ::std::decay_t;
using ::std::enable_if_t;
using ::std::declval;

template <typename T> using ConstSpan = absl::Span<const T>;
template <typename T> using MutableSpan = absl::Span<T>;
using ::absl::MakeConstSpan;
template <typename... Args> auto MakeMutableSpan(Args&&... args) { return absl::MakeSpan(std::forward<Args>(args)...); }

#define SCANN_INLINE inline ABSL_ATTRIBUTE_ALWAYS_INLINE
#define SCANN_INLINE_LAMBDA ABSL_ATTRIBUTE_ALWAYS_INLINE
#define SCANN_OUTLINE ABSL_ATTRIBUTE_NOINLINE

#define SCANN_DECLARE_COPYABLE_CLASS(ClassName) \
  ClassName(ClassName&&) = default;             \
  ClassName& operator=(ClassName&&) = default;  \
  ClassName(const ClassName&) = default;        \
  ClassName& operator=(const ClassName&) = default
#define SCANN_DECLARE_IMMOBILE_CLASS(ClassName) \
  ClassName(ClassName&&) = delete;              \
  ClassName& operator=(ClassName&&) = delete;   \
  ClassName(const ClassName&) = delete;         \
  ClassName& operator=(const ClassName&) = delete
#define SCANN_DECLARE_MOVE_ONLY_CLASS(ClassName) \
  ClassName(ClassName&&) = default;              \
  ClassName& operator=(ClassName&&) = default;   \
  ClassName(const ClassName&) = delete;          \
  ClassName& operator=(const ClassName&) = delete

template <typename T, typename U> inline constexpr bool IsSame() { return std::is_same_v<std::decay_t<T>, std::decay_t<U>>; }
template <typename T, typename... UU> inline constexpr bool IsSameAny() { return (IsSame<T, UU>() || ...); }
template <typename T> inline constexpr bool IsUint8() { return IsSame<T, uint8_t>(); }
template <typename T> inline constexpr bool IsFloat() { return IsSame<T, float>(); }
template <typename T> inline constexpr bool IsDouble() { return IsSame<T, double>(); }
template <typename T> inline constexpr bool IsFloatingType() { return std::is_floating_point_v<std::decay_t<T>>; }
template <typename T> inline constexpr bool IsIntegerType() { return std::is_integral_v<std::decay_t<T>>; }
template <typename T> inline constexpr bool IsSignedType() { return std::is_signed_v<std::decay_t<T>>; }

template <typename Int, typename DenomInt>
constexpr Int DivRoundUp(Int num, DenomInt denom) {
  return (num + static_cast<Int>(denom) - static_cast<Int>(1)) / static_cast<Int>(denom);
}
template <typename Int, typename DenomInt> constexpr Int NextMultipleOf(Int num, DenomInt denom) { return DivRoundUp(num, denom) * denom; }
template <typename Int, typename DenomInt> constexpr bool IsDivisibleBy(Int num, DenomInt denom) { return num % denom == 0; }

// for (size_t j : Seq(n)): iterates 0 .. n-1
class SeqRange {
 public:
  class It {
   public:
    SCANN_INLINE explicit It(size_t i) : i_(i) {}
    SCANN_INLINE size_t operator*() const { return i_; }
    SCANN_INLINE It& operator++() { ++i_; return *this; }
    SCANN_INLINE bool operator!=(It e) const { return i_ < e.i_; }
   private:
    size_t i_;
  };
  SCANN_INLINE SeqRange(size_t b, size_t e) : b_(b), e_(e) {}
  SCANN_INLINE It begin() const { return It(b_); }
  SCANN_INLINE It end() const { return It(e_); }
 private:
  size_t b_, e_;
};
SCANN_INLINE SeqRange Seq(size_t end) { return SeqRange(0, end); }
SCANN_INLINE SeqRange Seq(size_t begin, size_t end) { return SeqRange(begin, end); }
}  // namespace research_scann
