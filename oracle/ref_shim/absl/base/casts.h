#pragma once
#include <cstring>
namespace absl {
template <typename To, typename From>
inline To bit_cast(const From& f) { static_assert(sizeof(To) == sizeof(From)); To t; std::memcpy(&t, &f, sizeof t); return t; }
}  // namespace absl
