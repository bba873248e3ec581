#pragma once
#include <cstdint>
namespace absl {
inline int countl_zero(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline int countl_zero(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
inline int countr_zero(uint32_t x) { return x ? __builtin_ctz(x) : 32; }
inline int countr_zero(uint64_t x) { return x ? __builtin_ctzll(x) : 64; }
inline int bit_width(uint32_t x) { return 32 - countl_zero(x); }
inline int bit_width(uint64_t x) { return 64 - countl_zero(x); }
inline int popcount(uint32_t x) { return __builtin_popcount(x); }
inline int popcount(uint64_t x) { return __builtin_popcountll(x); }
}  // namespace absl
