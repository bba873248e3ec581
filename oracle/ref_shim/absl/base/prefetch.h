// shim (oracle/_ref build only): absl prefetch hints -> compiler builtins
#pragma once
namespace absl {
inline void PrefetchToLocalCache(const void* p) { __builtin_prefetch(p, 0, 3); }
inline void PrefetchToLocalCacheNta(const void* p) { __builtin_prefetch(p, 0, 0); }
}  // namespace absl
