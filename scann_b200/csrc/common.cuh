// common.cuh -- shared device helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr uint64_t kKeyMax = ~0ull;

// float -> u32 that orders like the float under DistanceComparator
// (utils/util_functions.h:94-107); -0.0 is canonicalised to +0.0.
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(__fadd_rn(f, 0.0f));
#else
  f = f + 0.0f;
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
  return ((uint64_t)f2ord(score) << 32) | idx;
}

__device__ __forceinline__ int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// Bitonic sort (ascending) of n = 2^k u64 keys in shared memory by the whole block.
__device__ __forceinline__ void block_bitonic_sort(uint64_t* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int l = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int r = l | j;
        const uint64_t a = s[l], b = s[r];
        const bool up = (l & k) == 0;
        if ((a > b) == up) { s[l] = b; s[r] = a; }
      }
      __syncthreads();
    }
  }
}

// streaming (read-once) 128/64/32-bit global loads that do not allocate in L1
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

}  // namespace sb
