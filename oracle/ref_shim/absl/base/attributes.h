#pragma once
#define ABSL_ATTRIBUTE_ALWAYS_INLINE __attribute__((always_inline))
#define ABSL_ATTRIBUTE_NOINLINE __attribute__((noinline))
#define ABSL_MUST_USE_RESULT
#define ABSL_ATTRIBUTE_NO_SANITIZE_UNDEFINED
