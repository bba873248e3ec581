// shim (oracle/_ref build only): CHECK / DCHECK / LOG as no-op streams (NDEBUG semantics) or abort
#pragma once
#include "absl/log/check.h"
#if 0
#include <cstdlib>
#include <iostream>
namespace shim_log {
struct Null { template <typename T> Null& operator<<(const T&) { return *this; } };
struct Fatal { template <typename T> Fatal& operator<<(const T& v) { std::cerr << v; return *this; } [[noreturn]] ~Fatal() { std::cerr << std::endl; std::abort(); } };
}  // namespace shim_log
#define SHIM_NULL_STREAM while (false) ::shim_log::Null()
#define CHECK(c) if (!(c)) ::shim_log::Fatal() << "CHECK failed: " #c " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define DCHECK(c) SHIM_NULL_STREAM
#define DCHECK_EQ(a, b) SHIM_NULL_STREAM
#define DCHECK_NE(a, b) SHIM_NULL_STREAM
#define DCHECK_LT(a, b) SHIM_NULL_STREAM
#define DCHECK_LE(a, b) SHIM_NULL_STREAM
#define DCHECK_GT(a, b) SHIM_NULL_STREAM
#define DCHECK_GE(a, b) SHIM_NULL_STREAM
#define SHIM_LOG_FATAL ::shim_log::Fatal()
#define SHIM_LOG_INFO ::shim_log::Null()
#define SHIM_LOG_WARNING ::shim_log::Null()
#define SHIM_LOG_ERROR ::shim_log::Null()
#define LOG(sev) SHIM_LOG_##sev
#define DLOG(sev) SHIM_NULL_STREAM
#define VLOG(n) SHIM_NULL_STREAM
#endif
