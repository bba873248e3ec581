// shim (oracle/_ref build only): the subset of absl::Span the compiled reference files use
#pragma once
#include <array>
#include <cstddef>
#include <type_traits>
#include <vector>
namespace absl {
template <typename T>
class Span {
 public:
  using value_type = std::remove_cv_t<T>;
  using iterator = T*;
  constexpr Span() : p_(nullptr), n_(0) {}
  constexpr Span(T* p, size_t n) : p_(p), n_(n) {}
  template <typename V, typename = std::enable_if_t<std::is_const_v<T> && std::is_same_v<std::remove_cv_t<T>, V>>>
  Span(const std::vector<V>& v) : p_(v.data()), n_(v.size()) {}
  template <typename V, typename = std::enable_if_t<!std::is_const_v<T> && std::is_same_v<T, V>>, int = 0>
  Span(std::vector<V>& v) : p_(v.data()), n_(v.size()) {}
  template <size_t N> Span(const std::array<value_type, N>& a) : p_(a.data()), n_(N) {}
  template <size_t N> Span(T (&a)[N]) : p_(a), n_(N) {}
  template <typename U, typename = std::enable_if_t<std::is_const_v<T> && std::is_same_v<U, value_type>>>
  Span(Span<U> o) : p_(o.data()), n_(o.size()) {}
  constexpr T* data() const { return p_; }
  constexpr size_t size() const { return n_; }
  constexpr bool empty() const { return n_ == 0; }
  constexpr T& operator[](size_t i) const { return p_[i]; }
  constexpr T* begin() const { return p_; }
  constexpr T* end() const { return p_ + n_; }
  constexpr T& front() const { return p_[0]; }
  constexpr T& back() const { return p_[n_ - 1]; }
  constexpr Span subspan(size_t pos, size_t len = static_cast<size_t>(-1)) const {
    return Span(p_ + pos, len == static_cast<size_t>(-1) ? n_ - pos : (len < n_ - pos ? len : n_ - pos));
  }
 private:
  T* p_;
  size_t n_;
};
template <typename T> Span<const T> MakeConstSpan(const T* p, size_t n) { return Span<const T>(p, n); }
template <typename T> Span<const T> MakeConstSpan(const std::vector<T>& v) { return Span<const T>(v.data(), v.size()); }
template <typename T> Span<const T> MakeConstSpan(Span<T> s) { return Span<const T>(s.data(), s.size()); }
template <typename T> Span<T> MakeSpan(T* p, size_t n) { return Span<T>(p, n); }
template <typename T> Span<T> MakeSpan(std::vector<T>& v) { return Span<T>(v.data(), v.size()); }
}  // namespace absl
