// Test-infrastructure shim (oracle/_ref build only): the subset of Boost.ContainerHash that
// /root/reference/lmc uses -- hash_combine + boost::hash<T> dispatching to ADL hash_value().
// The mixing constants differ from Boost's; consumers only need a valid hash.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <utility>
#include <type_traits>

namespace boost {
template <class T> struct hash;

namespace shim_detail {
inline std::size_t mix(std::size_t x) {
  x ^= x >> 32; x *= 0xe9846af9b1a615dULL; x ^= x >> 32; x *= 0xe9846af9b1a615dULL; x ^= x >> 28;
  return x;
}
template <class T, class = void> struct is_std_hashable : std::false_type {};
template <class T>
struct is_std_hashable<T, std::enable_if_t<std::is_arithmetic_v<T> || std::is_enum_v<T>>> : std::true_type {};
}  // namespace shim_detail

template <class T>
inline std::enable_if_t<shim_detail::is_std_hashable<T>::value, std::size_t> hash_value(const T &v) {
  return static_cast<std::size_t>(v);
}
template <class A, class B> inline std::size_t hash_value(const std::pair<A, B> &p);

template <class T> inline void hash_combine(std::size_t &seed, const T &v) {
  boost::hash<T> hasher;
  seed = shim_detail::mix(seed + 0x9e3779b97f4a7c15ULL + hasher(v));
}
template <class A, class B> inline std::size_t hash_value(const std::pair<A, B> &p) {
  std::size_t seed = 0;
  hash_combine(seed, p.first);
  hash_combine(seed, p.second);
  return seed;
}
template <class T> struct hash {
  std::size_t operator()(const T &v) const { return hash_value(v); }  // ADL finds friends
};
}  // namespace boost
