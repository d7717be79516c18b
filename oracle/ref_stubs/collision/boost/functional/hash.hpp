// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sources to compile where they lie (see oracle/Makefile, target ref).  boost::hash_combine / hash_range
// with Boost's published mixing step; the value only decides bucket order in the lattice's hash table.
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <typename T> inline void hash_combine(std::size_t& seed, const T& v)
{
    seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
template <typename It> inline std::size_t hash_range(It first, It last)
{
    std::size_t seed = 0;
    for (; first != last; ++first) hash_combine(seed, *first);
    return seed;
}
template <typename It> inline void hash_range(std::size_t& seed, It first, It last)
{
    for (; first != last; ++first) hash_combine(seed, *first);
}
} // namespace boost
