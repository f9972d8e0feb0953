// melonix_b200/host/range.hpp -- key type of the Spec cache.
// Keeps the three names the untouched front-end sees through app.hpp (reference range.hpp:4-18):
// MaxRanges, pair_hash, Range.
#pragma once
#include <cstddef>
#include <functional>
#include <utility>

static const auto MaxRanges = 4000;  // LRU capacity of Spec and SpecCache (reference range.hpp:4)

using Range = std::pair<int, int>;  // (start, end) sample indices of one spectrogram column

struct pair_hash
{
  template <class A, class B>
  std::size_t operator()(const std::pair<A, B> &p) const
  {
    // boost-style hash_combine of the two members
    std::size_t seed = std::hash<A>{}(p.first);
    seed ^= std::hash<B>{}(p.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};
