// melonix_b200/host/picks.cpp -- see picks.hpp / include/melonix_host.h.
#include "picks.hpp"

#include "colour_ramp.hpp"

#include "../../include/melonix_host.h"

#include <algorithm>
#include <cmath>

namespace melonix
{
Picks::Picks(std::span<const float> wav, std::vector<MinMax> flat, std::vector<int64_t> levelOff)
  : wav(wav), flat(std::move(flat)), levelOff(std::move(levelOff))
{
}

auto Picks::levels(int64_t n) -> int
{
  int lvl = 0;
  while (n > (int64_t{1} << (lvl + 1)))
    ++lvl;
  return lvl;
}

auto Picks::layout(int64_t n) -> std::vector<int64_t>
{
  const int L = levels(n);
  std::vector<int64_t> off(L + 1);
  int64_t o = 0;
  for (int l = 0; l < L; ++l)
  {
    off[l] = o;
    o += n >> (l + 1);
  }
  off[L] = o;
  return off;
}

auto Picks::level(int l) const -> std::span<const MinMax>
{
  return {flat.data() + levelOff[l], static_cast<size_t>(levelOff[l + 1] - levelOff[l])};
}

auto Picks::getMinMaxFromRange(int start, int end) const -> MinMax
{
  const int n = static_cast<int>(wav.size());
  if (start >= end)
  {
    if (start >= 0 && start < n)
      return {wav[start], wav[start]};
    return {0.f, 0.f};
  }
  if (start < 0 || end < 0)
    return {0.f, 0.f};
  if (start >= n || end >= n)
    return {0.f, 0.f};
  if (end - start == 1)
    return {wav[start], wav[start]};
  // the block of the largest level that fits the range and contains `start` ...
  const auto lvl = static_cast<size_t>(std::log2(end - start));
  const int lvlStart = start / (1 << lvl);
  MinMax mm{0.f, 0.f};
  if (lvl - 1 < static_cast<size_t>(levelCount()) && lvlStart < static_cast<int>(level(static_cast<int>(lvl) - 1).size()))
    mm = level(static_cast<int>(lvl) - 1)[lvlStart];
  // ... the sample at `start` when the block begins exactly there ...
  const int leftEnd = lvlStart * (1 << lvl);
  if (leftEnd >= start)
  {
    const auto l = getMinMaxFromRange(start, leftEnd);
    mm.first = std::min(mm.first, l.first);
    mm.second = std::max(mm.second, l.second);
  }
  // ... and whatever is left on the right
  const int rightStart = (lvlStart + 1) * (1 << lvl);
  if (rightStart < end)
  {
    const auto r = getMinMaxFromRange(rightStart, end);
    mm.first = std::min(mm.first, r.first);
    mm.second = std::max(mm.second, r.second);
  }
  return mm;
}
} // namespace melonix

extern "C" {
int mlxh_picks_levels(int64_t n)
{
  return melonix::Picks::levels(n);
}
int64_t mlxh_picks_layout(int64_t n, int64_t *level_off)
{
  const auto off = melonix::Picks::layout(n);
  std::copy(off.begin(), off.end(), level_off);
  return off.back();
}
void mlxh_minmax_ranges(const float *wav, int64_t n, const float *pairs, const int32_t *start_end, int count,
                        float *out)
{
  auto off = melonix::Picks::layout(n);
  const auto *p = reinterpret_cast<const melonix::Picks::MinMax *>(pairs);
  const melonix::Picks picks({wav, static_cast<size_t>(n)}, {p, p + off.back()}, off);
  for (int i = 0; i < count; ++i)
  {
    const auto mm = picks.getMinMaxFromRange(start_end[2 * i], start_end[2 * i + 1]);
    out[2 * i] = mm.first;
    out[2 * i + 1] = mm.second;
  }
}
void mlxh_colour_ramp(const float *mag, int count, float k, uint8_t *rgb)
{
  for (int i = 0; i < count; ++i)
  {
    const auto t = melonix::rampTexel(mag[i], k);
    std::copy(t.begin(), t.end(), rgb + 3 * static_cast<size_t>(i));
  }
}
}
