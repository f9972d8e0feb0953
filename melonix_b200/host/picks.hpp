// melonix_b200/host/picks.hpp -- host side of the waveform level-of-detail cache.
//
// The pyramid itself is built on the GPU (mlx_picks_build, include/melonix_gpu.h, replacing
// App::calcPicks, reference app.cpp:347-378).  The front-end then asks one (start, end) range per
// screen column per UI frame (App::getMinMaxFromRange, app.cpp:380-426, from glDraw); a device round
// trip per query would cost more than the query, so this class answers single queries on the host
// from the downloaded pyramid, step for step as the reference does (mlx_minmax_ranges is the batched
// device form for whole screens).
#pragma once
#include <cstdint>
#include <span>
#include <utility>
#include <vector>

namespace melonix
{
class Picks
{
public:
  using MinMax = std::pair<float, float>;
  Picks() = default;
  // `wav`: the non-owning view the reference's query also reads (app.cpp:385, :396);
  // `flat` / `levelOff`: what mlx_picks_build / mlx_picks_layout returned (levelOff has levels+1 entries)
  Picks(std::span<const float> wav, std::vector<MinMax> flat, std::vector<int64_t> levelOff);
  static auto levels(int64_t n) -> int;                            // app.cpp:352, :365
  static auto layout(int64_t n) -> std::vector<int64_t>;           // first entry of each level, + total
  auto getMinMaxFromRange(int start, int end) const -> MinMax;     // app.cpp:380-426
  auto level(int l) const -> std::span<const MinMax>;              // the reference's picks[l]
  auto levelCount() const -> int { return static_cast<int>(levelOff.size()) - 1; }

private:
  std::span<const float> wav;
  std::vector<MinMax> flat;
  std::vector<int64_t> levelOff{0};
};
} // namespace melonix
