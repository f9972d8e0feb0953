// melonix_b200/host/colour_ramp.hpp -- one texel of the colour ramp of SpecCache::populateTex
// (reference spec-cache.cpp:77-96) on the host.  New columns get their texels from the fused GPU
// epilogue (mlx_spec_batch_rgb); this is used only to recolour a column whose float spectrum is
// already cached when the brightness gain changes -- the reference recolours from the cached floats
// too.  Same arithmetic: float clamp, integer thresholds 85 / 170, the angle in double with the
// reference's 3.141592, truncating casts.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>

namespace melonix
{
inline auto rampTexel(float mag, float gain) -> std::array<unsigned char, 3>
{
  const float v = std::clamp(mag * gain, 0.f, 255.f);
  constexpr int third = 255 / 3;
  const auto trunc8 = [](auto x) { return static_cast<unsigned char>(x); };
  if (v < third)
    return {trunc8(v), 0, 0};
  if (v >= 2 * third)
  {
    const unsigned char side = trunc8((v - 2 * third) * 3);
    return {side, trunc8(v), side};
  }
  const double angle = (v - third) / third * 3.141592 / 2;
  return {trunc8(v * std::cos(angle)), trunc8(v * std::sin(angle)), 0};
}
} // namespace melonix
