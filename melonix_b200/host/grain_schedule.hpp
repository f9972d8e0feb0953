// melonix_b200/host/grain_schedule.hpp -- host side of the grain path (see include/melonix_host.h).
#pragma once
#include <cstdint>
#include <span>
#include <vector>

namespace melonix
{
struct MarkerView // the fields of the reference's Marker (marker.hpp:4-19) the warp maps read
{
  int sample;
  double dTime;
  double pitchBend;
};

struct Grain
{
  int start;
  int len;
};

struct RenderSchedule // input of mlx_grain_render
{
  std::vector<int32_t> gStart, gLen;
  std::vector<float> rate, next;
  std::vector<int64_t> outOff; // rows + 1
  int tailZeros = 0;
};

constexpr int PreferredGrainSize = 1500; // reference app.cpp:19

auto segmentGrains(std::span<const float> wav) -> std::vector<Grain>;

class WarpMaps // sample <-> time and time -> pitch bend through the sorted markers
{
public:
  WarpMaps(std::span<const MarkerView> markers, int sampleRate, int64_t nSamples)
    : markers(markers), sampleRate(sampleRate), nSamples(nSamples)
  {
  }
  auto sample2Time(int sample) const -> double;
  auto time2Sample(double t) const -> int;
  auto duration() const -> double;
  auto time2PitchBend(double t) const -> float;

private:
  std::span<const MarkerView> markers;
  int sampleRate;
  int64_t nSamples;
};

auto buildExportSchedule(std::span<const float> wav, int sampleRate, std::span<const MarkerView> markers,
                         std::span<const Grain> grains) -> RenderSchedule;
} // namespace melonix
