// melonix_b200/host/grain_schedule.cpp -- see grain_schedule.hpp / include/melonix_host.h.
#include "grain_schedule.hpp"

#include "../../include/melonix_host.h"

#include <algorithm>
#include <cmath>

namespace melonix
{
namespace
{
// "negative up to here, non-negative after": `look` samples on each side of the crossing
// (reference app.cpp:167-181 with look = 7, :203-217 with look = 3)
auto risingCrossing(std::span<const float> w, int idx, int look) -> bool
{
  const int n = static_cast<int>(w.size());
  if (idx < look || idx >= n - look - 1)
    return false;
  for (int j = 0; j < look; ++j)
    if (w[idx - j] >= 0 || w[idx + 1 + j] < 0)
      return false;
  return true;
}
} // namespace

auto segmentGrains(std::span<const float> wav) -> std::vector<Grain>
{
  std::vector<Grain> grains;
  const int n = static_cast<int>(wav.size());
  int start = 0;
  while (start < n - PreferredGrainSize - 1)
  {
    int cut = -1;
    // probe start+1500 + {0, 0, +1, -1, +2, -2, ...}: nearest crossing to the preferred size
    for (int i = 0; i < PreferredGrainSize && cut < 0; ++i)
    {
      const int idx = start + PreferredGrainSize + (i % 2 == 0 ? i / 2 : -i / 2);
      if (risingCrossing(wav, idx, 7))
        cut = idx;
    }
    // none within +-750: first looser crossing after start+2250
    for (int i = start + PreferredGrainSize + PreferredGrainSize / 2; cut < 0 && i < n - 1; ++i)
      if (risingCrossing(wav, i, 3))
        cut = i;
    if (cut < 0)
      break;
    grains.push_back({start, cut - start});
    start = cut;
  }
  return grains;
}

// ---- piece-wise linear warp through the markers (sorted by sample); reference app.cpp:1020-1122.
// The reference memoises these by int(val * sampleRate); in a fresh export every repeated key is
// hit with a bit-identical argument, so the memo does not change results and is not kept.
auto WarpMaps::sample2Time(int val) const -> double
{
  if (val <= 0)
    return 1. * val / sampleRate;
  int prevSample = 0;
  double prevTime = 0.0;
  for (const auto &mk : markers)
  {
    const double rightTime = prevTime + 1.0 * (mk.sample - prevSample) / sampleRate + mk.dTime;
    if (val > prevSample && val <= mk.sample)
      return prevTime + (val - prevSample) * (rightTime - prevTime) / (mk.sample - prevSample);
    prevSample = mk.sample;
    prevTime = rightTime;
  }
  return prevTime + 1. * (val - prevSample) / sampleRate;
}

auto WarpMaps::time2Sample(double val) const -> int
{
  if (val <= 0)
    return static_cast<int>(val * sampleRate);
  int prevSample = 0;
  double prevTime = 0.0;
  for (const auto &mk : markers)
  {
    const double rightTime = prevTime + 1.0 * (mk.sample - prevSample) / sampleRate + mk.dTime;
    if (val > prevTime && val <= rightTime)
      return static_cast<int>(prevSample + (val - prevTime) * (mk.sample - prevSample) / (rightTime - prevTime));
    prevSample = mk.sample;
    prevTime = rightTime;
  }
  return static_cast<int>(prevSample + (val - prevTime) * sampleRate);
}

auto WarpMaps::duration() const -> double { return sample2Time(static_cast<int>(nSamples - 1)); }

auto WarpMaps::time2PitchBend(double val) const -> float
{
  if (val <= 0)
    return 0;
  int prevSample = 0;
  double prevTime = 0.0, prevBend = 0.0;
  for (const auto &mk : markers)
  {
    const double rightTime = prevTime + 1.0 * (mk.sample - prevSample) / sampleRate + mk.dTime;
    if (val > prevTime && val <= rightTime)
      return static_cast<float>(prevBend + (val - prevTime) * (mk.pitchBend - prevBend) / (rightTime - prevTime));
    prevSample = mk.sample;
    prevTime = rightTime;
    prevBend = mk.pitchBend;
  }
  const double dur = duration();
  if (val > dur)
    return 0;
  return static_cast<float>(prevBend + (val - prevTime) * (0 - prevBend) / (dur - prevTime)); // ramps back to 0
}

// exportWav's loop (reference app.cpp:1201-1207) with process() reduced to its bookkeeping:
// which grain, which rate, how many samples, which sample closes the interpolation.
auto buildExportSchedule(std::span<const float> wav, int sampleRate, std::span<const MarkerView> markers,
                         std::span<const Grain> grains) -> RenderSchedule
{
  RenderSchedule s;
  const WarpMaps warp(markers, sampleRate, static_cast<int64_t>(wav.size()));
  const auto firstGrainAtOrAfter = [&](int sample) { // std::map::lower_bound on the grain starts
    return std::lower_bound(grains.begin(), grains.end(), sample,
                            [](const Grain &g, int v) { return g.start < v; });
  };
  int64_t out = 0;
  for (double cursor = 0.;;)
  {
    const float pitchBend = warp.time2PitchBend(cursor);
    const float rate = powf(2, pitchBend / 12); // float, as app.cpp:297
    const auto it = firstGrainAtOrAfter(warp.time2Sample(cursor));
    if (it == grains.end())
    {
      s.tailZeros = PreferredGrainSize; // app.cpp:303-309: 1500 zeros, then the export stops
      break;
    }
    // number of output samples: smallest i with trunc(float(i) * rate) >= len (app.cpp:314-322)
    int sz = 0;
    while (static_cast<size_t>(std::trunc(static_cast<double>(static_cast<float>(sz) * rate + 0.f))) <
           static_cast<size_t>(it->len))
      ++sz;
    const auto nextIt = firstGrainAtOrAfter(warp.time2Sample(cursor + 1. * sz / sampleRate));
    s.gStart.push_back(it->start);
    s.gLen.push_back(it->len);
    s.rate.push_back(rate);
    s.next.push_back(nextIt == grains.end() ? 0.f : wav[nextIt->start]);
    s.outOff.push_back(out);
    out += sz;
    const double dt = 1. * sz / sampleRate;
    if (dt <= 0.)
      break;
    cursor += dt;
  }
  s.outOff.push_back(out);
  return s;
}
} // namespace melonix

// ---------------------------------------------------------------------------------------------- C
using namespace melonix;

namespace
{
auto toViews(const mlxh_marker *m, int nm) -> std::vector<MarkerView>
{
  std::vector<MarkerView> v(static_cast<size_t>(nm > 0 ? nm : 0));
  for (int i = 0; i < nm; ++i)
    v[i] = {m[i].sample, m[i].dTime, m[i].pitchBend};
  return v;
}
} // namespace

extern "C" {
int mlxh_grain_segment(const float *wav, int64_t n, int32_t *g_start, int32_t *g_len, int cap)
{
  const auto g = segmentGrains({wav, static_cast<size_t>(n)});
  for (size_t i = 0; i < g.size() && static_cast<int>(i) < cap; ++i)
  {
    g_start[i] = g[i].start;
    g_len[i] = g[i].len;
  }
  return static_cast<int>(g.size());
}
double mlxh_sample2time(const mlxh_marker *m, int nm, int sr, int sample)
{
  const auto v = toViews(m, nm);
  return WarpMaps(v, sr, 0).sample2Time(sample);
}
int mlxh_time2sample(const mlxh_marker *m, int nm, int sr, double t)
{
  const auto v = toViews(m, nm);
  return WarpMaps(v, sr, 0).time2Sample(t);
}
double mlxh_duration(const mlxh_marker *m, int nm, int sr, int64_t n)
{
  const auto v = toViews(m, nm);
  return WarpMaps(v, sr, n).duration();
}
float mlxh_time2pitchbend(const mlxh_marker *m, int nm, int sr, int64_t n, double t)
{
  const auto v = toViews(m, nm);
  return WarpMaps(v, sr, n).time2PitchBend(t);
}
int mlxh_export_schedule(const float *wav, int64_t n, int sr, const mlxh_marker *m, int nm, const int32_t *g_start,
                         const int32_t *g_len, int ngrains, int32_t *s_gstart, int32_t *s_glen, float *s_rate,
                         int64_t *s_out_off, float *s_next, int cap, int *tail_zeros)
{
  const auto v = toViews(m, nm);
  std::vector<Grain> grains(static_cast<size_t>(ngrains));
  for (int i = 0; i < ngrains; ++i)
    grains[i] = {g_start[i], g_len[i]};
  const auto s = buildExportSchedule({wav, static_cast<size_t>(n)}, sr, v, grains);
  const int rows = static_cast<int>(s.gStart.size());
  if (tail_zeros)
    *tail_zeros = s.tailZeros;
  if (rows > cap)
    return -rows;
  for (int i = 0; i < rows; ++i)
  {
    s_gstart[i] = s.gStart[i];
    s_glen[i] = s.gLen[i];
    s_rate[i] = s.rate[i];
    s_next[i] = s.next[i];
    s_out_off[i] = s.outOff[i];
  }
  s_out_off[rows] = s.outOff[rows];
  return rows;
}
}
