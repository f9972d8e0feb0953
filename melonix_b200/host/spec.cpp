// melonix_b200/host/spec.cpp -- see spec.hpp.  Replaces reference spec.cpp:10-106.
#include "spec.hpp"

#include "colour_ramp.hpp"

#include "../../include/melonix_gpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace
{
auto envInt(const char *name, int fallback) -> int
{
  const char *v = std::getenv(name);
  return v ? std::atoi(v) : fallback;
}

} // namespace

auto Spec::spectrSize() -> int
{
  const int n = envInt("MELONIX_SPECTR_SIZE", 8 * 4096);
  const bool pow2 = n > 0 && (n & (n - 1)) == 0;
  return (pow2 && n >= 512 && n <= 32768) ? n : 8 * 4096;
}

Spec::Spec(std::span<float> aWav) : wav(aWav), fftSize(spectrSize())
{
  if (mlx_create(&ctx, envInt("MELONIX_DEVICE", 0)) != MLX_OK)
    throw std::runtime_error(std::string("Spec: ") + mlx_last_error());
  const float *ptr = wav.data();
  const int64_t n = static_cast<int64_t>(wav.size());
  if (mlx_upload_tracks(ctx, &ptr, &n, 1) != MLX_OK)
  {
    const std::string msg = mlx_last_error();
    mlx_destroy(ctx);
    throw std::runtime_error("Spec: " + msg);
  }
  running = true;
  thread = std::thread(&Spec::run, this);
}

Spec::~Spec()
{
  {
    std::lock_guard<std::mutex> lock(mutex);
    running = false;
  }
  wake.notify_all();
  if (thread.joinable())
    thread.join();
  mlx_destroy(ctx);
}

auto Spec::touch(const Range &key, Entry &e) const -> void
{
  age.erase(e.age);
  age.push_front(key);
  e.age = std::begin(age);
}

// caller holds the mutex.  Same bookkeeping as the reference's miss path (spec.cpp:30-41):
// queue the job, create a placeholder, evict the least recently used entry beyond MaxRanges.
auto Spec::enqueue(const Range &key, bool wantRgb, float k) const -> void
{
  jobs.insert(key);
  age.push_front(key);
  Entry e;
  e.age = std::begin(age);
  e.wantRgb = wantRgb;
  e.wantSpec = !wantRgb;
  e.pending = true;
  e.rgbGain = k;
  range2Spec.emplace(key, std::move(e));
  if (range2Spec.size() > static_cast<size_t>(MaxRanges))
  {
    const Range oldest = age.back();
    range2Spec.erase(oldest);
    jobs.erase(oldest);
    age.pop_back();
  }
  wake.notify_one();
}

auto Spec::getSpec(int start, int end) const -> std::vector<float>
{
  const Range key{start, end};
  std::lock_guard<std::mutex> lock(mutex);
  const auto it = range2Spec.find(key);
  if (it != std::end(range2Spec))
  {
    Entry &e = it->second;
    touch(key, e);
    if (!e.wantSpec || (e.spec.empty() && !e.pending))
    {
      // the column so far exists as texels only (getSpecRgb), or its launch failed: queue the floats
      e.wantSpec = true;
      e.pending = true;
      jobs.insert(key);
      wake.notify_one();
    }
    return e.spec; // copy; empty while the job is still in flight
  }
  enqueue(key, false, 0.f);
  return {};
}

auto Spec::getSpecRgb(int start, int end, float k) const -> std::vector<Rgb>
{
  const Range key{start, end};
  std::lock_guard<std::mutex> lock(mutex);
  const auto it = range2Spec.find(key);
  if (it != std::end(range2Spec))
  {
    Entry &e = it->second;
    touch(key, e);
    if (e.wantRgb && e.rgbGain == k && !e.rgb.empty())
      return e.rgb;
    if (!e.wantRgb || e.rgbGain != k)
      e.rgb.clear();
    e.wantRgb = true;
    e.rgbGain = k;
    if (!e.spec.empty())
    {
      // the float spectrum is cached: recolour at once on the host, exactly as the reference does on
      // a brightness change (populateTex from getSpec's cached floats) -- no black flash, no relaunch
      e.rgb.resize(e.spec.size());
      std::transform(std::begin(e.spec), std::end(e.spec), std::begin(e.rgb),
                     [k](float m) { return melonix::rampTexel(m, k); });
      return e.rgb;
    }
    // texels-only column at a new gain (or first RGB request): one fused launch on the worker
    if (!e.pending)
    {
      e.pending = true;
      jobs.insert(key);
      wake.notify_one();
    }
    return {};
  }
  enqueue(key, true, k);
  return {};
}

// Worker.  The reference pops ONE arbitrary job, runs one FFT, sleeps 20 ms when idle
// (spec.cpp:68-97).  Here every wake-up takes the whole pending set and issues one batched launch
// per kind: float spectra for the columns somebody asked floats of, RGB columns (grouped by gain) for
// the ones SpecCache asked texels of -- a texels-only column never goes through the float batch.
// A failed launch drops the placeholders of its columns, so that the next getSpec / getSpecRgb
// misses and enqueues them again instead of returning {} forever.
auto Spec::run() -> void
{
  struct Job
  {
    Range key;
    bool spec, rgb;
    float k;
    bool failed = false;
  };
  std::vector<Job> batch;
  std::vector<int32_t> se;
  std::vector<int> idx;
  std::vector<float> out;
  std::vector<unsigned char> rgbOut;
  const int half = fftSize / 2;
  for (;;)
  {
    batch.clear();
    {
      std::unique_lock<std::mutex> lock(mutex);
      wake.wait(lock, [&] { return !running || !jobs.empty(); });
      if (!running)
        return;
      for (const auto &key : jobs)
      {
        const auto it = range2Spec.find(key);
        if (it == std::end(range2Spec))
          continue;
        const Entry &e = it->second;
        batch.push_back({key, e.wantSpec && e.spec.empty(), e.wantRgb && e.rgb.empty(), e.rgbGain});
      }
      jobs.clear();
    }
    if (batch.empty())
      continue;
    const int count = static_cast<int>(batch.size());

    // gathers the (start, end) pairs of the jobs selected by `pick` into se / idx
    const auto select = [&](auto pick) {
      se.clear();
      idx.clear();
      for (int j = 0; j < count; ++j)
        if (pick(batch[j]))
        {
          idx.push_back(j);
          se.push_back(batch[j].key.first);
          se.push_back(batch[j].key.second);
        }
      return static_cast<int>(idx.size());
    };

    // float spectra
    std::vector<int> specAt(count, -1);
    if (const int n = select([](const Job &j) { return j.spec; }); n > 0)
    {
      out.resize(static_cast<size_t>(n) * half);
      const bool ok = mlx_spec_batch(ctx, 0, fftSize, se.data(), n, out.data()) == MLX_OK;
      for (int u = 0; u < n; ++u)
      {
        specAt[idx[u]] = ok ? u : -1;
        batch[idx[u]].failed |= !ok;
      }
    }

    // RGB columns: one launch per distinct gain (in practice one: SpecCache has a single k)
    std::vector<std::vector<Spec::Rgb>> rgbCols(count);
    std::vector<char> rgbTried(count, 0);
    for (int j = 0; j < count; ++j)
    {
      if (!batch[j].rgb || rgbTried[j])
        continue;
      const float gain = batch[j].k;
      const int n = select([&](const Job &q) { return q.rgb && q.k == gain; });
      rgbOut.resize(static_cast<size_t>(n) * half * 3);
      const bool ok = mlx_spec_batch_rgb(ctx, 0, fftSize, se.data(), n, gain, rgbOut.data()) == MLX_OK;
      for (int u = 0; u < n; ++u)
      {
        rgbTried[idx[u]] = 1;
        batch[idx[u]].failed |= !ok;
        if (!ok)
          continue;
        auto &col = rgbCols[idx[u]];
        col.resize(half);
        const unsigned char *src = rgbOut.data() + static_cast<size_t>(u) * half * 3;
        for (int b = 0; b < half; ++b)
          col[b] = {src[3 * b], src[3 * b + 1], src[3 * b + 2]};
      }
    }

    std::lock_guard<std::mutex> lock(mutex);
    for (int j = 0; j < count; ++j)
    {
      const auto it = range2Spec.find(batch[j].key);
      if (it == std::end(range2Spec))
        continue; // evicted meanwhile (reference spec.cpp:91-93)
      Entry &e = it->second;
      e.pending = jobs.count(batch[j].key) != 0; // a request that arrived during the launch stays queued
      if (specAt[j] >= 0)
        e.spec.assign(out.begin() + static_cast<size_t>(specAt[j]) * half,
                      out.begin() + static_cast<size_t>(specAt[j] + 1) * half);
      if (!rgbCols[j].empty() && e.wantRgb && e.rgbGain == batch[j].k)
        e.rgb = std::move(rgbCols[j]);
      if (batch[j].failed && e.spec.empty() && e.rgb.empty())
      {
        // nothing usable was produced (transient CUDA error): forget the placeholder; the caller's
        // next request is a miss and queues the column again (the reference has no failure mode here)
        age.erase(e.age);
        range2Spec.erase(it);
      }
    }
  }
}
