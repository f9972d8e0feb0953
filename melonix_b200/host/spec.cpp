// melonix_b200/host/spec.cpp -- see spec.hpp.  Replaces reference spec.cpp:10-106.
#include "spec.hpp"

#include "../../include/melonix_gpu.h"

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
    touch(key, it->second);
    return it->second.spec; // copy; empty while the job is still in flight
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
    if (e.wantRgb && e.rgbGain == k)
      return e.rgb;
    // first RGB request for this column (or the gain changed): recompute it with the ramp fused
    e.wantRgb = true;
    e.rgbGain = k;
    e.rgb.clear();
    jobs.insert(key);
    wake.notify_one();
    return {};
  }
  enqueue(key, true, k);
  return {};
}

// Worker.  The reference pops ONE arbitrary job, runs one FFT, sleeps 20 ms when idle
// (spec.cpp:68-97).  Here every wake-up takes the whole pending set and issues one batched launch
// per kind (float spectra; RGB columns grouped by gain).
auto Spec::run() -> void
{
  struct Job
  {
    Range key;
    bool rgb;
    float k;
  };
  std::vector<Job> batch;
  std::vector<int32_t> se;
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
        if (it != std::end(range2Spec))
          batch.push_back({key, it->second.wantRgb, it->second.rgbGain});
      }
      jobs.clear();
    }
    if (batch.empty())
      continue;

    const int count = static_cast<int>(batch.size());
    se.resize(2 * static_cast<size_t>(count));
    for (int j = 0; j < count; ++j)
    {
      se[2 * j] = batch[j].key.first;
      se[2 * j + 1] = batch[j].key.second;
    }
    out.resize(static_cast<size_t>(count) * half);
    const bool ok = mlx_spec_batch(ctx, 0, fftSize, se.data(), count, out.data()) == MLX_OK;

    // RGB columns: one launch per distinct gain (in practice one: SpecCache has a single k)
    std::vector<char> rgbDone(count, 0);
    std::vector<std::vector<Spec::Rgb>> rgbCols(count);
    for (int j = 0; ok && j < count; ++j)
    {
      if (!batch[j].rgb || rgbDone[j])
        continue;
      std::vector<int> idx;
      std::vector<int32_t> seK;
      for (int i = j; i < count; ++i)
        if (batch[i].rgb && !rgbDone[i] && batch[i].k == batch[j].k)
        {
          idx.push_back(i);
          seK.push_back(se[2 * i]);
          seK.push_back(se[2 * i + 1]);
        }
      rgbOut.resize(idx.size() * static_cast<size_t>(half) * 3);
      if (mlx_spec_batch_rgb(ctx, 0, fftSize, seK.data(), static_cast<int>(idx.size()), batch[j].k, rgbOut.data()) !=
          MLX_OK)
        break;
      for (size_t u = 0; u < idx.size(); ++u)
      {
        auto &col = rgbCols[idx[u]];
        col.resize(half);
        const unsigned char *src = rgbOut.data() + u * static_cast<size_t>(half) * 3;
        for (int b = 0; b < half; ++b)
          col[b] = {src[3 * b], src[3 * b + 1], src[3 * b + 2]};
        rgbDone[idx[u]] = 1;
      }
    }

    if (!ok)
      continue; // the reference has no error channel either; the column stays "not ready"
    std::lock_guard<std::mutex> lock(mutex);
    for (int j = 0; j < count; ++j)
    {
      const auto it = range2Spec.find(batch[j].key);
      if (it == std::end(range2Spec))
        continue; // evicted meanwhile (reference spec.cpp:91-93)
      it->second.spec.assign(out.begin() + static_cast<size_t>(j) * half,
                             out.begin() + static_cast<size_t>(j + 1) * half);
      if (rgbDone[j] && it->second.wantRgb && it->second.rgbGain == batch[j].k)
        it->second.rgb = std::move(rgbCols[j]);
    }
  }
}
