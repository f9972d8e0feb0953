// melonix_b200/host/spec.hpp -- drop-in replacement for the reference's `Spec` (spec.hpp:11-39).
//
// Same public surface, same contract:
//   Spec(std::span<float> wav)   non-owning view; wav outlives Spec (the samples are copied to HBM)
//   ~Spec()                      joins the worker
//   getSpec(start, end) const    NON-BLOCKING: {} = "not ready, enqueued"; later SpectrSize/2 floats
// What changed underneath: instead of one 32768-point FFTW transform per job on the worker thread
// (reference spec.cpp:44-66,68-97) the worker drains ALL pending jobs and sends them to the GPU as
// one batched launch through the C ABI (include/melonix_gpu.h, mlx_spec_batch).  No CPU fallback:
// if no B200 is present the constructor throws.
#pragma once
#include "range.hpp"

#include <array>
#include <atomic>
#include <condition_variable>
#include <list>
#include <mutex>
#include <span>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

struct mlx_ctx;

class Spec
{
public:
  Spec(std::span<float> wav);
  ~Spec();
  auto getSpec(int start, int end) const -> std::vector<float>;

  // Extension used by SpecCache (not in the reference): the same job with the colour ramp of
  // SpecCache::populateTex fused into the GPU epilogue (reference spec-cache.cpp:77-96).
  // Non-blocking like getSpec: {} until the worker has produced the column for this gain `k`.
  using Rgb = std::array<unsigned char, 3>;
  auto getSpecRgb(int start, int end, float k) const -> std::vector<Rgb>;

  // FFT size of every job (reference: const SpectrSize = 8 * 4096, spec.cpp:8).  Overridable with
  // the MELONIX_SPECTR_SIZE environment variable (power of two, 512..32768) for tests / benchmarks.
  static auto spectrSize() -> int;

private:
  struct Entry
  {
    std::vector<float> spec;
    std::vector<Rgb> rgb;
    float rgbGain = 0.f;
    bool wantRgb = false;  // SpecCache asked for texels at rgbGain
    bool wantSpec = false; // somebody asked for the float magnitudes (getSpec)
    bool pending = false;  // queued or being computed by the worker
    std::list<Range>::iterator age;
  };

  std::span<float> wav;
  mlx_ctx *ctx = nullptr;
  int fftSize;
  mutable std::mutex mutex;
  mutable std::condition_variable wake;
  mutable std::unordered_set<Range, pair_hash> jobs;
  mutable std::unordered_map<Range, Entry, pair_hash> range2Spec;
  mutable std::list<Range> age;
  std::atomic<bool> running{false};
  std::thread thread;  // last member: started after everything above is constructed

  auto touch(const Range &key, Entry &e) const -> void;
  auto enqueue(const Range &key, bool wantRgb, float k) const -> void;
  auto run() -> void;
};
