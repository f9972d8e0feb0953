// melonix_b200/host/host_test.cpp -- headless driver for the C++ drop-in classes (built with
// -DMELONIX_HEADLESS; runs on the GPU box).  tests/test_host_cpp_gpu.py feeds it raw files and
// compares what it writes with the oracle.
//   host_test spec      wav.f32 jobs.i32 out.f32        Spec::getSpec for every job (async contract checked)
//   host_test recolour  wav.f32 jobs.i32 k out.u8       getSpec, then getSpecRgb(k): served at once from the cached floats
//   host_test speccache wav.f32 k width rangeTime out.u8   SpecCache::getTex for every column
//   host_test shard     wav.f32 fftN rate ngpu out.f32   one file by time range over ngpu GPUs (one host thread each)
//   host_test export    wav.f32 sampleRate semitones out.i16   segment -> schedule -> mlx_grain_render
#include "grain_schedule.hpp"
#include "spec-cache.hpp"
#include "spec.hpp"

#include "../../include/melonix_gpu.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>

template <typename T>
static auto readAll(const char *path) -> std::vector<T>
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f)
  {
    std::fprintf(stderr, "cannot open %s\n", path);
    std::exit(2);
  }
  const auto bytes = static_cast<size_t>(f.tellg());
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
  return v;
}

template <typename T>
static void writeAll(const char *path, const T *p, size_t n)
{
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char *>(p), static_cast<std::streamsize>(n * sizeof(T)));
}

static int runSpec(char **a)
{
  auto wav = readAll<float>(a[0]);
  const auto jobs = readAll<int32_t>(a[1]);
  const int count = static_cast<int>(jobs.size() / 2);
  const int half = Spec::spectrSize() / 2;
  Spec spec(std::span<float>{wav.data(), wav.size()});
  // KAT-3: the first call of a key is a miss and must return {} without blocking
  int firstEmpty = 0;
  for (int j = 0; j < count; ++j)
    firstEmpty += spec.getSpec(jobs[2 * j], jobs[2 * j + 1]).empty();
  std::vector<float> out(static_cast<size_t>(count) * half, -1.f);
  std::vector<char> done(count, 0);
  int remaining = count;
  for (int spin = 0; remaining > 0 && spin < 20000; ++spin)
  {
    for (int j = 0; j < count; ++j)
    {
      if (done[j])
        continue;
      const auto s = spec.getSpec(jobs[2 * j], jobs[2 * j + 1]);
      if (s.empty())
        continue;
      if (static_cast<int>(s.size()) != half)
      {
        std::fprintf(stderr, "bad length %zu\n", s.size());
        return 3;
      }
      std::memcpy(out.data() + static_cast<size_t>(j) * half, s.data(), sizeof(float) * half);
      done[j] = 1;
      --remaining;
    }
    if (remaining)
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  writeAll(a[2], out.data(), out.size());
  std::printf("spec: jobs=%d first_call_empty=%d remaining=%d half=%d\n", count, firstEmpty, remaining, half);
  return remaining == 0 ? 0 : 4;
}

// Brightness change on warm columns: once getSpec has delivered the floats, getSpecRgb must answer on
// its first call (host colour ramp on the cached floats, as the reference's populateTex does) instead
// of going back to the GPU and returning {} meanwhile.
static int runRecolour(char **a)
{
  auto wav = readAll<float>(a[0]);
  const auto jobs = readAll<int32_t>(a[1]);
  const float k = static_cast<float>(std::atof(a[2]));
  const int count = static_cast<int>(jobs.size() / 2);
  const int half = Spec::spectrSize() / 2;
  Spec spec(std::span<float>{wav.data(), wav.size()});
  int remaining = count;
  for (int spin = 0; remaining > 0 && spin < 20000; ++spin)
  {
    remaining = 0;
    for (int j = 0; j < count; ++j)
      remaining += spec.getSpec(jobs[2 * j], jobs[2 * j + 1]).empty();
    if (remaining)
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  std::vector<unsigned char> out(static_cast<size_t>(count) * half * 3);
  int immediate = 0;
  for (int j = 0; j < count && remaining == 0; ++j)
  {
    const auto rgb = spec.getSpecRgb(jobs[2 * j], jobs[2 * j + 1], k);
    if (static_cast<int>(rgb.size()) != half)
      continue;
    ++immediate;
    std::memcpy(out.data() + static_cast<size_t>(j) * half * 3, rgb.data(), static_cast<size_t>(half) * 3);
  }
  writeAll(a[3], out.data(), out.size());
  std::printf("recolour: jobs=%d not_ready=%d immediate=%d half=%d\n", count, remaining, immediate, half);
  return (remaining == 0 && immediate == count) ? 0 : 4;
}

static int runSpecCache(char **a)
{
  auto wav = readAll<float>(a[0]);
  const float k = static_cast<float>(std::atof(a[1]));
  const int width = std::atoi(a[2]);
  const double rangeTime = std::atof(a[3]);
  const int sampleRate = 48000;
  const int half = Spec::spectrSize() / 2;
  Spec spec(std::span<float>{wav.data(), wav.size()});
  SpecCache cache(spec, k, width, rangeTime, [&](double t) { return static_cast<int>(t * sampleRate); });
  std::vector<GLuint> names(width);
  auto &gl = gl_headless::state();
  int black = width;
  for (int spin = 0; black > 0 && spin < 20000; ++spin)
  {
    black = 0;
    for (int x = 0; x < width; ++x)
    {
      names[x] = cache.getTex(x * rangeTime / width + 1e-9);
      black += gl.tex[names[x]].size() != static_cast<size_t>(half) * 3;
    }
    if (black)
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  std::vector<unsigned char> out(static_cast<size_t>(width) * half * 3);
  for (int x = 0; x < width && black == 0; ++x)
    std::memcpy(out.data() + static_cast<size_t>(x) * half * 3, gl.tex[names[x]].data(), static_cast<size_t>(half) * 3);
  writeAll(a[4], out.data(), out.size());
  std::printf("speccache: columns=%d not_ready=%d half=%d\n", width, black, half);
  cache.clear();
  return black == 0 ? 0 : 4;
}

static int runExport(char **a)
{
  auto wav = readAll<float>(a[0]);
  const int sampleRate = std::atoi(a[1]);
  const double semis = std::atof(a[2]);
  const auto grains = melonix::segmentGrains(wav);
  // a constant shift needs a marker near each end
  std::vector<melonix::MarkerView> markers;
  if (semis != 0.0)
    markers = {{10, 0.0, semis}, {static_cast<int>(wav.size()) - 10, 0.0, semis}};
  const auto s = melonix::buildExportSchedule(wav, sampleRate, markers, grains);
  mlx_ctx *ctx = nullptr;
  if (mlx_create(&ctx, 0) != MLX_OK)
  {
    std::fprintf(stderr, "%s\n", mlx_last_error());
    return 5;
  }
  const float *ptr = wav.data();
  const int64_t n = static_cast<int64_t>(wav.size());
  int rc = mlx_upload_tracks(ctx, &ptr, &n, 1);
  const int rows = static_cast<int>(s.gStart.size());
  std::vector<int16_t> pcm16(static_cast<size_t>(s.outOff.back() + s.tailZeros));
  if (rc == MLX_OK)
    rc = mlx_grain_render(ctx, 0, s.gStart.data(), s.gLen.data(), s.rate.data(), s.outOff.data(), s.next.data(), rows,
                          s.tailZeros, nullptr, pcm16.data());
  if (rc != MLX_OK)
    std::fprintf(stderr, "%s\n", mlx_last_error());
  mlx_destroy(ctx);
  writeAll(a[3], pcm16.data(), pcm16.size());
  std::printf("export: grains=%zu rows=%d samples=%zu\n", grains.size(), rows, pcm16.size());
  return rc == MLX_OK ? 0 : 6;
}

// One long file phase-vocoded by time range across `ngpu` GPUs from a plain C++ host: one thread per
// GPU, each with its own context and its rank of one NCCL communicator behind the C ABI
// (mlx_comm_create); every rank hands in only its owned samples and receives only its owned output.
static int runShard(char **a)
{
  const auto wav = readAll<float>(a[0]);
  const int fftN = std::atoi(a[1]);
  const float rate = std::strtof(a[2], nullptr); // the pitch ratio itself (%.9g round-trips a float exactly)
  const int ngpu = std::atoi(a[3]);
  const int64_t n = static_cast<int64_t>(wav.size());
  std::vector<float> out(wav.size(), 0.f);
  char id[128];
  if (mlx_comm_unique_id(id) != MLX_OK)
  {
    std::fprintf(stderr, "%s\n", mlx_last_error());
    return 5;
  }
  mlx_pv_params p{};
  p.fftN = fftN;
  p.hop = fftN / 4;
  p.rate = rate;
  p.sample_rate = 48000.0;
  p.frame_begin = p.frame_end = -1;
  std::vector<int> rcs(ngpu, 0);
  std::vector<std::string> errs(ngpu);
  std::vector<std::thread> ranks;
  for (int r = 0; r < ngpu; ++r)
    ranks.emplace_back([&, r] {
      mlx_ctx *ctx = nullptr;
      mlx_comm *comm = nullptr;
      int rc = mlx_create(&ctx, r);
      if (rc == MLX_OK)
        rc = mlx_comm_create(&comm, ctx, id, ngpu, r);
      mlx_time_shard sh{};
      if (rc == MLX_OK)
        rc = mlx_shard_frames(n, fftN, fftN / 4, ngpu, r, &sh);
      if (rc == MLX_OK)
      {
        const float *own = wav.data() + sh.own_lo;
        float *y = out.data() + sh.own_lo;
        rc = mlx_pv_run_sharded(ctx, comm, &p, &own, 1, n, &y, nullptr, nullptr);
      }
      if (rc != MLX_OK)
        errs[r] = mlx_last_error();
      rcs[r] = rc;
      mlx_comm_destroy(comm);
      mlx_destroy(ctx);
    });
  for (auto &t : ranks)
    t.join();
  int bad = 0;
  for (int r = 0; r < ngpu; ++r)
    if (rcs[r] != MLX_OK)
    {
      std::fprintf(stderr, "rank %d: %s\n", r, errs[r].c_str());
      ++bad;
    }
  writeAll(a[4], out.data(), out.size());
  std::printf("shard: gpus=%d samples=%zu failed_ranks=%d\n", ngpu, out.size(), bad);
  return bad ? 6 : 0;
}

int main(int argc, char **argv)
{
  if (argc >= 5 && !std::strcmp(argv[1], "spec"))
    return runSpec(argv + 2);
  if (argc >= 6 && !std::strcmp(argv[1], "recolour"))
    return runRecolour(argv + 2);
  if (argc >= 7 && !std::strcmp(argv[1], "speccache"))
    return runSpecCache(argv + 2);
  if (argc >= 7 && !std::strcmp(argv[1], "shard"))
    return runShard(argv + 2);
  if (argc >= 6 && !std::strcmp(argv[1], "export"))
    return runExport(argv + 2);
  std::fprintf(stderr, "usage: host_test spec|recolour|speccache|shard|export ...\n");
  return 1;
}
