// melonix_b200/csrc/picks_kernels.cu -- K9: waveform min/max pyramid for sm_100a.
//
// Replaces App::calcPicks (reference app.cpp:347-378) and App::getMinMaxFromRange (app.cpp:380-426).
// Level l holds floor(n / 2^(l+1)) pairs (min, max) over samples [i 2^(l+1), (i+1) 2^(l+1)); levels
// exist while n > 2^(l+1).  The reference builds level l from level l-1 with std::min / std::max on
// the (first, second) halves, re-reading each level from memory; here one CTA reduces a tile of 4096
// samples through levels 0..11 in registers (four float4 per thread, lane-contiguous), warp shuffles and
// one shared-memory hop, writing every level as it appears -- each sample is read once (4 B) and the pyramid written once
// (8 B per sample in total): an HBM-bound streaming kernel.  The few levels above the tile are
// finished by one CTA.  std::min(a, b) is (b < a) ? b : a and std::max(a, b) is (a < b) ? b : a with
// a = the LOWER-indexed half: that order decides the result for NaN and signed zeros and is kept.
#include <cuda_runtime.h>

#include "kernels.h"

namespace mlx {

__device__ __forceinline__ float smin(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float smax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float2 comb(float2 lo, float2 hi) { return make_float2(smin(lo.x, hi.x), smax(lo.y, hi.y)); }

constexpr int kPickTile = 4096;      // samples per CTA
constexpr int kPickTileLevels = 12;  // levels 0..11 are complete inside a tile

__global__ void __launch_bounds__(256) picks_tile_kernel(const PicksArgs* __restrict__ tracks) {
  __shared__ float2 s_part[32];  // level-6 entries (128 samples each) of the tile, in sample order
  __shared__ PicksArgs a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tile = blockIdx.x;
  if (tile * kPickTile >= tracks[blockIdx.y].n) return;  // grid.x is sized for the longest track
  if (tid < (int)(sizeof(PicksArgs) / sizeof(int)))
    reinterpret_cast<int*>(&a)[tid] = reinterpret_cast<const int*>(tracks + blockIdx.y)[tid];
  __syncthreads();
  const bool aligned = (reinterpret_cast<unsigned long long>(a.pairs) & 15ull) == 0ull;  // level 0 starts at pairs
  // Lane-contiguous I/O: in round v the thread takes float4 number f = v*256 + tid of the tile (samples
  // 4f .. 4f+3), so every load and every store of levels 0 and 1 is one contiguous run per warp.
  // A span that crosses n produces no entry, so the values read past n never matter.
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int f = v * 256 + tid;
    const long long s0 = tile * kPickTile + 4LL * f;
    float4 q;
    if (s0 + 4 <= a.n) {
      q = __ldg(reinterpret_cast<const float4*>(a.x + s0));
    } else {
      q.x = s0 < a.n ? __ldg(a.x + s0) : 0.f;
      q.y = s0 + 1 < a.n ? __ldg(a.x + s0 + 1) : 0.f;
      q.z = s0 + 2 < a.n ? __ldg(a.x + s0 + 2) : 0.f;
      q.w = 0.f;
    }
    const float2 p0 = make_float2(smin(q.x, q.y), smax(q.x, q.y));
    const float2 p1 = make_float2(smin(q.z, q.w), smax(q.z, q.w));
    if (a.levels > 0) {  // level 0: entries 2f, 2f+1 of the tile
      const long long idx = tile * (kPickTile >> 1) + 2LL * f;
      const long long cnt = a.n >> 1;
      float2* out = a.pairs + a.level_off[0] + idx;
      if (idx + 2 <= cnt && aligned) {
        *reinterpret_cast<float4*>(out) = make_float4(p0.x, p0.y, p1.x, p1.y);
      } else {
        if (idx < cnt) out[0] = p0;
        if (idx + 1 < cnt) out[1] = p1;
      }
    }
    float2 e = comb(p0, p1);
    if (a.levels > 1) {  // level 1: entry f
      const long long idx = tile * (kPickTile >> 2) + f;
      if (idx < (a.n >> 2)) a.pairs[a.level_off[1] + idx] = e;
    }
    // levels 2..6: pairs of lanes, the lower lane holds the first half
#pragma unroll
    for (int sft = 0; sft < 5; ++sft) {
      const int l = 2 + sft;
      float2 o;
      o.x = __shfl_down_sync(0xffffffffu, e.x, 1 << sft);
      o.y = __shfl_down_sync(0xffffffffu, e.y, 1 << sft);
      e = comb(e, o);
      if (l < a.levels && (lane & ((2 << sft) - 1)) == 0) {
        const long long idx = tile * (kPickTile >> (l + 1)) + (f >> (sft + 1));
        if (idx < (a.n >> (l + 1))) a.pairs[a.level_off[l] + idx] = e;
      }
    }
    if (lane == 0) s_part[v * 8 + warp] = e;  // level 6: samples [128 (v*8 + warp), +128) of the tile
  }
  __syncthreads();
  if (warp == 0) {  // levels 7..11 from the 32 level-6 entries
    float2 e = s_part[lane];
#pragma unroll
    for (int sft = 0; sft < 5; ++sft) {
      const int l = 7 + sft;
      float2 o;
      o.x = __shfl_down_sync(0xffffffffu, e.x, 1 << sft);
      o.y = __shfl_down_sync(0xffffffffu, e.y, 1 << sft);
      e = comb(e, o);
      if (l < a.levels && (lane & ((2 << sft) - 1)) == 0) {
        const long long idx = tile * (kPickTile >> (l + 1)) + (lane >> (sft + 1));
        if (idx < (a.n >> (l + 1))) a.pairs[a.level_off[l] + idx] = e;
      }
    }
  }
}

// levels >= 12: one CTA, level after level (at most n / 8192 entries on the first of them)
__global__ void __launch_bounds__(1024) picks_top_kernel(const PicksArgs* __restrict__ tracks) {
  const PicksArgs a = tracks[blockIdx.x];
  for (int l = kPickTileLevels; l < a.levels; ++l) {
    const long long cnt = a.n >> (l + 1);
    const float2* prev = a.pairs + a.level_off[l - 1];
    float2* cur = a.pairs + a.level_off[l];
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) cur[i] = comb(prev[2 * i], prev[2 * i + 1]);
    __syncthreads();  // one CTA: its global writes are visible to its own threads after the barrier
  }
}

// App::getMinMaxFromRange (app.cpp:380-426), one thread per (start, end).  The reference recurses on the
// right remainder and folds outer-first: min(A0, min(A1, ...)); the per-level terms are collected and
// folded from the innermost outwards so that the std::min / std::max argument order is the reference's.
__global__ void __launch_bounds__(128) minmax_ranges_kernel(const PicksArgs a, const int* __restrict__ start_end,
                                                            int count, float2* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= count) return;
  int start = start_end[2 * r];
  const int end = start_end[2 * r + 1];
  const int n = (int)a.n;
  float2 term[32];
  bool has_left[32];
  float left[32];
  int depth = 0;
  float2 tail;
  while (true) {
    if (start >= end) {  // :382-387
      const float w = (start >= 0 && start < n) ? a.x[start] : 0.f;
      tail = make_float2(w, w);
      break;
    }
    if (start < 0 || end < 0 || start >= n || end >= n) {  // :389-393
      tail = make_float2(0.f, 0.f);
      break;
    }
    if (end - start == 1) {  // :395-396
      tail = make_float2(a.x[start], a.x[start]);
      break;
    }
    const int lvl = 31 - __clz(end - start);  // (size_t)std::log2(end - start), :399
    const int lvlStart = start >> lvl;         // :401 (start >= 0)
    float2 mm = make_float2(0.f, 0.f);         // :402-408
    if (lvl - 1 < a.levels && lvlStart < (a.n >> lvl)) mm = a.pairs[a.level_off[lvl - 1] + lvlStart];
    const int leftEnd = lvlStart << lvl;       // :410-416: recursion on (start, leftEnd) with leftEnd <= start
    has_left[depth] = leftEnd >= start;
    left[depth] = a.x[start];
    term[depth] = mm;
    ++depth;
    const long long rightStart = ((long long)lvlStart + 1) << lvl;  // :418-424
    if (rightStart < end) {
      start = (int)rightStart;
      continue;
    }
    // no right remainder: fold what has been collected
    --depth;
    tail = term[depth];
    if (has_left[depth]) tail = make_float2(smin(tail.x, left[depth]), smax(tail.y, left[depth]));
    break;
  }
  while (depth > 0) {
    --depth;
    float2 mm = term[depth];
    if (has_left[depth]) mm = make_float2(smin(mm.x, left[depth]), smax(mm.y, left[depth]));
    tail = make_float2(smin(mm.x, tail.x), smax(mm.y, tail.y));
  }
  out[r] = tail;
}

cudaError_t launch_picks_build(const PicksArgs* tracks_dev, int ntracks, long long max_n, int max_levels,
                               cudaStream_t st) {
  if (ntracks <= 0 || max_levels <= 0) return cudaSuccess;
  const long long tiles = (max_n + kPickTile - 1) / kPickTile;
  dim3 grid((unsigned)tiles, ntracks);
  picks_tile_kernel<<<grid, 256, 0, st>>>(tracks_dev);
  if (max_levels > kPickTileLevels) picks_top_kernel<<<ntracks, 1024, 0, st>>>(tracks_dev);
  return cudaGetLastError();
}

cudaError_t launch_minmax_ranges(const PicksArgs& a, const int* start_end_dev, int count, float* out_dev,
                                 cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  minmax_ranges_kernel<<<(count + 127) / 128, 128, 0, st>>>(a, start_end_dev, count, reinterpret_cast<float2*>(out_dev));
  return cudaGetLastError();
}

}  // namespace mlx
