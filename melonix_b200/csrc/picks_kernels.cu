// melonix_b200/csrc/picks_kernels.cu -- K9: waveform min/max pyramid for sm_100a.
//
// Replaces App::calcPicks (reference app.cpp:347-378) and App::getMinMaxFromRange (app.cpp:380-426).
// Level l holds floor(n / 2^(l+1)) pairs (min, max) over samples [i 2^(l+1), (i+1) 2^(l+1)); levels
// exist while n > 2^(l+1).  The reference builds level l from level l-1 with std::min / std::max on
// the (first, second) halves, re-reading each level from memory; here one CTA reduces a tile of 4096
// samples through levels 0..11 in registers (16 samples per thread), warp shuffles and one shared-memory
// hop, writing every level as it appears -- each sample is read once (4 B) and the pyramid written once
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
  __shared__ float2 s_warp[8];
  __shared__ PicksArgs a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tile = blockIdx.x;
  if (tile * kPickTile >= tracks[blockIdx.y].n) return;  // grid.x is sized for the longest track
  if (tid < (int)(sizeof(PicksArgs) / sizeof(int)))
    reinterpret_cast<int*>(&a)[tid] = reinterpret_cast<const int*>(tracks + blockIdx.y)[tid];
  __syncthreads();
  const long long base = tile * kPickTile + tid * 16;
  // 16 consecutive samples; a span that crosses n produces no entry, so its values never matter
  float s[16];
  if (base + 16 <= a.n) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(a.x + base) + v);
      s[4 * v] = q.x; s[4 * v + 1] = q.y; s[4 * v + 2] = q.z; s[4 * v + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = (base + i < a.n) ? __ldg(a.x + base + i) : 0.f;
  }
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = make_float2(smin(s[2 * i], s[2 * i + 1]), smax(s[2 * i], s[2 * i + 1]));
  // levels 0..3 from registers: level l has 8 >> l entries per thread
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int per = 8 >> l;
    if (l < a.levels) {
      const long long first = tile * (kPickTile >> (l + 1)) + (long long)tid * per;
      const long long cnt = a.n >> (l + 1);
      float2* out = a.pairs + a.level_off[l] + first;
      if (first + per <= cnt && per >= 2 && (reinterpret_cast<unsigned long long>(out) & 15ull) == 0ull) {
#pragma unroll
        for (int i = 0; i < per; i += 2)
          *reinterpret_cast<float4*>(out + i) = make_float4(p[i].x, p[i].y, p[i + 1].x, p[i + 1].y);
      } else {
#pragma unroll
        for (int i = 0; i < per; ++i)
          if (first + i < cnt) out[i] = p[i];
      }
    }
#pragma unroll
    for (int i = 0; i < per / 2; ++i) p[i] = comb(p[2 * i], p[2 * i + 1]);
  }
  // p[0] is now the thread's level-3 entry (16 samples).  Levels 4..8: pairs of lanes, the lower lane
  // holds the first half.
  float2 v = p[0];
#pragma unroll
  for (int sft = 0; sft < 5; ++sft) {
    const int l = 4 + sft;
    float2 o;
    o.x = __shfl_down_sync(0xffffffffu, v.x, 1 << sft);
    o.y = __shfl_down_sync(0xffffffffu, v.y, 1 << sft);
    v = comb(v, o);
    if (l < a.levels && (lane & ((2 << sft) - 1)) == 0) {
      const long long idx = tile * (kPickTile >> (l + 1)) + (tid >> (sft + 1));
      if (idx < (a.n >> (l + 1))) a.pairs[a.level_off[l] + idx] = v;
    }
  }
  if (lane == 0) s_warp[warp] = v;  // level 8: 512 samples per warp
  __syncthreads();
  if (warp == 0) {
    v = s_warp[lane & 7];
#pragma unroll
    for (int sft = 0; sft < 3; ++sft) {
      const int l = 9 + sft;
      float2 o;
      o.x = __shfl_down_sync(0xffffffffu, v.x, 1 << sft);
      o.y = __shfl_down_sync(0xffffffffu, v.y, 1 << sft);
      v = comb(v, o);
      if (l < a.levels && lane < 8 && (lane & ((2 << sft) - 1)) == 0) {
        const long long idx = tile * (kPickTile >> (l + 1)) + (lane >> (sft + 1));
        if (idx < (a.n >> (l + 1))) a.pairs[a.level_off[l] + idx] = v;
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
