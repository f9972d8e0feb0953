// melonix_b200/csrc/grain_seg.cuh -- bit logic of the grain segmentation kernels (K8).
//
// App::preproc (reference app.cpp:156-235) cuts the track into grains at negative->positive zero
// crossings: position idx qualifies with look-around `look` when
//     look <= idx < n - look - 1,   !(w[idx-j] >= 0)  and  !(w[idx+1+j] < 0)   for j in [0, look)
// (app.cpp:167-181 with look = 7, :203-217 with look = 3).  From `start` the reference probes
// idx = start + 1500 + {0, 0, +1, -1, +2, -2, ..., +749, -749} with look 7 and takes the first hit
// (app.cpp:164-166), else scans forward from start + 2250 with look 3 (app.cpp:198-201).
//
// On the device the predicate is evaluated for every sample at once into two bit arrays (Z7, Z3;
// bit b of word k <-> idx = 32 k + b), and the chain over grains becomes "nearest set bit to
// start + 1500, ties to the right" on Z7 / "first set bit at or after start + 2250" on Z3.
// These helpers are host/device so that tests/host/grain_seg_emul.cpp runs the same bit logic on
// the CPU against the oracle's restatement.
#pragma once
#include <stdint.h>

#include "fft.cuh"  // MLX_HD

namespace mlx {

constexpr int kGrainPreferred = 1500;                       // preferredGrainSize, app.cpp:19
constexpr int kGrainHalfSpan = kGrainPreferred / 2 - 1;     // probes reach +-749 around start + 1500
constexpr int kGrainWindowWords = 64;                       // words that cover the 1499-sample window from its first word

// Word k of the crossing bit array for look-around `look` from the sign words around it:
//   l_prev, l_cur: bit b <-> !(w[i] >= 0) for i = 32(k-1)+b, 32k+b
//   r_cur, r_next: bit b <-> !(w[i] <  0) for i = 32k+b, 32(k+1)+b
MLX_HD uint32_t seg_cross_word(uint32_t l_prev, uint32_t l_cur, uint32_t r_cur, uint32_t r_next, long long k,
                               long long n, int look) {
  const unsigned long long L = (unsigned long long)l_prev | ((unsigned long long)l_cur << 32);
  const unsigned long long R = (unsigned long long)r_cur | ((unsigned long long)r_next << 32);
  uint32_t z = 0xffffffffu;
  for (int j = 0; j < look; ++j) {
    z &= (uint32_t)(L >> (32 - j));  // bit b = !(w[32k + b - j] >= 0)
    z &= (uint32_t)(R >> (1 + j));   // bit b = !(w[32k + b + 1 + j] < 0)
  }
  // range: look <= idx < n - look - 1
  const long long p0 = 32 * k, lo = look, hi = n - look - 2;  // inclusive bounds
  if (hi < lo || p0 > hi || p0 + 31 < lo) return 0u;
  if (lo > p0) z &= 0xffffffffu << (int)(lo - p0);
  if (hi < p0 + 31) z &= 0xffffffffu >> (int)(31 - (hi - p0));
  return z;
}

// Best probe of the reference's search order inside one word (positions p0 .. p0+31) restricted to
// [c - 749, c + 749]: key = 2 |p - c| + (p < c), smaller = probed earlier; 0xffffffff when none.
MLX_HD uint32_t seg_word_key(uint32_t word, int p0, int c) {
  const int lo = c - kGrainHalfSpan, hi = c + kGrainHalfSpan;
  if (word == 0u || p0 > hi || p0 + 31 < lo) return 0xffffffffu;
  uint32_t m = word;
  if (lo > p0) m &= 0xffffffffu << (lo - p0);
  if (hi < p0 + 31) m &= 0xffffffffu >> (31 - (hi - p0));
  uint32_t ge = 0xffffffffu;  // bits with p >= c
  if (c > p0) ge = (c - p0 >= 32) ? 0u : (0xffffffffu << (c - p0));
  const uint32_t m_hi = m & ge, m_lo = m & ~ge;
  uint32_t key = 0xffffffffu;
  if (m_hi) {
#ifdef __CUDA_ARCH__
    const int b = __ffs((int)m_hi) - 1;
#else
    const int b = __builtin_ctz(m_hi);
#endif
    key = 2u * (uint32_t)(p0 + b - c);
  }
  if (m_lo) {
#ifdef __CUDA_ARCH__
    const int b = 31 - __clz((int)m_lo);
#else
    const int b = 31 - __builtin_clz(m_lo);
#endif
    const uint32_t k2 = 2u * (uint32_t)(c - (p0 + b)) + 1u;
    key = k2 < key ? k2 : key;
  }
  return key;
}

// ---- warp-cooperative form of the same search (used by the chain kernel) -----------------------
// The 1499-sample window [lo, hi] = [c - 749, c + 749] starts at bit s = lo & 31 of word w0 = lo >> 5.
// Lane l (0..23) owns the 64 bits [64 l, 64 l + 64) relative to bit 0 of word w0: two staged words.
// seg_lane_split() masks the lane's bits to the window and splits them at the centre c (relative
// position s + 749): `ge` = probes at or after c, `lt` = probes before c.  The nearest probe on each
// side is then the lowest set bit of the lowest lane with ge != 0 and the highest set bit of the
// highest lane with lt != 0 (two ballots on the device); seg_pick() applies the reference's order
// (+d is probed before -d, app.cpp:166) and returns the position relative to bit 0 of word w0, or -1.
struct SegSplit {
  unsigned long long ge, lt;
};
MLX_HD SegSplit seg_lane_split(unsigned long long bits, int lane, int s) {
  const int first = 64 * lane;                       // relative position of the lane's bit 0
  const int last_rel = s + 2 * kGrainHalfSpan;       // relative position of hi
  const int c_rel = s + kGrainHalfSpan;              // relative position of c
  unsigned long long m = bits;
  if (lane == 0) m &= ~0ull << s;                    // below lo (s < 32: only lane 0 is cut)
  if (first > last_rel) m = 0ull;                    // lanes past hi
  else if (first + 63 > last_rel) m &= ~0ull >> (63 - (last_rel - first));
  unsigned long long gem;                            // bits at or after c
  if (first >= c_rel) gem = ~0ull;
  else if (first + 63 < c_rel) gem = 0ull;
  else gem = ~0ull << (c_rel - first);
  return SegSplit{m & gem, m & ~gem};
}
MLX_HD int seg_ctz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
MLX_HD int seg_clz64(unsigned long long v) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}
// lane_ge / lane_lt: lowest lane with ge != 0 / highest lane with lt != 0 (or -1); w_ge / w_lt: their words
MLX_HD int seg_pick(int lane_ge, unsigned long long w_ge, int lane_lt, unsigned long long w_lt, int s) {
  const int c_rel = s + kGrainHalfSpan;
  const int p_ge = lane_ge >= 0 ? 64 * lane_ge + seg_ctz64(w_ge) : -1;
  const int p_lt = lane_lt >= 0 ? 64 * lane_lt + 63 - seg_clz64(w_lt) : -1;
  if (p_ge < 0) return p_lt;
  if (p_lt < 0) return p_ge;
  return (p_ge - c_rel <= c_rel - p_lt) ? p_ge : p_lt;
}

// index the key stands for
MLX_HD int seg_key_index(uint32_t key, int c) {
  const int d = (int)(key >> 1);
  return (key & 1u) ? c - d : c + d;
}

}  // namespace mlx
