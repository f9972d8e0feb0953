/* oracle/picks_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the waveform min/max pyramid of the reference:
 *   App::calcPicks            reference app.cpp:347-378
 *   App::getMinMaxFromRange   reference app.cpp:380-426
 * PINNED against the reference itself (app.cpp compiled unmodified, oracle/_ref/libapp_ref.so, driver
 * oracle/ref_app.cpp): identical pyramids and range-query results, bit for bit; the brute-force min/max
 * over aligned ranges is a second, analytic anchor (tests/test_oracle.py).
 *
 * Level l holds floor(n / 2^(l+1)) pairs (min, max) over samples [i 2^(l+1), (i+1) 2^(l+1)); levels
 * exist while n > 2^(l+1) (app.cpp:352, :365).  The levels are stored back to back; level_off[l] is
 * the index of the first pair of level l (in pairs), level_off[levels] the total.
 * std::min(a, b) is (b < a) ? b : a and std::max(a, b) is (a < b) ? b : a -- that is what decides the
 * result for NaN and signed zeros, so it is spelled out.
 */
#include "oracle.h"
#include <math.h>
#include <stddef.h>

static float smin(float a, float b) { return (b < a) ? b : a; }
static float smax(float a, float b) { return (a < b) ? b : a; }

int mlxo_picks_levels(int64_t n) {
  int lvl = 0;
  while (n > ((int64_t)1 << (lvl + 1))) ++lvl; /* app.cpp:352, :365 */
  return lvl;
}

int64_t mlxo_picks_layout(int64_t n, int64_t *level_off /* [levels + 1] */) {
  const int L = mlxo_picks_levels(n);
  int64_t off = 0;
  for (int l = 0; l < L; ++l) {
    level_off[l] = off;
    off += n / ((int64_t)1 << (l + 1)); /* app.cpp:356, :369 */
  }
  level_off[L] = off;
  return off;
}

void mlxo_picks_build(const float *wav, int64_t n, float *pairs /* [total][2] */, const int64_t *level_off) {
  const int L = mlxo_picks_levels(n);
  if (L == 0) return;
  for (int64_t i = 0; i < n / 2; ++i) { /* app.cpp:356-361 */
    pairs[2 * i] = smin(wav[2 * i], wav[2 * i + 1]);
    pairs[2 * i + 1] = smax(wav[2 * i], wav[2 * i + 1]);
  }
  for (int l = 1; l < L; ++l) { /* app.cpp:363-375 */
    const float *prev = pairs + 2 * level_off[l - 1];
    float *cur = pairs + 2 * level_off[l];
    const int64_t cnt = n / ((int64_t)1 << (l + 1));
    for (int64_t i = 0; i < cnt; ++i) {
      cur[2 * i] = smin(prev[2 * (2 * i)], prev[2 * (2 * i + 1)]);
      cur[2 * i + 1] = smax(prev[2 * (2 * i) + 1], prev[2 * (2 * i + 1) + 1]);
    }
  }
}

static void range_rec(const float *wav, int64_t n, const float *pairs, const int64_t *level_off, int L,
                      int start, int end, float *mn, float *mx) {
  if (start >= end) { /* app.cpp:382-387 */
    if (start >= 0 && start < (int)n) {
      *mn = *mx = wav[start];
    } else {
      *mn = *mx = 0.f;
    }
    return;
  }
  if (start < 0 || end < 0 || start >= (int)n || end >= (int)n) { /* :389-393 */
    *mn = *mx = 0.f;
    return;
  }
  if (end - start == 1) { /* :395-396 */
    *mn = *mx = wav[start];
    return;
  }
  const size_t lvl = (size_t)log2((double)(end - start)); /* :399 */
  const int lvlStart = start / (1 << lvl);                /* :401 */
  float a = 0.f, b = 0.f;                                 /* :402-408 */
  if (lvl - 1 < (size_t)L) {
    const int64_t cnt = level_off[lvl] - level_off[lvl - 1];
    if (lvlStart < (int)cnt) {
      a = pairs[2 * (level_off[lvl - 1] + lvlStart)];
      b = pairs[2 * (level_off[lvl - 1] + lvlStart) + 1];
    }
  }
  const int leftEnd = lvlStart * (1 << lvl); /* :410-416 */
  if (leftEnd >= start) {
    float l0, l1;
    range_rec(wav, n, pairs, level_off, L, start, leftEnd, &l0, &l1);
    a = smin(a, l0);
    b = smax(b, l1);
  }
  const int rightStart = (lvlStart + 1) * (1 << lvl); /* :418-424 */
  if (rightStart < end) {
    float r0, r1;
    range_rec(wav, n, pairs, level_off, L, rightStart, end, &r0, &r1);
    a = smin(a, r0);
    b = smax(b, r1);
  }
  *mn = a;
  *mx = b;
}

void mlxo_minmax_ranges(const float *wav, int64_t n, const float *pairs, const int64_t *level_off,
                        const int32_t *start_end, int count, float *out /* [count][2] */) {
  const int L = mlxo_picks_levels(n);
  for (int i = 0; i < count; ++i)
    range_rec(wav, n, pairs, level_off, L, start_end[2 * i], start_end[2 * i + 1], out + 2 * i, out + 2 * i + 1);
}
