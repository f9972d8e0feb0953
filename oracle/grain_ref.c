/* oracle/grain_ref.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the reference's time-domain pitch shifter (this is what melonix actually ships where
 * the north-star says "pitch-shift inner loop"):
 *   grain segmentation   App::preproc   reference app.cpp:156-235
 *   warp maps            sample2Time / time2Sample / duration / time2PitchBend  app.cpp:1020-1122
 *   grain resampler      App::process   app.cpp:294-345
 *   export driver        App::exportWav app.cpp:1194-1215
 * PINNED against the reference itself: app.cpp + save-wav.cpp compile unmodified against the no-op
 * UI / audio / codec headers of oracle/shim_app (oracle/_ref/libapp_ref.so, driver oracle/ref_app.cpp)
 * and give identical grains, float samples, int16 samples and warp-map values (tests/test_oracle.py);
 * KAT-4 / KAT-5 are the analytic anchors.  The reference has no tests of its own.
 *
 * The reference memoises the warp maps by int(val*sampleRate) (app.cpp:1058-1060, 1093-1095).
 * In a fresh export every repeated key is hit with a bit-identical argument (cursor + sz/sr is
 * recomputed identically by the caller, app.cpp:325 and :1206), so the memo is transparent and is
 * not restated.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { kPreferredGrainSize = 1500 }; /* app.cpp:19 */

static int zero_cross(const float *w, int64_t n, int idx, int look) {
  /* app.cpp:167-181 (look=7) and :203-217 (look=3) */
  if (idx < look) return 0;
  if (idx >= (int)(n - look - 1)) return 0;
  for (int j = 0; j < look; ++j) {
    if (w[idx - j] >= 0) return 0;
    if (w[idx + 1 + j] < 0) return 0;
  }
  return 1;
}

int mlxo_grain_segment(const float *wav, int64_t n, int32_t *g_start, int32_t *g_len, int cap) {
  int count = 0;
  int start = 0;
  while (start < (int)(n - kPreferredGrainSize - 1)) { /* app.cpp:161 */
    int found = 0;
    for (int i = 0; i < kPreferredGrainSize; ++i) { /* :164 */
      const int idx = start + kPreferredGrainSize + (i % 2 == 0 ? i / 2 : -i / 2); /* :166 */
      if (zero_cross(wav, n, idx, 7)) {
        if (count < cap) {
          g_start[count] = start;
          g_len[count] = idx - start;
        }
        ++count;
        start = idx;
        found = 1;
        break;
      }
    }
    if (!found) { /* :194-231 */
      for (int i = start + kPreferredGrainSize + kPreferredGrainSize / 2; i < (int)(n - 1); ++i) {
        if (zero_cross(wav, n, i, 3)) {
          if (count < cap) {
            g_start[count] = start;
            g_len[count] = i - start;
          }
          ++count;
          start = i;
          found = 1;
          break;
        }
      }
      if (!found) break;
    }
  }
  return count;
}

double mlxo_sample2time(const mlxo_marker *m, int nm, int sr, int val) { /* app.cpp:1020-1050 */
  if (val <= 0) return 1. * val / sr;
  int prevSample = 0;
  double prevTime = 0.0;
  for (int i = 0; i < nm; ++i) {
    const double rightTime = prevTime + 1.0 * (m[i].sample - prevSample) / sr + m[i].dTime;
    if (val > prevSample && val <= m[i].sample)
      return prevTime + (val - prevSample) * (rightTime - prevTime) / (m[i].sample - prevSample);
    prevSample = m[i].sample;
    prevTime = rightTime;
  }
  return prevTime + 1. * (val - prevSample) / sr;
}

int mlxo_time2sample(const mlxo_marker *m, int nm, int sr, double val) { /* app.cpp:1052-1082 */
  if (val <= 0) return (int)(val * sr);
  int prevSample = 0;
  double prevTime = 0.0;
  for (int i = 0; i < nm; ++i) {
    const double rightTime = prevTime + 1.0 * (m[i].sample - prevSample) / sr + m[i].dTime;
    if (val > prevTime && val <= rightTime)
      return (int)(prevSample +
                   (val - prevTime) * (m[i].sample - prevSample) / (rightTime - prevTime));
    prevSample = m[i].sample;
    prevTime = rightTime;
  }
  return (int)(prevSample + (val - prevTime) * sr);
}

double mlxo_duration(const mlxo_marker *m, int nm, int sr, int64_t n) { /* app.cpp:1084-1087 */
  return mlxo_sample2time(m, nm, sr, (int)(n - 1));
}

float mlxo_time2pitchbend(const mlxo_marker *m, int nm, int sr, int64_t n, double val) {
  /* app.cpp:1089-1122 */
  if (val <= 0) return 0;
  int prevSample = 0;
  double prevTime = 0.0, prevPitchBend = 0.0;
  for (int i = 0; i < nm; ++i) {
    const double rightTime = prevTime + 1.0 * (m[i].sample - prevSample) / sr + m[i].dTime;
    if (val > prevTime && val <= rightTime)
      return (float)(prevPitchBend +
                     (val - prevTime) * (m[i].pitchBend - prevPitchBend) / (rightTime - prevTime));
    prevSample = m[i].sample;
    prevTime = rightTime;
    prevPitchBend = m[i].pitchBend;
  }
  const double dur = mlxo_duration(m, nm, sr, n);
  if (val > dur) return 0;
  return (float)(prevPitchBend + (val - prevTime) * (0 - prevPitchBend) / (dur - prevTime));
}

/* std::map::lower_bound over grain starts (ascending): first grain with start >= sample */
static int lower_bound_grain(const int32_t *g_start, int ng, int sample) {
  int lo = 0, hi = ng;
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (g_start[mid] < sample) lo = mid + 1; else hi = mid;
  }
  return lo;
}

int64_t mlxo_grain_export(const float *wav, int64_t n, int sr, const mlxo_marker *m, int nm,
                          const int32_t *g_start, const int32_t *g_len, int ng, float *pcm,
                          int16_t *pcm16, int64_t cap, int32_t *s_gstart, int32_t *s_glen,
                          float *s_rate, int64_t *s_out_off, float *s_next, int *nsched,
                          int cap_sched) {
  const float bias = 0.f; /* app.hpp:66, never assigned */
  int64_t len = 0;
  int ns = 0;
  for (double cursor = 0.;;) { /* exportWav loop, app.cpp:1201-1207 */
    /* ---- App::process(cursor, pcm), app.cpp:294-345 ---- */
    const float pitchBend = mlxo_time2pitchbend(m, nm, sr, n, cursor);
    const float rate = powf(2, pitchBend / 12);
    const int gi = lower_bound_grain(g_start, ng, mlxo_time2sample(m, nm, sr, cursor));
    double dt;
    if (gi == ng) { /* :303-309 */
      for (int i = 0; i < kPreferredGrainSize; ++i, ++len)
        if (pcm && len < cap) pcm[len] = 0.f;
      dt = 0;
    } else {
      const float *grain = wav + g_start[gi];
      const size_t gsz = (size_t)g_len[gi];
      int sz = 0; /* :313-322 */
      for (int i = 0;; ++i) {
        double idxF;
        modf((double)(i * rate + bias), &idxF);
        if ((size_t)idxF >= gsz) break;
        ++sz;
      }
      float next = 0.f; /* :323-328 */
      {
        const int g2 = lower_bound_grain(g_start, ng,
                                         mlxo_time2sample(m, nm, sr, cursor + 1. * sz / sr));
        if (g2 != ng) next = wav[g_start[g2]];
      }
      if (ns < cap_sched && s_gstart) {
        s_gstart[ns] = g_start[gi];
        s_glen[ns] = g_len[gi];
        s_rate[ns] = rate;
        s_out_off[ns] = len;
        s_next[ns] = next;
      }
      ++ns;
      sz = 0; /* :331-343 */
      for (int i = 0;; ++i) {
        float idxF;
        const float curBias = modff(i * rate + bias, &idxF);
        const size_t idx = (size_t)idxF;
        if (idx >= gsz) break;
        const float v =
            (1.f - curBias) * grain[idx] + curBias * (idx + 1 < gsz ? grain[idx + 1] : next);
        if (pcm && len < cap) pcm[len] = v;
        ++len;
        ++sz;
      }
      dt = 1. * sz / sr;
    }
    if (dt <= 0.) break;
    cursor += dt;
  }
  if (nsched) *nsched = ns;
  if (pcm && pcm16) { /* app.cpp:1209-1212 */
    const int64_t m_ = len < cap ? len : cap;
    for (int64_t i = 0; i < m_; ++i) pcm16[i] = (int16_t)(pcm[i] * 32767.);
  }
  return len;
}
