// tests/host/grain_seg_emul.cpp -- the bit logic of the device grain segmentation
// (melonix_b200/csrc/grain_seg.cuh) run sequentially on the CPU against the oracle's restatement of
// App::preproc (oracle/grain_ref.c, reference app.cpp:156-235).  Built and run by
// tests/test_host_side.py.
#include "../../melonix_b200/csrc/grain_seg.cuh"
#include "../../oracle/oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace mlx;

static int run_case(const char* name, const std::vector<float>& w) {
  const long long n = (long long)w.size();
  const long long nwords = (n + 31) / 32 + 1;
  std::vector<uint32_t> L(nwords + 2, 0u), R(nwords + 2, 0u), z7(nwords + kGrainWindowWords + 8, 0u),
      z3(nwords + kGrainWindowWords + 2, 0u);
  auto Lw = [&](long long k) { return (k < 0 || k >= nwords) ? 0u : L[k]; };
  auto Rw = [&](long long k) { return (k < 0 || k >= nwords) ? 0u : R[k]; };
  for (long long i = 0; i < n; ++i) {
    if (!(w[i] >= 0)) L[i >> 5] |= 1u << (i & 31);
    if (!(w[i] < 0)) R[i >> 5] |= 1u << (i & 31);
  }
  for (long long k = 0; k < nwords; ++k) {
    z7[k] = seg_cross_word(Lw(k - 1), Lw(k), Rw(k), Rw(k + 1), k, n, 7);
    z3[k] = seg_cross_word(Lw(k - 1), Lw(k), Rw(k), Rw(k + 1), k, n, 3);
  }
  // the chain, as the device walks it
  std::vector<int> gs, gl;
  const int lim = (int)(n - kGrainPreferred - 1);
  int start = 0;
  while (start < lim) {
    const int c = start + kGrainPreferred;
    const int w0 = (c - kGrainHalfSpan) >> 5;
    uint32_t best = 0xffffffffu;
    for (int i = 0; i < kGrainWindowWords; ++i) {
      const uint32_t key = seg_word_key(z7[w0 + i], (w0 + i) * 32, c);
      if (key < best) best = key;
    }
    int idx = -1;
    if (best != 0xffffffffu) {
      idx = seg_key_index(best, c);
    } else {
      const long long s = (long long)start + kGrainPreferred + kGrainPreferred / 2;
      for (long long k = s >> 5; k < nwords && idx < 0; ++k) {
        uint32_t word = z3[k];
        if (k == (s >> 5)) word &= 0xffffffffu << (int)(s & 31);
        if (word) idx = (int)(k * 32 + __builtin_ctz(word));
      }
      if (idx < 0) break;
    }
    gs.push_back(start);
    gl.push_back(idx - start);
    start = idx;
  }
  // the same chain with the warp-cooperative helpers (lanes emulated one after the other)
  std::vector<int> gs2, gl2;
  start = 0;
  while (start < lim) {
    const int c = start + kGrainPreferred, lo = c - kGrainHalfSpan;
    const int sbit = lo & 31, w0 = lo >> 5;
    int lane_ge = -1, lane_lt = -1;
    unsigned long long w_ge = 0, w_lt = 0;
    for (int lane = 0; lane < 32; ++lane) {
      const unsigned long long bits = (unsigned long long)z7[w0 + 2 * lane] | ((unsigned long long)z7[w0 + 2 * lane + 1] << 32);
      const SegSplit sp = seg_lane_split(lane < 24 ? bits : 0xdeadbeefdeadbeefull, lane, sbit);
      if (sp.ge && lane_ge < 0) { lane_ge = lane; w_ge = sp.ge; }
      if (sp.lt) { lane_lt = lane; w_lt = sp.lt; }
    }
    const int rel = seg_pick(lane_ge, w_ge, lane_lt, w_lt, sbit);
    int idx = -1;
    if (rel >= 0) {
      idx = w0 * 32 + rel;
    } else {
      const long long s0 = (long long)start + kGrainPreferred + kGrainPreferred / 2;
      for (long long k = s0 >> 5; k < nwords && idx < 0; ++k) {
        uint32_t word = z3[k];
        if (k == (s0 >> 5)) word &= 0xffffffffu << (int)(s0 & 31);
        if (word) idx = (int)(k * 32 + __builtin_ctz(word));
      }
      if (idx < 0) break;
    }
    gs2.push_back(start);
    gl2.push_back(idx - start);
    start = idx;
  }
  int bad2 = gs2 != gs || gl2 != gl;
  std::vector<int32_t> os(n / 700 + 8), ol(n / 700 + 8);
  const int cnt = mlxo_grain_segment(w.data(), n, os.data(), ol.data(), (int)os.size());
  int bad = (cnt != (int)gs.size()) || bad2;
  for (int i = 0; i < cnt && i < (int)gs.size() && !bad; ++i) bad = (os[i] != gs[i]) || (ol[i] != gl[i]);
  std::printf("%-28s n=%8lld grains oracle %6d emul %6zu %s\n", name, n, cnt, gs.size(), bad ? "MISMATCH" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  const int fs = 48000;
  auto tone = [&](double f, double sec, double noise, unsigned seed) {
    std::vector<float> w((size_t)(sec * fs));
    srand(seed);
    for (size_t i = 0; i < w.size(); ++i)
      w[i] = (float)(0.4 * std::sin(2 * M_PI * f * i / fs) + 0.2 * std::sin(2 * M_PI * 2.01 * f * i / fs + 1.0) +
                     noise * (rand() / (double)RAND_MAX - 0.5));
    return w;
  };
  bad |= run_case("tone 220 Hz", tone(220, 20, 0, 1));
  bad |= run_case("tone 220 Hz + noise", tone(220, 20, 0.05, 2));
  bad |= run_case("tone 33 Hz (sparse)", tone(33, 20, 0, 3));
  bad |= run_case("tone 9 Hz (fallback)", tone(9, 30, 0, 4));
  bad |= run_case("tone 9 Hz + noise", tone(9, 30, 0.02, 5));
  bad |= run_case("white noise", tone(0, 10, 1.0, 6));
  bad |= run_case("silence", std::vector<float>(100000, 0.f));
  bad |= run_case("negative DC", std::vector<float>(100000, -0.25f));
  bad |= run_case("short clip 1400", tone(220, 1400.0 / fs, 0, 7));
  bad |= run_case("short clip 1502", tone(220, 1502.0 / fs, 0, 8));
  bad |= run_case("short clip 3100", tone(220, 3100.0 / fs, 0, 9));
  bad |= run_case("empty", std::vector<float>());
  {
    std::vector<float> w = tone(440, 5, 0, 10);
    for (size_t i = 0; i < w.size(); i += 97) w[i] = -0.0f;     // -0.0 counts as ">= 0"
    for (size_t i = 5; i < w.size(); i += 1013) w[i] = NAN;     // NaN passes both sign tests
    bad |= run_case("tone with -0.0 and NaN", w);
  }
  std::printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
