// tests/host/spec_frame_emul.cpp -- sequential emulation of one thread group of the regular-hop
// Spec kernel (melonix_b200/csrc/spec_frame.cuh + fft.cuh) against the oracle restatement of
// Spec::internalGetSpec (oracle/spec_ref.c, reference spec.cpp:44-66).  Verifies, without a GPU,
// the window/tile addressing, the upper-slot exchange and that every bin in [0, N/2) is emitted
// exactly once.  Built and run by tests/test_host_side.py.
#include "../../melonix_b200/csrc/spec_frame.cuh"
#include "../../oracle/oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace mlx;

template <int NC, int S>
struct Stages {
  using F = Fft<float, NC, -1>;
  using P = FftPlan<NC>;
  static void go(std::vector<cplx<float>>& regs, std::vector<cplx<float>>& buf,
                 std::vector<FftTwiddles<float, NC, -1>>& tw) {
    if constexpr (S < P::NSTAGES) {
      if constexpr (S > 0) {
        for (int t = 0; t < P::TPF; ++t)
          F::load(*reinterpret_cast<cplx<float>(*)[16]>(&regs[t * 16]), buf.data(), t);
        for (auto& v : buf) v = cplx<float>{NAN, NAN};  // nothing stale may be read later
      }
      for (int t = 0; t < P::TPF; ++t)
        F::template compute<S>(*reinterpret_cast<cplx<float>(*)[16]>(&regs[t * 16]), buf.data(), t, tw[t]);
      Stages<NC, S + 1>::go(regs, buf, tw);
    }
  }
};

template <int N>
int check(int hop, long long n, int first_frame, int count) {
  using SF = SpecFrame<N>;
  constexpr int NC = N / 2, TPF = SF::TPF;
  using P = FftPlan<NC>;
  const int padf = 8192, padb = 16384;
  std::vector<float> track(padf + n + padb, 0.f);
  float* x = track.data() + padf;
  srand(N + hop);
  for (long long i = 0; i < n; ++i)
    x[i] = 0.4f * std::sin(0.01 * i * (1.0 + 1e-4 * i)) + 0.1f * (rand() / (float)RAND_MAX - 0.5f);
  std::vector<cplx<float>> table(NC), twr(NC / 2 + 1);
  for (int m = 0; m < NC; ++m) {
    const double a = 2.0 * M_PI * m / NC;
    table[m] = cplx<float>{(float)std::cos(a), (float)-std::sin(a)};
  }
  for (int k = 0; k <= NC / 2; ++k) {
    const double a = 2.0 * M_PI * k / N;
    twr[k] = cplx<float>{(float)std::cos(a), (float)-std::sin(a)};
  }
  const int ndec = N - hop;
  std::vector<float> decay(N + 1), win(N);
  for (int d = 0; d <= N; ++d) decay[d] = expf(-2.5e-4f * (float)d);
  for (int p = 0; p < N; ++p) win[p] = p < ndec ? decay[ndec - p] : 1.f;

  std::vector<FftTwiddles<float, NC, -1>> tw(TPF);
  for (int t = 0; t < TPF; ++t) tw[t].init(t, table.data());
  std::vector<cplx<float>> regs(TPF * 16), buf(P::BUF);
  std::vector<float> got(NC), ref(NC);
  std::vector<int> hits(NC);
  const float scale = 0.5f / (float)N;
  double err = 0, nrm = 0;
  int bad = 0;
  for (int f = first_frame; f < first_frame + count; ++f) {
    const long long start = (long long)f * hop, end = start + hop;
    const float* cur = x + end - N;  // inside the zero padding when outside the track
    for (int t = 0; t < TPF; ++t)
      SF::load(*reinterpret_cast<cplx<float>(*)[16]>(&regs[t * 16]), cur, t,
               [&](int p) { return *reinterpret_cast<const cplx<float>*>(win.data() + p); });
    for (auto& v : buf) v = cplx<float>{NAN, NAN};
    Stages<NC, 0>::go(regs, buf, tw);
    for (auto& v : buf) v = cplx<float>{NAN, NAN};
    for (int t = 0; t < TPF; ++t)
      SF::stage_upper(*reinterpret_cast<cplx<float>(*)[16]>(&regs[t * 16]), buf.data(), t);
    std::fill(hits.begin(), hits.end(), 0);
    for (int t = 0; t < TPF; ++t)
      SF::emit_bins(*reinterpret_cast<cplx<float>(*)[16]>(&regs[t * 16]), buf.data(), t, scale,
                    [&](int k) { return twr[k]; },
                    [&](int k, float v) {
                      if (k < 0 || k >= NC) { ++bad; return; }
                      got[k] = v;
                      ++hits[k];
                    });
    for (int k = 0; k < NC; ++k)
      if (hits[k] != 1) ++bad;
    mlxo_spec_frame(x, n, (int)start, (int)end, N, ref.data());
    for (int k = 0; k < NC; ++k) {
      if (!(got[k] == got[k])) ++bad;  // NaN: a stale slot was read
      const double d = (double)got[k] - ref[k];
      err += d * d;
      nrm += (double)ref[k] * ref[k];
    }
  }
  const double rel = std::sqrt(err / (nrm > 0 ? nrm : 1));
  std::printf("N=%5d hop=%5d frames=%d  rel rms %.2e  bad %d\n", N, hop, count, rel, bad);
  return (bad != 0) || !(rel < 2e-6);
}

int main() {
  int bad = 0;
  bad |= check<512>(128, 3000, 0, 26);   // covers the leading zero padding, the interior and the tail
  bad |= check<1024>(256, 6000, 0, 26);
  bad |= check<1024>(100, 3000, 0, 34);  // hop not N/4
  bad |= check<2048>(512, 9000, 0, 20);
  bad |= check<2048>(2048, 9000, 0, 6);  // hop == N: nothing decays
  bad |= check<4096>(1024, 20000, 0, 22);
  bad |= check<8192>(2048, 30000, 0, 17);
  std::printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
