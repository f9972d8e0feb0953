// tests/host/fft_emul.cpp -- host-side sequential emulation of the device FFT (melonix_b200/csrc/fft.cuh).
// Verifies the Stockham index arithmetic, the in-register DFT-2/4/8/16 and the twiddle plan against
// the oracle's double FFT, without a GPU.  Built and run by tests/test_host_fft.py.
#include "../../melonix_b200/csrc/fft.cuh"
#include "../../oracle/fft64.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace mlx;

template <typename T, int NC, int DIR, int S, bool PRE1 = false>
struct Emul {
  using F = Fft<T, NC, DIR>;
  using P = FftPlan<NC>;
  static void go(std::vector<cplx<T>>& regs, std::vector<cplx<T>>& buf,
                 std::vector<FftTwiddles<T, NC, DIR, PRE1>>& tw) {
    if constexpr (S < P::NSTAGES) {
      if constexpr (S > 0) {
        // "barrier", then every thread loads, "barrier"
        for (int t = 0; t < P::TPF; ++t)
          F::load(*reinterpret_cast<cplx<T>(*)[16]>(&regs[t * 16]), buf.data(), t);
        // poison the buffer to prove that nothing stale is read later
        for (auto& v : buf) v = cplx<T>{T(NAN), T(NAN)};
      }
      for (int t = 0; t < P::TPF; ++t)
        F::template compute<S>(*reinterpret_cast<cplx<T>(*)[16]>(&regs[t * 16]), buf.data(), t, tw[t]);
      Emul<T, NC, DIR, S + 1, PRE1>::go(regs, buf, tw);
    }
  }
};

template <typename T, int NC, int DIR, bool PRE1 = false>
double check() {
  using P = FftPlan<NC>;
  std::vector<cplx<T>> table(NC);
  for (int m = 0; m < NC; ++m) {
    const double a = 2.0 * M_PI * m / NC;
    table[m] = cplx<T>{T(std::cos(a)), T(-std::sin(a))};
  }
  std::vector<double> in(2 * NC), ref(2 * NC), scratch(4 * NC);
  srand(NC + DIR);
  for (auto& v : in) v = (rand() / (double)RAND_MAX) * 2.0 - 1.0;
  mlxo_fft_plan* plan = mlxo_fft_plan_create(NC);
  mlxo_fft_c2c(plan, in.data(), ref.data(), scratch.data(), DIR);
  mlxo_fft_plan_destroy(plan);

  std::vector<cplx<T>> regs(P::TPF * 16), buf(P::BUF);
  std::vector<FftTwiddles<T, NC, DIR, PRE1>> tw(P::TPF);
  for (int t = 0; t < P::TPF; ++t) {
    tw[t].init(t, table.data());
    for (int m = 0; m < 16; ++m) {
      const int i = t + m * P::TPF;
      regs[t * 16 + m] = cplx<T>{T(in[2 * i]), T(in[2 * i + 1])};
    }
  }
  Emul<T, NC, DIR, 0, PRE1>::go(regs, buf, tw);
  double err = 0, nrm = 0;
  for (int t = 0; t < P::TPF; ++t)
    for (int m = 0; m < 16; ++m) {
      const int i = t + m * P::TPF;
      const double dx = regs[t * 16 + m].x - ref[2 * i], dy = regs[t * 16 + m].y - ref[2 * i + 1];
      err += dx * dx + dy * dy;
      nrm += ref[2 * i] * ref[2 * i] + ref[2 * i + 1] * ref[2 * i + 1];
    }
  return std::sqrt(err / nrm);
}

template <int NC>
int check_all() {
  int bad = 0;
  const double ef = check<float, NC, -1>(), eb = check<float, NC, +1>();
  const double df = check<double, NC, -1>(), db = check<double, NC, +1>();
  // stage-1 twiddle powers kept in registers: the same product tree, hence the same error
  const double pf = check<float, NC, -1, true>(), pb = check<float, NC, +1, true>();
  if (pf != ef || pb != eb) bad = 1;
  std::printf("NC=%5d  float fwd %.2e inv %.2e   double fwd %.2e inv %.2e\n", NC, ef, eb, df, db);
  if (!(ef < 2e-6 && eb < 2e-6 && df < 1e-14 && db < 1e-14)) bad = 1;
  return bad;
}

int main() {
  int bad = 0;
  bad |= check_all<256>();
  bad |= check_all<512>();
  bad |= check_all<1024>();
  bad |= check_all<2048>();
  bad |= check_all<4096>();
  bad |= check_all<8192>();
  bad |= check_all<16384>();
  std::printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
