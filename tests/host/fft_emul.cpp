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

// The 32-points-per-thread plan of the synthesis kernel at fftN = 2048 (Fft32x32: two radix-32 stages, one exchange):
// 32 emulated threads, the buffer poisoned between the stages.
template <int DIR>
int check_32x32() {
  constexpr int NC = Fft32x32::NC;
  using C = cplx<float>;
  std::vector<double> in(2 * NC), ref(2 * NC), scratch(4 * NC);
  srand(77 + DIR);
  for (auto& v : in) v = (rand() / (double)RAND_MAX) * 2.0 - 1.0;
  mlxo_fft_plan* plan = mlxo_fft_plan_create(NC);
  mlxo_fft_c2c(plan, in.data(), ref.data(), scratch.data(), DIR);
  mlxo_fft_plan_destroy(plan);
  std::vector<C> regs(32 * 32), buf(Fft32x32::BUF, C{NAN, NAN});
  for (int t = 0; t < 32; ++t)
    for (int r = 0; r < 32; ++r) regs[t * 32 + r] = C{(float)in[2 * (t + 32 * r)], (float)in[2 * (t + 32 * r) + 1]};
  for (int t = 0; t < 32; ++t) Fft32x32::stage0<DIR>(*reinterpret_cast<C(*)[32]>(&regs[t * 32]), buf.data(), t);
  int bad = 0;
  for (int t = 0; t < 32; ++t) {
    C p1[31];
    for (int r = 1; r < 32; ++r) {
      const double a = DIR * 2.0 * M_PI * (double)((t * r) % NC) / NC;
      p1[r - 1] = C{(float)std::cos(a), (float)std::sin(a)};
    }
    Fft32x32::stage1<DIR>(*reinterpret_cast<C(*)[32]>(&regs[t * 32]), buf.data(), t, p1);
  }
  for (auto& v : buf) v = C{NAN, NAN};
  for (int t = 0; t < 32; ++t) Fft32x32::store(*reinterpret_cast<C(*)[32]>(&regs[t * 32]), buf.data(), t);
  double err = 0, nrm = 0;
  for (int i = 0; i < NC; ++i) {
    const C v = buf[pad32(i)];  // natural order, padded: what the overlap-add reads
    const C w = regs[(i % 32) * 32 + i / 32];
    if (!(v.x == w.x && v.y == w.y)) ++bad;
    const double dx = v.x - ref[2 * i], dy = v.y - ref[2 * i + 1];
    err += dx * dx + dy * dy;
    nrm += ref[2 * i] * ref[2 * i] + ref[2 * i + 1] * ref[2 * i + 1];
  }
  const double rel = std::sqrt(err / nrm);
  std::printf("32 x 32 plan, DIR %+d: rel rms %.2e  bad %d\n", DIR, rel, bad);
  return (bad != 0) || !(rel < 2e-6);
}

int main() {
  int bad = 0;
  bad |= check_32x32<+1>();
  bad |= check_32x32<-1>();
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
