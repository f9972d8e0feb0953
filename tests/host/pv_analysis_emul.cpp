// tests/host/pv_analysis_emul.cpp -- the analysis-bin arithmetic of the phase vocoder
// (melonix_b200/csrc/pv_analysis.cuh, the code the GPU runs; the host build uses IEEE division / sqrt
// where the device uses the SFU approximations) against the oracle's evaluation of PV-spec A.3
// (oracle/pv_ref.c:99-114): d = arg(X_f conj(X_{f-1}) (-i)^k) in double.
//
// What is checked: the wrapped phase advance d' = d32 (-+2^32 on a flip) lies on the SAME side of the
// +-pi cut as the oracle's atan2 -- for random bins and for sequences built to sit within 1e-3 ... 1e-15
// rad of the cut -- and agrees with it to 1e-6 rad; the magnitude to 1e-6 relative; and the integer
// phases telescope (sum of d' over frames = P_last - P_first - F k 2^30 mod 2^32, exactly).
// Built and run by tests/test_host_side.py.
#include "../../melonix_b200/csrc/pv_analysis.cuh"

#include <cmath>
#include <cstdio>
#include <random>

using namespace mlx;

struct Ref {
  double dd, zi_abs, scale;
  bool gated;
};
// oracle/pv_ref.c:99-114, verbatim arithmetic (plain double products, rotation by (-i)^k, atan2)
static Ref ref_bin(double a, double b, double c, double d, int k, bool real_bin) {
  double zr = a * c + b * d, zi = b * c - a * d;
  double t;
  switch (k & 3) {
    case 1: t = zr; zr = zi; zi = -t; break;
    case 2: zr = -zr; zi = -zi; break;
    case 3: t = zr; zr = -zi; zi = t; break;
    default: break;
  }
  if (real_bin) zi = 0.0;
  Ref r;
  r.gated = zr * zr + zi * zi <= 1e-36;
  r.dd = r.gated ? 0.0 : std::atan2(zi, zr);
  r.zi_abs = std::fabs(zi);
  r.scale = std::fabs(a * c) + std::fabs(b * d) + std::fabs(b * c) + std::fabs(a * d);
  return r;
}

int main() {
  std::mt19937_64 rng(777);
  std::uniform_real_distribution<double> uni(-1.0, 1.0), ang(-M_PI, M_PI);
  const double twopi = 2.0 * M_PI;
  long long n = 0, wrong_side = 0, indeterminate = 0, bad_val = 0, bad_mag = 0, bad_tel = 0;
  double max_err = 0;
  const double eps_list[] = {1e-3, 1e-4, 1e-6, 1e-9, 1e-12, 1e-15, 0.3, 1.0, 2.0, 3.0};
  for (int trial = 0; trial < 4000; ++trial) {
    const int k = (int)(rng() % 1025);
    // a track of 64 frames of this bin; magnitudes over 12 decades, phases either random or steered so that
    // the advance lands eps away from the cut on alternating sides
    double pa = 1.0, pb = 0.0;  // X_{-1} = 1
    uint32_t p_prev = 0u, p_first = 0u;
    float m_prev = 1.f;
    uint32_t sum = 0u;
    double phi = 0.0;
    const bool steer = trial & 1;
    const int F = 64;
    for (int f = 0; f < F; ++f, ++n) {
      const double mag = std::pow(10.0, 6.0 * uni(rng) - 2.0);
      if (steer) {
        const double eps = eps_list[rng() % 10] * ((rng() & 1) ? 1.0 : -1.0);
        phi += M_PI * 0.5 * k + M_PI - eps;  // advance = k pi/2 + (pi - eps)  ->  d = pi - eps (mod 2 pi)
      } else {
        phi = ang(rng);
      }
      const double a = mag * std::cos(phi), b = mag * std::sin(phi);
      float gm;
      int d32;
      bool flip;
      const uint32_t before = p_prev;
      analysis_bin(a, b, pa, pb, p_prev, m_prev, k, false, gm, d32, flip);
      if (f == 0) p_first = before;
      const Ref r = ref_bin(a, b, pa, pb, k, false);
      const long long dprime = (long long)d32 + (flip ? (d32 < 0 ? 4294967296LL : -4294967296LL) : 0LL);
      const double got = (double)dprime * twopi / 4294967296.0;
      if (!r.gated) {
        const double err = std::fabs(got - r.dd);
        if (err > 1.0) {  // the other side of the cut
          // legitimate only when Im Z is below the rounding noise of its own evaluation
          if (r.zi_abs <= 8.0 * 2.3e-16 * r.scale) ++indeterminate; else ++wrong_side;
        } else {
          if (err > 1e-6) ++bad_val;
          if (err > max_err) max_err = err;
        }
        if (std::fabs((double)gm - mag) > 1e-6 * mag) ++bad_mag;
      }
      sum += (uint32_t)d32;  // flips add multiples of 2^32: invisible mod 2^32
      pa = a;
      pb = b;
    }
    // telescoping (no gated frame in these tracks): sum d32 = P_last - P_before_first - F k 2^30  (mod 2^32)
    const uint32_t expect = p_prev - p_first - (uint32_t)((unsigned long long)F * ((unsigned long long)k << 30));
    if (sum != expect) ++bad_tel;
  }
  // the purely real bins: d in {0, +pi} exactly as the spec defines them
  long long bad_real = 0;
  {
    uint32_t pp = 0u;
    float pm = 1.f;
    double prev = 1.0;
    for (int f = 0; f < 2000; ++f) {
      const double x = uni(rng) * std::pow(10.0, 3.0 * uni(rng));
      const MagD m = analysis_real_bin(x, pp, pm);
      const Ref r = ref_bin(x, 0.0, prev, 0.0, 0, true);
      const long long dprime = (long long)m.d + ((pv_f2bits(m.mag) >> 31) ? (m.d < 0 ? 4294967296LL : -4294967296LL) : 0LL);
      const double got = (double)dprime * twopi / 4294967296.0;
      if (!r.gated && std::fabs(got - r.dd) > 1e-9) ++bad_real;
      prev = x;
    }
  }
  std::printf("%lld bin-frames: wrong side of the cut %lld (numerically indeterminate: %lld), value errors %lld "
              "(max %.2e rad), magnitude errors %lld, telescoping failures %lld, real-bin errors %lld\n",
              n, wrong_side, indeterminate, bad_val, max_err, bad_mag, bad_tel, bad_real);
  const bool bad = wrong_side || bad_val || bad_mag || bad_tel || bad_real;
  std::printf(bad ? "FAIL\n" : "OK\n");
  return bad ? 1 : 0;
}
