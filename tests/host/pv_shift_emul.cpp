// tests/host/pv_shift_emul.cpp -- the bin-shift phase increment of the analysis kernel
// (melonix_b200/csrc/pv_shift.cuh, the code the GPU runs) against
//   (1) a 128-bit evaluation of its defining formula  inc = (r_fix (kh 2^30 + d') + 2^25) >> 26 mod 2^32,
//   (2) the PV-spec formula in double  inc = llrint(frac(r nu / 4) 2^32),  nu = kh + 4 d' / 2^32
//       (SURVEY Appendix A.6): equal up to the one rounding both perform.
// Built and run by tests/test_host_side.py.
#include "../../melonix_b200/csrc/pv_shift.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

using namespace mlx;

int main() {
  std::mt19937_64 rng(12345);
  const int zslot = 1089;
  long long bad_exact = 0, bad_spec = 0, n = 0;
  const float rates[] = {0.25f, 0.5f, 0.7491536f, 0.8908987f, 1.0f, 1.0594631f, 1.1892071f, 1.4983071f, 2.0f, 3.999f, 4.0f};
  for (float rate : rates) {
    const uint32_t r_fix = (uint32_t)((double)rate * 67108864.0);  // rate * 2^26, exact (24-bit float)
    for (int it = 0; it < 400000; ++it, ++n) {
      const int kh = (int)(rng() % 4097);
      int d;
      switch (it % 8) {
        case 0: d = INT32_MIN; break;
        case 1: d = INT32_MAX; break;
        case 2: d = 0; break;
        case 3: d = -1; break;
        default: d = (int)(uint32_t)rng();
      }
      const bool flip = (rng() & 1) != 0;
      const ShiftConst c = make_shift_const(0, (uint32_t)kh | ((uint32_t)kh << 16), r_fix, zslot);
      if (c.slot != (uint32_t)fft_pad(kh)) ++bad_exact;
      const uint32_t mb = flip ? 0x80000000u | 0x3f000000u : 0x3f000000u;
      const uint32_t got = shift_inc(c.base, d, mb, (int)r_fix);
      // (1) exact
      const __int128 dprime = (__int128)d + (flip ? (d < 0 ? ((__int128)1 << 32) : -((__int128)1 << 32)) : 0);
      const __int128 turns = ((__int128)kh << 30) + dprime;  // nu / 4 in 2^-32 turns
      const __int128 val = (__int128)r_fix * turns + ((__int128)1 << 25);
      const uint32_t exact = (uint32_t)(unsigned long long)(val >> 26);
      if (got != exact) ++bad_exact;
      // the 32-bit-constant form used by the kernels' fast paths, and where its record lives
      const ShiftConstA ca = make_shift_const_a(0, (uint32_t)kh | ((uint32_t)kh << 16), r_fix, 4096, true, 16u * zslot);
      if (shift_inc_a(ca.A, d, mb, (int)r_fix) != exact) ++bad_exact;
      if (ca.off != 16u * (uint32_t)fft_pad(kh <= 2048 ? kh : 4096 - kh) + (kh <= 2048 ? 0u : 8u)) ++bad_exact;
      // (2) the spec's double formula: one rounding each, so they agree to one count
      const double x = (double)rate * ((double)kh / 4.0 + (double)(long long)dprime / 4294967296.0);
      const double fr = x - std::floor(x);
      const uint32_t spec = (uint32_t)(unsigned long long)std::llrint(fr * 4294967296.0);
      const uint32_t diff = got - spec;
      if (!(diff == 0u || diff == 1u || diff == 0xffffffffu)) ++bad_spec;
    }
    // empty K_j: s_nu = j, inc = frac(j / 4) 2^32, from an all-zero record
    for (int j = 0; j < 64; ++j) {
      const ShiftConst c = make_shift_const(j, 1u /* klo = 1, khi = 0 */, r_fix, zslot);
      if (c.slot != (uint32_t)zslot || shift_inc(c.base, 0, 0u, (int)r_fix) != ((uint32_t)(j & 3) << 30)) ++bad_exact;
      const ShiftConstA ca = make_shift_const_a(j, 1u, r_fix, 4096, true, 16u * zslot);
      if (ca.off != 16u * zslot || shift_inc_a(ca.A, 0, 0u, (int)r_fix) != ((uint32_t)(j & 3) << 30)) ++bad_exact;
    }
  }
  std::printf("%lld increments: %lld differ from the exact formula, %lld differ from the spec formula by more than one count\n",
              n, bad_exact, bad_spec);
  std::printf(bad_exact || bad_spec ? "FAIL\n" : "OK\n");
  return bad_exact || bad_spec ? 1 : 0;
}
