/* oracle/ref_spec.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's own `Spec` class (compiled UNMODIFIED from /root/reference/spec.cpp by
 * oracle/Makefile target `ref`) and implements the six FFTW entry points it needs on top of the
 * oracle's double FFT.  Used to pin oracle/spec_ref.c against the real reference code and as the
 * `kind: "reference"` CPU baseline of the Spec-only configuration.
 */
#include "fft64.h"
#include <spec.hpp> /* the reference header, found via -I /root/reference */

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

struct mlxo_shim_plan_s {
  int n;
  fftw_complex *in, *out;
  int sign;
  mlxo_fft_plan *plan;
  double *scratch;
};

extern "C" {
fftw_complex *fftw_alloc_complex(size_t n) {
  void *p = nullptr;
  if (posix_memalign(&p, 64, n * sizeof(fftw_complex))) return nullptr;
  return static_cast<fftw_complex *>(p);
}
void fftw_free(void *p) { free(p); }
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned) {
  auto *pl = new mlxo_shim_plan_s{n, in, out, sign, mlxo_fft_plan_create(n), nullptr};
  pl->scratch = static_cast<double *>(malloc(sizeof(double) * 2 * static_cast<size_t>(n)));
  return pl;
}
void fftw_execute(const fftw_plan p) {
  mlxo_fft_c2c(p->plan, reinterpret_cast<const double *>(p->in), reinterpret_cast<double *>(p->out),
               p->scratch, p->sign);
}
void fftw_destroy_plan(fftw_plan p) {
  mlxo_fft_plan_destroy(p->plan);
  free(p->scratch);
  delete p;
}

/* Runs `count` jobs through the reference Spec exactly as SpecCache does (spec-cache.cpp:63-72):
 * call getSpec until it stops returning {} (the worker fills it asynchronously).
 * out: [count][16384] floats (SpectrSize/2, spec.cpp:8).  Returns the spectrum length or <0. */
int mlxo_ref_spec_run(const float *wav, long long n, const int *start_end, int count, float *out) {
  std::vector<float> copy(wav, wav + n); /* Spec takes std::span<float> (non-const) */
  Spec spec(std::span<float>{copy.data(), copy.size()});
  /* the reference starts its worker in the mem-init list before `plan` exists (SURVEY.md
     section 5); give the constructor time to finish planning before the first job is queued. */
  int len = 0;
  /* enqueue everything first (bounded by MaxRanges = 4000, range.hpp:4) then poll */
  const int window = 2000;
  for (int base = 0; base < count; base += window) {
    const int hi = base + window < count ? base + window : count;
    std::vector<char> done(hi - base, 0);
    int remaining = hi - base;
    while (remaining > 0) {
      for (int j = base; j < hi; ++j) {
        if (done[j - base]) continue;
        auto s = spec.getSpec(start_end[2 * j], start_end[2 * j + 1]);
        if (s.empty()) continue;
        len = static_cast<int>(s.size());
        std::memcpy(out + static_cast<size_t>(j) * s.size(), s.data(), s.size() * sizeof(float));
        done[j - base] = 1;
        --remaining;
      }
      if (remaining > 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
  }
  return len;
}
}
