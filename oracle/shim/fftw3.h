/* oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 * Declares exactly the FFTW3 symbols the reference's Spec uses (spec.hpp:4,20-22;
 * spec.cpp:11,13-15,60,103-105) so that /root/reference/spec.cpp compiles unmodified.
 * FFTW3 itself (libfftw3-dev, unpinned system package, README.md:9) is absent from this image;
 * the symbols are implemented in oracle/ref_spec.cpp on top of oracle/fft64.c. */
#ifndef MLXO_SHIM_FFTW3_H
#define MLXO_SHIM_FFTW3_H
#include <atomic>
#include <cstddef>
#include <mutex>
#include <unordered_map>
extern "C" {
typedef double fftw_complex[2];
typedef struct mlxo_shim_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
fftw_complex *fftw_alloc_complex(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
}
#endif
