/* oracle/fft64.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Not part of the shipped product path.
 *
 * Double-precision power-of-two FFT used by the CPU oracle in place of FFTW3 (the reference's
 * FFT engine, reference spec.hpp:4, spec.cpp:15,60 -- libfftw3-dev is a system package that is
 * absent from this image and from /root/reference).  Any correct double FFT agrees with FFTW to
 * ~1e-13 relative, far inside the 1e-4 parity budget.
 */
#ifndef MLXO_FFT64_H
#define MLXO_FFT64_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlxo_fft_plan mlxo_fft_plan;

/* n: complex length, power of two >= 1. */
mlxo_fft_plan *mlxo_fft_plan_create(int n);
void mlxo_fft_plan_destroy(mlxo_fft_plan *p);
int mlxo_fft_plan_size(const mlxo_fft_plan *p);

/* Complex transform, interleaved (re,im) doubles. dir = -1 forward (e^{-i..}), +1 backward
 * (unnormalised, like FFTW_BACKWARD). in and out may alias. scratch: 2*n doubles. */
void mlxo_fft_c2c(const mlxo_fft_plan *p, const double *in, double *out, double *scratch, int dir);

/* Real transforms of length N = 2*plan_size (plan is the N/2-point complex plan).
 * rfft: x[N] -> X[N/2+1] interleaved.  irfft: X[N/2+1] -> y[N], includes the 1/N factor.
 * scratch: 2*N doubles. */
void mlxo_rfft(const mlxo_fft_plan *p, const double *x, double *X, double *scratch);
void mlxo_irfft(const mlxo_fft_plan *p, const double *X, double *y, double *scratch);

#ifdef __cplusplus
}
#endif
#endif
